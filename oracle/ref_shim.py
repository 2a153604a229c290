"""Import the UNMODIFIED reference `models/` package from /root/reference (build container only).

TEST INFRASTRUCTURE.  The reference needs timm, hydra, omegaconf and submitit, none of which are
installed here; only four tiny stand-ins are required for `models/`:

  timm.models.layers.trunc_normal_  -> torch.nn.init.trunc_normal_      (init only, fusion.py:13,25)
  hydra.utils.instantiate           -> import `_target_`, call with merged kwargs (base_model.py:22-25)
  omegaconf.{DictConfig, OmegaConf} -> attribute-access dict             (future_prediction.py:24)
  submitit                          -> empty module (unused import via common/utils.py:10)

`transformers` must be imported before the timm stub is registered (its availability probe chokes on a
spec-less module).  fusion.py hard-codes device 'cuda' for the T-SA / CA fusers (:170,187,254-255);
`cuda_redirect()` maps that to the CPU for the duration of a forward.

/root/reference does not exist on the GPU box: nothing here may be used by `-m gpu` tests, smoke() or
bench.py.  available() says whether the reference can be imported.
"""
from __future__ import annotations

import contextlib
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("AFFT_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "base_model.py"))


class DictConfig(dict):
    """Attribute-access dict standing in for omegaconf.DictConfig."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def to_cfg(obj):
    if isinstance(obj, dict):
        return DictConfig({k: to_cfg(v) for k, v in obj.items()})
    if isinstance(obj, (list, tuple)):
        return [to_cfg(v) for v in obj]
    return obj


def _instantiate(cfg, *args, _recursive_=True, **kwargs):
    params = dict(cfg)
    target = params.pop("_target_")
    params.update(kwargs)
    mod_name, cls_name = target.rsplit(".", 1)
    cls = getattr(importlib.import_module(mod_name), cls_name)
    return cls(*args, **params)


_installed = False


def install_stubs():
    global _installed
    if _installed:
        return
    import torch
    import transformers  # noqa: F401  (must precede the timm stub)

    om = types.ModuleType("omegaconf")
    om.DictConfig = DictConfig
    om.ListConfig = list

    class OmegaConf:  # noqa: D401
        @staticmethod
        def get_type(x):
            return type(x)

        @staticmethod
        def to_yaml(x):
            return repr(x)

    om.OmegaConf = OmegaConf
    sys.modules.setdefault("omegaconf", om)

    hy = types.ModuleType("hydra")
    hyu = types.ModuleType("hydra.utils")
    hyu.instantiate = _instantiate
    hy.utils = hyu
    sys.modules.setdefault("hydra", hy)
    sys.modules.setdefault("hydra.utils", hyu)

    tm = types.ModuleType("timm")
    tmm = types.ModuleType("timm.models")
    tml = types.ModuleType("timm.models.layers")
    tml.trunc_normal_ = torch.nn.init.trunc_normal_
    tm.models = tmm
    tmm.layers = tml
    sys.modules.setdefault("timm", tm)
    sys.modules.setdefault("timm.models", tmm)
    sys.modules.setdefault("timm.models.layers", tml)

    sys.modules.setdefault("submitit", types.ModuleType("submitit"))
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed = True


@contextlib.contextmanager
def cuda_redirect():
    """Run reference code that says .to('cuda') / device='cuda' on the CPU."""
    import torch
    orig_to, orig_arange = torch.Tensor.to, torch.arange

    def to(self, *a, **k):
        a = tuple("cpu" if (isinstance(x, str) and x == "cuda") else x for x in a)
        if k.get("device") == "cuda":
            k["device"] = "cpu"
        return orig_to(self, *a, **k)

    def arange(*a, **k):
        if k.get("device") == "cuda":
            k["device"] = "cpu"
        return orig_arange(*a, **k)

    torch.Tensor.to, torch.arange = to, arange
    try:
        yield
    finally:
        torch.Tensor.to, torch.arange = orig_to, orig_arange


def build_reference_model(model_cfg: dict, num_classes: dict):
    """The reference's BaseModel(cfg.model, num_classes, class_mappings={}) - models/base_model.py:15-29."""
    if not available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    install_stubs()
    ref_models = importlib.import_module("models.base_model")
    if not os.path.abspath(ref_models.__file__).startswith(os.path.abspath(REFERENCE_ROOT)):
        raise RuntimeError(f"`models` resolved to {ref_models.__file__}, not the reference")
    return ref_models.BaseModel(to_cfg(model_cfg), num_classes=dict(num_classes), class_mappings={})


def reference_forward(model, feats_6d: dict):
    """test.py:72-82 - model(feature_dict, mixup_fn=None, target=None, ...) in eval / no_grad."""
    import torch
    kwargs = dict(mixup_fn=None, target=None, target_subclips=None, target_subclips_ignore_index=None)
    model.eval()
    with torch.no_grad(), cuda_redirect():
        outputs, _ = model({m: t.clone() for m, t in feats_6d.items()}, **kwargs)
    return outputs
