"""Mirror of reference models/future_prediction.py for the early-fusion path (CMFPEarly + BaseFuturePredictor).

The module tree and parameter names equal the reference's (``mapping.<mod>.mapping.0``, ``fuser.*``,
``dim_encoder``, ``dim_decoder``, ``future_predictor.gpt_model.*``, ``classifiers.<cls>.all-fused.1``), so
checkpoints load with the reference's own ``init_model``.  ``CMFPEarly.forward`` hands the whole chain
feature_mapping -> fuser -> dim_encoder -> GPT-2 -> dim_decoder -> classifiers to one native call.
"""
from __future__ import annotations

import importlib
import logging
import math
from collections.abc import Mapping
from typing import Dict

import torch
import torch.nn as nn
from torch import Tensor

from .. import _capi
from ..engine import Engine

PAST_LOGITS_PREFIX = 'past_'


# ------------------------------------------------------------------------------------------------
# config helpers (accept omegaconf DictConfig, plain dicts or attribute objects)
# ------------------------------------------------------------------------------------------------
def cfg_get(cfg, key, default=None):
    if isinstance(cfg, Mapping):
        return cfg.get(key, default)
    return getattr(cfg, key, default)


def cfg_items(cfg):
    return cfg.items() if isinstance(cfg, Mapping) else vars(cfg).items()


_TARGETS = {
    "models.fusion.ModalTokenCMFuser": ("afft_b200.models.fusion", "ModalTokenCMFuser"),
    "models.fusion.CMFuser": ("afft_b200.models.fusion", "CMFuser"),
    "models.fusion.TemporalCMFuser": ("afft_b200.models.fusion", "TemporalCMFuser"),
    "models.fusion.TemporalCrossAttentFuser": ("afft_b200.models.fusion", "TemporalCrossAttentFuser"),
    "models.fusion.MATT": ("afft_b200.models.fusion", "MATT"),
    "models.feature_mapping.Linear": ("afft_b200.models.feature_mapping", "Linear"),
    "models.feature_mapping.GatedLinear": ("afft_b200.models.feature_mapping", "GatedLinear"),
    "models.feature_mapping.NonLinear": ("afft_b200.models.feature_mapping", "NonLinear"),
    "models.future_prediction.BaseFuturePredictor": ("afft_b200.models.future_prediction", "BaseFuturePredictor"),
    "models.future_prediction.CMFPEarly": ("afft_b200.models.future_prediction", "CMFPEarly"),
    "models.future_prediction.IndividualFuturePrediction": ("afft_b200.models.future_prediction", "IndividualFuturePrediction"),
    "models.future_prediction.CMFPScoreFusion": ("afft_b200.models.future_prediction", "CMFPScoreFusion"),
}


def instantiate(cfg, *args, **kwargs):
    """Minimal ``hydra.utils.instantiate``: resolve ``_target_`` (the reference's ``models.*`` names map onto
    this package, anything else is imported as is) and call it with the remaining keys + kwargs."""
    params = {k: v for k, v in cfg_items(cfg) if k not in ("_target_", "_recursive_")}
    kwargs.pop("_recursive_", None)
    params.update(kwargs)
    target = cfg_get(cfg, "_target_")
    if target.startswith("afft_b200.models."):
        target = target[len("afft_b200."):]
    if target in _TARGETS:
        mod_name, cls_name = _TARGETS[target]
    elif target.startswith("models."):
        raise NotImplementedError(f"{target} is outside the hot path rebuilt here (SURVEY.md section 8f)")
    else:
        mod_name, cls_name = target.rsplit(".", 1)
    return getattr(importlib.import_module(mod_name), cls_name)(*args, **params)


# ------------------------------------------------------------------------------------------------
# GPT-2 parameter tree (transformers.GPT2Model layout, wte deleted - reference future_prediction.py:372-385)
# ------------------------------------------------------------------------------------------------
class Conv1D(nn.Module):
    """transformers' Conv1D: y = x @ weight + bias with weight stored [in, out]."""

    def __init__(self, nf, nx):
        super().__init__()
        self.nf = nf
        self.weight = nn.Parameter(torch.empty(nx, nf))
        self.bias = nn.Parameter(torch.zeros(nf))
        nn.init.normal_(self.weight, std=0.02)


class _GPT2Attention(nn.Module):
    def __init__(self, n_embd, attn_pdrop, resid_pdrop):
        super().__init__()
        self.c_attn = Conv1D(3 * n_embd, n_embd)
        self.c_proj = Conv1D(n_embd, n_embd)
        self.attn_dropout = nn.Dropout(attn_pdrop)
        self.resid_dropout = nn.Dropout(resid_pdrop)


class _GPT2MLP(nn.Module):
    def __init__(self, n_embd, resid_pdrop):
        super().__init__()
        self.c_fc = Conv1D(4 * n_embd, n_embd)
        self.c_proj = Conv1D(n_embd, 4 * n_embd)
        self.dropout = nn.Dropout(resid_pdrop)


class _GPT2Block(nn.Module):
    def __init__(self, n_embd, attn_pdrop, resid_pdrop, eps=1e-5):
        super().__init__()
        self.ln_1 = nn.LayerNorm(n_embd, eps=eps)
        self.attn = _GPT2Attention(n_embd, attn_pdrop, resid_pdrop)
        self.ln_2 = nn.LayerNorm(n_embd, eps=eps)
        self.mlp = _GPT2MLP(n_embd, resid_pdrop)


class _GPT2Model(nn.Module):
    def __init__(self, n_embd, n_layer, n_head, n_positions, embd_pdrop, resid_pdrop, attn_pdrop):
        super().__init__()
        self.n_embd, self.n_layer, self.n_head, self.n_positions = n_embd, n_layer, n_head, n_positions
        self.wpe = nn.Embedding(n_positions, n_embd)
        nn.init.normal_(self.wpe.weight, std=0.02)
        self.drop = nn.Dropout(embd_pdrop)
        self.h = nn.ModuleList([_GPT2Block(n_embd, attn_pdrop, resid_pdrop) for _ in range(n_layer)])
        self.ln_f = nn.LayerNorm(n_embd, eps=1e-5)
        for blk in self.h:  # GPT-2's scaled init of the residual projections
            nn.init.normal_(blk.attn.c_proj.weight, std=0.02 / math.sqrt(2 * n_layer))
            nn.init.normal_(blk.mlp.c_proj.weight, std=0.02 / math.sqrt(2 * n_layer))


class BaseFuturePredictor(nn.Module):
    """reference models/future_prediction.py:354-415.  Holds the GPT-2 weights (gelu_new MLP, n_inner = 4 n_embd,
    1024 positions, LayerNorm eps 1e-5 - the GPT2Config defaults the reference relies on)."""

    def __init__(self, in_features, inter_dim=2048, n_layer=6, n_head=4, embd_pdrop=0.1, resid_pdrop=0.1,
                 attn_pdrop=0.1, output_attentions=False, dimension_mapping=False):
        super().__init__()
        if dimension_mapping:
            raise NotImplementedError("dimension_mapping inside GPT-2 is deprecated in the reference and not supported")
        self.in_features = in_features
        self.output_attentions = output_attentions
        self.encoder = nn.Identity()
        self.decoder = nn.Identity()
        self.gpt_model = _GPT2Model(inter_dim, n_layer, n_head, 1024, embd_pdrop, resid_pdrop, attn_pdrop)

    precision = "bf16"  # 'bf16' | 'fp16' | 'strict' (set by the owning head)
    max_batch = 64

    def forward(self, feats, output_len: int = 1):
        """The reference's inner seam ``predictor(feats (B, T, C), output_len) -> ((B, T + output_len - 1, C), {})``
        (models/future_prediction.py:387-415; encoder / decoder are Identity), evaluated by a native GPT-2-only handle
        (AFFT_STAGE_GPT): position embedding add, the GPT-2 blocks, ln_f, and the KV-cache roll-out for output_len > 1.
        Inference only."""
        if self.training:
            raise NotImplementedError("the standalone predictor seam is inference-only; training runs through CMFPEarly.forward")
        if feats.device.type != "cuda":
            raise _capi.AfftError("afft_b200 runs on CUDA devices only (sm_100a); there is no CPU path")
        B, T, G = feats.shape
        gpt = self.gpt_model
        if G != gpt.n_embd:
            raise ValueError(f"predictor input width {G} != inter_dim {gpt.n_embd}")
        engines = self.__dict__.setdefault("_seam_engines", {})
        key = (T, int(output_len), str(feats.device), self.precision)
        eng = engines.get(key)
        if eng is not None and B > eng.max_batch:
            eng.close()
            eng = None
        if eng is None:
            eng = Engine(fuser_kind=_capi.FUSER_NONE, T=T, mod_names=["feats"], mod_dims=[G], dim=G, fuser_depth=0,
                         fuser_heads=1, modal_encoding=False, frame_level_token=False, cross_attn=False,
                         norm_elementwise=True, gpt_dim=G, gpt_layers=gpt.n_layer, gpt_heads=gpt.n_head, cls_names=[],
                         cls_dims=[], precision=self.precision, max_batch=max(B, self.max_batch), device=feats.device,
                         fp_output_len=int(output_len), stages=_capi.STAGE_GPT)
            engines[key] = eng
        eng.sync_weights({"future_predictor." + n: p for n, p in self.named_parameters()})
        if self.output_attentions and output_len != 1:
            raise NotImplementedError("output_attentions with output_len > 1: the single-query roll-out kernel does not keep its probabilities")
        hidden, att = eng.forward_gpt(feats.to(torch.float32).contiguous(), want_attn=bool(self.output_attentions))
        # reference :403-409: {'gpt2_att_<output_id>': (B, n_layer, n_head, T, T)} when output_attentions is set
        return hidden, ({"gpt2_att_0": att} if att is not None else {})


# ------------------------------------------------------------------------------------------------
# CMFPEarly
# ------------------------------------------------------------------------------------------------
class CMFPEarly(nn.Module):
    """Early-fusion cross-modal future predictor - reference models/future_prediction.py:19-186,228-291.

    precision (include/afft_b200.h AFFT_PREC_*; accumulation / residual stream / LayerNorm / softmax are fp32 in all):
      "bf16"   bf16 tensor-core operands (the default).
      "fp16"   fp16 operands at the same tensor rate, 8x less operand rounding (saturating conversions).
      "strict" error-compensated bf16x3 GEMMs (hi.hi + hi.lo + lo.hi), fp32 activations between kernels; the mode
               whose top-5 indices are required to be identical to the fp32 reference (``strict=True`` selects it).
    """

    def __init__(self, model_cfg, num_classes, strict: bool = False, max_batch: int = 64, precision: str = None):
        super().__init__()
        logger = logging.getLogger(__name__)
        common = cfg_get(model_cfg, "common")
        modal_dims = cfg_get(model_cfg, "modal_dims")
        assert isinstance(modal_dims, Mapping), 'cfg.model.modal_dims must be a Dict!'
        for flag in ("share_classifiers", "share_predictors"):
            if not cfg_get(common, flag):
                logger.warning("Enforcing %s for early CMFP.", flag)
        if cfg_get(common, "modality_cls") or not cfg_get(common, "fusion_cls"):
            raise NotImplementedError("CMFPEarly here supports modality_cls=false, fusion_cls=true (every shipped config)")
        self.cfg = model_cfg
        self.num_classes = dict(num_classes)
        self.latent_dim = cfg_get(common, "in_features")
        self.fp_inter_dim = cfg_get(common, "fp_inter_dim")
        self.modality_dims = dict(modal_dims)
        self.fp_output_len = int(cfg_get(common, "fp_output_len", 1))
        if self.fp_output_len < 1:
            raise ValueError("fp_output_len must be >= 1")
        self.modal_feature_order = list(cfg_get(model_cfg, "modal_feature_order"))

        self.mapping = nn.ModuleDict()
        for mod, dim in self.modality_dims.items():  # reference :47-54
            self.mapping[mod] = instantiate(cfg_get(model_cfg, "mapping"), in_features=dim, out_features=self.latent_dim)
        self.fuser = instantiate(cfg_get(model_cfg, "fuser"))  # reference :74-76
        if self.latent_dim != self.fp_inter_dim:  # reference :245-255 (Identity when the widths match)
            self.dim_encoder = nn.Linear(self.latent_dim, self.fp_inter_dim, bias=False)
            self.dim_decoder = nn.Linear(self.fp_inter_dim, self.latent_dim, bias=False)
        else:
            self.dim_encoder, self.dim_decoder = nn.Identity(), nn.Identity()
        self.future_predictor = instantiate(cfg_get(model_cfg, "future_predictor"), in_features=self.fp_inter_dim,
                                            dimension_mapping=False)  # reference :84-87
        self.classifiers = nn.ModuleDict()  # reference :97-122 (shared classifier, fusion_cls only)
        dropout = cfg_get(model_cfg, "dropout")
        for cls_type, cls_dim in self.num_classes.items():
            self.classifiers[cls_type] = nn.ModuleDict(
                {'all-fused': nn.Sequential(nn.Dropout(dropout), nn.Linear(self.latent_dim, cls_dim))})

        self.precision = _capi.resolve_precision(strict, precision)
        self.max_batch = max_batch
        self.return_attentions = True
        # upper bound on the K splits of skinny GEMMs (small batches); 1 = off, which makes a clip's result bitwise
        # independent of the batch size it is computed in (see afft_set_max_ksplit)
        self.max_ksplit = 4
        self._engines: Dict[tuple, Engine] = {}

    @property
    def strict(self) -> bool:
        return self.precision == "strict"

    @strict.setter
    def strict(self, on: bool):
        self.precision = "strict" if on else "bf16"

    # ---- reference helpers kept for API parity ----
    @staticmethod
    def ordered_feature_list(x_d: Dict[str, Tensor], feats_order):
        return [x_d[m] for m in feats_order]

    # ---- native path ----
    def _fused_mapping(self, mod: str) -> bool:
        return bool(getattr(self.mapping[mod], "fused_in_forward", False))

    def _named_weights(self) -> Dict[str, Tensor]:
        # mappings that run ahead of afft_forward (feature_mapping.py ablation variants) keep their weights in Python
        skip = tuple(f"mapping.{m}." for m in self.mapping if not self._fused_mapping(m))
        return {n: p for n, p in self.named_parameters() if not n.startswith(skip)} if skip else \
            {n: p for n, p in self.named_parameters()}

    def _engine(self, feats_order, T: int, B: int, device: torch.device) -> Engine:
        key = (tuple(feats_order), T, str(device), self.precision, self.fp_output_len)
        eng = self._engines.get(key)
        if eng is not None and B > eng.max_batch:
            eng.close()
            eng = None
        if eng is None:
            f = self.fuser
            gpt = self.future_predictor.gpt_model
            eng = Engine(fuser_kind=f.afft_kind, T=T, mod_names=list(feats_order),
                         mod_dims=[self.modality_dims[m] if self._fused_mapping(m) else self.latent_dim
                                   for m in feats_order], dim=self.latent_dim,
                         fuser_depth=f.depth, fuser_heads=f.num_heads, modal_encoding=bool(f.modal_encoding),
                         frame_level_token=bool(f.frame_level_token), cross_attn=bool(f.cross_attn),
                         norm_elementwise=bool(f.norm_elementwise), gpt_dim=gpt.n_embd, gpt_layers=gpt.n_layer,
                         gpt_heads=gpt.n_head, cls_names=list(self.num_classes.keys()),
                         cls_dims=list(self.num_classes.values()), precision=self.precision,
                         max_batch=max(B, self.max_batch), device=device, fp_output_len=self.fp_output_len)
            self._engines[key] = eng
        if eng.max_ksplit != self.max_ksplit:
            eng.set_max_ksplit(self.max_ksplit)
        return eng

    def forward(self, feats: Dict[str, Tensor]) -> Dict[str, Dict[str, Tensor]]:
        if self.training:
            # training step (BASELINE config 5): differentiable path built from the same native kernels
            from ..train import forward_train
            if next(iter(feats.values())).device.type != "cuda":
                raise _capi.AfftError("afft_b200 runs on CUDA devices only (sm_100a); there is no CPU path")
            return forward_train(self, feats)
        feats_order = [mod for mod in self.modal_feature_order if mod in feats]  # reference :258
        if set(feats_order) != set(self.modality_dims):
            raise ValueError(f"features {sorted(feats)} do not match modal_dims {sorted(self.modality_dims)}")
        first = feats[feats_order[0]]
        shape = first.shape
        assert all(feats[m].shape[:2] == shape[:2] for m in feats_order), \
            'The shape of all inputs of the fusion module should be the same!'
        B, T = shape[0], shape[1]
        f = self.fuser
        if getattr(f, "frame_level_token", False) and f.temporal_sequence_length is not None:
            assert f.temporal_sequence_length == T, f"Temporal sequence length not valid {f.temporal_sequence_length} vs {T}"
        self.fuser.precision = self.future_predictor.precision = self.precision  # the inner seams follow the head's mode
        eng = self._engine(feats_order, T, B, first.device)
        eng.sync_weights(self._named_weights())
        xs = []
        for m in feats_order:
            x = feats[m].to(torch.float32).contiguous()
            if not self._fused_mapping(m):  # ablation mapping: library kernels ahead of the fused call
                self.mapping[m].precision = self.precision
                x = self.mapping[m](x).contiguous()
            xs.append(x)
        want_t = bool(getattr(self.future_predictor, "output_attentions", False))
        if want_t and self.fp_output_len != 1:
            raise NotImplementedError("fp_output_attentions with fp_output_len > 1: the single-query roll-out kernel does not keep its probabilities")
        res = eng.forward(xs, want_attn=self.return_attentions, want_gpt_attn=want_t)
        z, pf, logits, attn = res[:4]
        temporal = {"gpt2_att_0": res[4]} if want_t else {}  # reference future_prediction.py:403-409

        out = {  # reference prepare_output :155-182 (views into the native output buffers)
            'orig_past': {'all-fused': z},
            'future': {'all-fused': pf[:, T:]},
            'all-fused': {'all-fused': z[:, T - 1:]},
            'past_futures': {'all-fused': pf[:, :T]},
        }
        for k, (cls, c) in enumerate(self.num_classes.items()):  # reference apply_classifier :144-153
            out[f'{PAST_LOGITS_PREFIX}logits/{cls}'] = {'all-fused': logits[k][:, :T, :c]}
            out[f'logits/{cls}'] = {'all-fused': logits[k][:, T:, :c]}
        if attn is None:
            attn = torch.zeros(B)  # CA-Fuser's dummy attention (reference fusion.py:269)
        out['attentions'] = {'all-fused': {'modality_attns': attn, 'temporal_attns': temporal}}
        return out

    def last_launch_count(self) -> int:
        return max((e.launch_count() for e in self._engines.values()), default=0)


# ------------------------------------------------------------------------------------------------
# per-modality heads: IndividualFuturePrediction, CMFPScoreFusion
# ------------------------------------------------------------------------------------------------
class _UnimodalPrediction(nn.Module):
    """What reference IndividualFuturePrediction and CMFPScoreFusion share (models/future_prediction.py:57-95,97-122,
    203-214): per modality a bias-free dim_encoder / dim_decoder around a (shared or private) GPT-2 predictor and a
    classifier on the modality's own width.  One native handle per modality (AFFT_FUSER_NONE) runs
    dim_encoder -> GPT-2 -> dim_decoder -> prepare_output -> classifier."""

    @property
    def strict(self) -> bool:
        return self.precision == "strict"

    @strict.setter
    def strict(self, on: bool):
        self.precision = "strict" if on else "bf16"

    def _init_common(self, model_cfg, num_classes, strict, max_batch, precision=None):
        common = cfg_get(model_cfg, "common")
        modal_dims = cfg_get(model_cfg, "modal_dims")
        assert isinstance(modal_dims, Mapping), 'cfg.model.modal_dims must be a Dict!'
        self.cfg = model_cfg
        self.num_classes = dict(num_classes)
        self.latent_dim = cfg_get(common, "in_features")
        self.fp_inter_dim = cfg_get(common, "fp_inter_dim")
        self.modality_dims = dict(modal_dims)
        self.common_predictor = bool(cfg_get(common, "share_predictors"))
        self.common_classifier = bool(cfg_get(common, "share_classifiers"))
        self.fp_output_len = int(cfg_get(common, "fp_output_len", 1))
        self.modal_feature_order = list(cfg_get(model_cfg, "modal_feature_order"))
        self.precision, self.max_batch = _capi.resolve_precision(strict, precision), max_batch
        self._engines: Dict[tuple, Engine] = {}

    def _init_future_predictor(self, model_cfg):  # reference :78-95
        for d in self.modality_dims.values():
            if d == self.fp_inter_dim:
                raise NotImplementedError("modality width == fp_inter_dim (Identity dim_encoder) is not supported")
        self.dim_encoder = nn.ModuleDict({m: nn.Linear(d, self.fp_inter_dim, bias=False) for m, d in self.modality_dims.items()})
        self.dim_decoder = nn.ModuleDict({m: nn.Linear(self.fp_inter_dim, d, bias=False) for m, d in self.modality_dims.items()})
        fp_cfg = cfg_get(model_cfg, "future_predictor")
        if self.common_predictor:
            self.future_predictor = instantiate(fp_cfg, in_features=self.fp_inter_dim, dimension_mapping=False)
        else:
            self.future_predictor = nn.ModuleDict({m: instantiate(fp_cfg, in_features=self.fp_inter_dim, dimension_mapping=False)
                                                   for m in self.modality_dims})

    def _init_classifiers(self, model_cfg):  # reference :97-122 with modality_cls=true, fusion_cls=false
        dropout = cfg_get(model_cfg, "dropout")
        self.classifiers = nn.ModuleDict()
        for cls_type, cls_dim in self.num_classes.items():
            shared = nn.Sequential(nn.Dropout(dropout), nn.Linear(self.latent_dim, cls_dim)) if self.common_classifier else None
            self.classifiers[cls_type] = nn.ModuleDict(
                {m: shared if shared is not None else nn.Sequential(nn.Dropout(dropout), nn.Linear(d, cls_dim))
                 for m, d in self.modality_dims.items()})

    def _predictor(self, mod):
        return self.future_predictor if self.common_predictor else self.future_predictor[mod]

    def _weights_for(self, mod) -> Dict[str, Tensor]:
        """This modality's tensors under the names a fused-head handle expects (include/afft_b200.h)."""
        w = {"dim_encoder.weight": self.dim_encoder[mod].weight, "dim_decoder.weight": self.dim_decoder[mod].weight}
        for n, p in self._predictor(mod).named_parameters():
            w["future_predictor." + n] = p
        for cls in self.num_classes:
            lin = self.classifiers[cls][mod][1]
            w[f"classifiers.{cls}.all-fused.1.weight"] = lin.weight
            w[f"classifiers.{cls}.all-fused.1.bias"] = lin.bias
        return w

    def _run_modality(self, mod, x: Tensor):
        """-> (past_futures buffer (B, T+O, C_mod): slot 0 = x[:, 0], slots 1.. = predictions; [logits buffers (B, T+O, ld)])"""
        if self.training:
            raise NotImplementedError(f"{type(self).__name__}: the training-step path is built for CMFPEarly + SA-Fuser only")
        if x.device.type != "cuda":
            raise _capi.AfftError("afft_b200 runs on CUDA devices only (sm_100a); there is no CPU path")
        B, T, d = x.shape
        key = (mod, T, str(x.device), self.precision, self.fp_output_len)
        eng = self._engines.get(key)
        if eng is not None and B > eng.max_batch:
            eng.close()
            eng = None
        if eng is None:
            gpt = self._predictor(mod).gpt_model
            eng = Engine(fuser_kind=_capi.FUSER_NONE, T=T, mod_names=[mod], mod_dims=[d], dim=d, fuser_depth=0,
                         fuser_heads=1, modal_encoding=False, frame_level_token=False, cross_attn=False,
                         norm_elementwise=True, gpt_dim=gpt.n_embd, gpt_layers=gpt.n_layer, gpt_heads=gpt.n_head,
                         cls_names=list(self.num_classes.keys()), cls_dims=list(self.num_classes.values()),
                         precision=self.precision, max_batch=max(B, self.max_batch), device=x.device,
                         fp_output_len=self.fp_output_len)
            self._engines[key] = eng
        eng.sync_weights(self._weights_for(mod))
        _, pf, logits, _ = eng.forward([x.to(torch.float32).contiguous()], want_attn=False)
        return pf, logits

    def _unimodal_outputs(self, z: Dict[str, Tensor]):
        """prepare_output (:155-182) + apply_classifier (:144-153) per modality, as views into the native buffers."""
        out = {'orig_past': dict(z), 'future': {}, 'all-fused': {}, 'past_futures': {}}
        for cls in self.num_classes:
            out[f'{PAST_LOGITS_PREFIX}logits/{cls}'] = {}
            out[f'logits/{cls}'] = {}
        bufs = {}
        for mod, x in z.items():
            T = x.shape[1]
            pf, logits = self._run_modality(mod, x)
            bufs[mod] = (pf, logits)
            out['past_futures'][mod] = pf[:, :T]
            out['future'][mod] = pf[:, T:]
            for k, (cls, c) in enumerate(self.num_classes.items()):
                out[f'{PAST_LOGITS_PREFIX}logits/{cls}'][mod] = logits[k][:, :T, :c]
                out[f'logits/{cls}'][mod] = logits[k][:, T:, :c]
        return out, bufs

    def last_launch_count(self) -> int:
        return sum(e.launch_count() for e in self._engines.values())


class IndividualFuturePrediction(_UnimodalPrediction):
    """Individual modality future predictor - reference models/future_prediction.py:189-225 (expts/00_*)."""

    def __init__(self, model_cfg, num_classes, strict: bool = False, max_batch: int = 64, precision: str = None):
        super().__init__()
        assert not cfg_get(cfg_get(model_cfg, "common"), "fusion_cls")  # reference :194
        self._init_common(model_cfg, num_classes, strict, max_batch, precision)
        self._init_classifiers(model_cfg)   # registration order of the reference: classifiers, then predictors (:196-198)
        self._init_future_predictor(model_cfg)

    def forward(self, z: Dict[str, Tensor]) -> Dict[str, Dict[str, Tensor]]:
        out, _ = self._unimodal_outputs(z)
        return out


class CMFPScoreFusion(_UnimodalPrediction):
    """Late (score) fusion with MATT - reference models/future_prediction.py:294-351 (expts/05_MATT_ek100_train.txt)."""

    def __init__(self, model_cfg, num_classes, strict: bool = False, max_batch: int = 64, precision: str = None):
        super().__init__()
        common = cfg_get(model_cfg, "common")
        assert not cfg_get(common, "fusion_cls")  # reference :298
        if not cfg_get(common, "modality_cls"):
            logging.getLogger(__name__).warning("Enforcing modality classification for CMFPScoreFusion.")
        self._init_common(model_cfg, num_classes, strict, max_batch, precision)
        if self.fp_output_len != 1:
            raise NotImplementedError("CMFPScoreFusion broadcasts one attention row over the future logits: fp_output_len must be 1")
        self.mapping = nn.ModuleDict()  # reference :47-54
        for mod, dim in self.modality_dims.items():
            self.mapping[mod] = instantiate(cfg_get(model_cfg, "mapping"), in_features=dim, out_features=self.latent_dim)
        self.fuser = instantiate(cfg_get(model_cfg, "fuser"))
        self._init_future_predictor(model_cfg)
        self._init_classifiers(model_cfg)

    @staticmethod
    def ordered_feature_list(x_d: Dict[str, Tensor], feats_order):
        return [x_d[m] for m in feats_order]

    def forward(self, z: Dict[str, Tensor]) -> Dict[str, Dict[str, Tensor]]:
        feats_order = [mod for mod in self.modal_feature_order if mod in z]  # reference :308
        out, bufs = self._unimodal_outputs(z)
        first = z[feats_order[0]]
        B, T = first.shape[0], first.shape[1]
        # [first frame | predictions] (:327-330) is exactly the native past_futures buffer; map it to the common width
        mapped = {}
        for mod in feats_order:
            self.mapping[mod].precision = self.precision
            mapped[mod] = self.mapping[mod](bufs[mod][0])
        self.fuser.precision = self.precision
        scores = self.fuser.attn_logits(mapped, lambda d: [d[m] for m in feats_order])  # (B*(T+1), M), pre-softmax
        attn = torch.empty(scores.shape[0], len(feats_order), device=scores.device, dtype=torch.float32)
        for k, (cls, c) in enumerate(self.num_classes.items()):  # :341-350, softmax fused with the weighted sum
            per_mod = [bufs[m][1][k] for m in feats_order]  # (B, T+1, ld) each
            fused = torch.empty_like(per_mod[0])
            _capi.score_fusion(scores, [t.view(-1, t.shape[-1]) for t in per_mod], c, attn=attn,
                               out=fused.view(-1, fused.shape[-1]))
            out[f'{PAST_LOGITS_PREFIX}logits/{cls}'] = {'all-fused': fused[:, :T, :c]}
            out[f'logits/{cls}'] = {'all-fused': fused[:, T:, :c]}
        self.last_modality_attns = attn.view(B, T + 1, len(feats_order))  # not a reference output; kept for inspection
        return out
