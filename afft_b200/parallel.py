"""Thread-per-GPU inference replicas with persistent packed weights: a drop-in for the ``nn.DataParallel`` wrapper of the
reference's test.py:130.

``nn.DataParallel`` re-creates its replicas on every forward by broadcasting the parameters; a replica's parameter
tensors are therefore new objects each time and an engine on device k != 0 cannot tell that nothing changed - it would
re-pack all 772 MB of weights per forward.  This wrapper keeps one replica module per device alive (module deep-copied
to the device once), re-copies a parameter only when the SOURCE parameter's version counter moved, and leaves the
engines' packed bf16 / fp16 weights alone otherwise.  Same call signature and result as ``nn.DataParallel``:

    model = afft_b200.parallel.DataParallel(model, device_ids=range(cfg.num_gpus))     # test.py:130
    outputs, outputs_target = model(feature_dict, mixup_fn=None, target=None, ...)     # test.py:72-82

The batch is split contiguously over the devices (no data-path collective: clips are independent units), each shard
runs in its own thread on its own handle (the C ABI is re-entrant across handles), outputs are gathered on
``output_device``.
"""
from __future__ import annotations

import copy
import threading
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn as nn

from .dist import shard_bounds


def _gather(outs: List, device):
    """Concatenate the shards' outputs along the batch dimension (dicts of dicts of tensors; None stays None)."""
    first = outs[0]
    if first is None:
        return None
    if isinstance(first, torch.Tensor):
        if first.dim() == 0:
            return first.to(device)
        return torch.cat([o.to(device, non_blocking=True) for o in outs], dim=0)
    if isinstance(first, dict):
        return {k: _gather([o[k] for o in outs], device) for k in first}
    if isinstance(first, (list, tuple)):
        return type(first)(_gather([o[i] for o in outs], device) for i in range(len(first)))
    return first


class DataParallel(nn.Module):
    def __init__(self, module: nn.Module, device_ids: Optional[Sequence[int]] = None, output_device: Optional[int] = None):
        super().__init__()
        if not torch.cuda.is_available():
            raise RuntimeError("afft_b200.parallel.DataParallel needs CUDA devices (the hot path has no CPU fallback)")
        self.module = module
        self.device_ids = list(device_ids) if device_ids is not None else list(range(torch.cuda.device_count()))
        self.output_device = self.device_ids[0] if output_device is None else output_device
        self._replicas: Dict[int, nn.Module] = {}   # slot -> replica (slot 0 is the module itself)
        self._versions: Dict[int, tuple] = {}

    def _replica(self, slot: int) -> nn.Module:
        dev = torch.device("cuda", self.device_ids[slot])
        src_params = dict(self.module.named_parameters())
        src_bufs = dict(self.module.named_buffers())
        if slot == 0 and next(self.module.parameters()).device == dev:
            return self.module
        rep = self._replicas.get(slot)
        if rep is None:
            rep = copy.deepcopy(self.module).to(dev)
            for m in rep.modules():  # native handles are per device: the copy starts without any
                if isinstance(getattr(m, "_engines", None), dict):
                    m._engines = {}
                m.__dict__.pop("_seam_engines", None)
            self._replicas[slot] = rep
            self._versions[slot] = tuple((n, p.data_ptr(), p._version) for n, p in src_params.items())
            return rep
        versions = tuple((n, p.data_ptr(), p._version) for n, p in src_params.items())
        if versions != self._versions[slot]:  # load_state_dict / init_model / an optimizer step touched the source
            old = {t[0]: t for t in self._versions[slot]}
            with torch.no_grad():
                for (n, p), dst in zip(src_params.items(), rep.parameters()):
                    if old.get(n) != (n, p.data_ptr(), p._version):
                        dst.copy_(p)  # bumps the replica parameter's version -> its engine re-packs just this tensor
                for (n, b), dst in zip(src_bufs.items(), rep.buffers()):
                    dst.copy_(b)
            self._versions[slot] = versions
        rep.train(self.module.training)
        return rep

    def forward(self, video_data: Dict[str, torch.Tensor], *args, **kwargs):
        n_dev = len(self.device_ids)
        B = next(iter(video_data.values())).shape[0]
        shards = [shard_bounds(B, r, n_dev) for r in range(n_dev)]
        shards = [(r, lo, hi) for r, (lo, hi) in enumerate(shards) if hi > lo]
        results: List = [None] * len(shards)
        errors: List = [None] * len(shards)

        def sl(v, lo, hi):  # per-sample tensors are split, everything else is passed through
            return v[lo:hi] if isinstance(v, torch.Tensor) and v.dim() > 0 and v.shape[0] == B else v

        def run(i, slot, lo, hi):
            try:
                dev = torch.device("cuda", self.device_ids[slot])
                with torch.cuda.device(dev):
                    rep = self._replica(slot)
                    data = {m: t[lo:hi].to(dev, non_blocking=True) for m, t in video_data.items()}
                    kw = {k: (sl(v, lo, hi).to(dev) if isinstance(v, torch.Tensor) else v) for k, v in kwargs.items()}
                    results[i] = rep(data, *args, **kw)
                    torch.cuda.current_stream(dev).synchronize()
            except BaseException as exc:  # noqa: BLE001 - re-raised in the calling thread
                errors[i] = exc

        if len(shards) == 1:
            run(0, *shards[0])
        else:
            threads = [threading.Thread(target=run, args=(i, slot, lo, hi)) for i, (slot, lo, hi) in enumerate(shards)]
            for t in threads:
                t.start()
            for t in threads:
                t.join()
        for e in errors:
            if e is not None:
                raise e
        return _gather(results, torch.device("cuda", self.output_device))
