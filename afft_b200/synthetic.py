"""Deterministic synthetic inputs and weights (there are no checkpoints or LMDB features offline).

Both are pure functions of (name/shape, seed) through ``torch.Generator`` on the CPU, so the same
tensors can be regenerated on any box and loaded into this package's modules, the oracle and the
reference module alike (``load_state_dict``).
"""
from __future__ import annotations

import hashlib
from typing import Dict

import torch


def _seed_for(name: str, seed: int) -> int:
    h = hashlib.sha256(f"{seed}:{name}".encode()).digest()
    return int.from_bytes(h[:8], "little") & 0x7FFFFFFFFFFFFFFF


def synthetic_features(modal_dims: Dict[str, int], B: int, T: int, seed: int = 123, family: str = "randn",
                       six_d: bool = False) -> Dict[str, torch.Tensor]:
    """Per modality (B, T, C) fp32 (or the loader's (B, T, C, 1, 1, 1) when six_d).

    family "randn": standard normal, the reference authors' own smoke-test distribution (tmp.py:59,118).
    family "relu":  non-negative, closer to post-ReLU TSN features; 'objects' sparse uniform.
    """
    out = {}
    for mod, c in modal_dims.items():
        g = torch.Generator().manual_seed(_seed_for(f"feat.{mod}", seed))
        x = torch.randn(B, T, c, generator=g)
        if family == "relu":
            if mod == "objects":
                mask = torch.rand(B, T, c, generator=g) < 0.1
                x = torch.rand(B, T, c, generator=g) * mask
            else:
                x = x.abs()
        elif family != "randn":
            raise ValueError(family)
        out[mod] = x.reshape(B, T, c, 1, 1, 1) if six_d else x
    return out


def synthetic_state_dict(module: torch.nn.Module, seed: int = 0, std: float = 0.02) -> Dict[str, torch.Tensor]:
    """A full state dict for `module` (this package's BaseModel or the reference's): N(0, std) for
    matrices/tokens/embeddings, LayerNorm weights 1 + N(0, 0.1), all biases N(0, std) (non-zero so bias paths
    are exercised).  Keys GPT-2 registers as buffers (attn.bias / masked_bias) are left untouched."""
    sd = {}
    for name, p in module.named_parameters():
        g = torch.Generator().manual_seed(_seed_for(name, seed))
        leaf = name.rsplit(".", 2)
        is_ln = any(k in name for k in (".norm", "ln_1", "ln_2", "ln_f", "norm_self", "norm_q", "norm_kv", "norm_mlp"))
        if is_ln and name.endswith(".weight"):
            t = 1.0 + 0.1 * torch.randn(p.shape, generator=g)
        elif is_ln and name.endswith(".bias"):
            t = 0.05 * torch.randn(p.shape, generator=g)
        else:
            t = std * torch.randn(p.shape, generator=g)
        del leaf
        sd[name] = t.to(p.dtype)
    return sd
