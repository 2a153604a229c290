"""Host-side composition of the stateless C-ABI operators for the head / mapping variants outside the fused
``afft_forward`` call (GatedLinear / NonLinear / layer-normed mappings, MATT).

Every arithmetic step is a library kernel (afft_convert_bf16, afft_gemm, afft_layernorm); PyTorch only owns the
buffers.  Weights are packed to the precision's 16-bit operand format (bf16; hi/lo bf16 pairs in strict mode; fp16)
once per parameter version.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from . import _capi


class WeightCache:
    """bf16 (hi, lo) copies of nn.Linear weights [N, K], re-packed when the parameter changes."""

    def __init__(self):
        self._packed: Dict[int, Tuple[tuple, torch.Tensor, Optional[torch.Tensor]]] = {}

    def get(self, weight: torch.Tensor, precision: str):
        key = (weight.data_ptr(), weight._version, precision, str(weight.device))
        hit = self._packed.get(id(weight))
        if hit is not None and hit[0] == key:
            return hit[1], hit[2]
        w = weight.detach()
        if w.dtype != torch.float32 or not w.is_contiguous():
            w = w.float().contiguous()
        hi, lo = to_operand(w, precision)
        self._packed[id(weight)] = (key, hi, lo)
        return hi, lo


def to_operand(x: torch.Tensor, precision: str):
    """fp32 [R, K] -> the precision's GEMM operand through afft_convert_operand: bf16 hi (and lo = bf16(x - hi) in strict
    mode), or saturated fp16.  The result is [R, Kp] with Kp = K rounded up to a multiple of 8 (TMA needs 16-byte row
    pitches); the pad columns are zero and add nothing to a contraction (MATT with dim=None: 3424 -> 856 -> 428)."""
    if not x.is_cuda:
        raise _capi.AfftError("afft_b200 runs on CUDA devices only (sm_100a); there is no CPU path")
    R, K = x.shape
    Kp = (K + 7) // 8 * 8
    dt = _capi.operand_dtype(precision)
    hi = torch.empty(R, Kp, device=x.device, dtype=dt) if Kp == K else torch.zeros(R, Kp, device=x.device, dtype=dt)
    lo = (torch.empty_like(hi) if Kp == K else torch.zeros_like(hi)) if precision == "strict" else None
    _capi.check(_capi.lib().afft_convert_operand(x.data_ptr(), x.stride(0), R, K, hi.data_ptr(), _capi.ptr(lo), Kp, 0,
                                                 _capi.PRECISIONS[precision], _capi.current_stream_ptr(x.device)))
    return hi, lo


def dense(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], cache: WeightCache, *, precision: str,
          act: int = _capi.ACT_NONE, res: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """epilogue(x . weight^T + bias) for fp32 x [R, K] -> fp32 [R, N] (pitch padded to 4 floats)."""
    x = x.contiguous()
    R = x.shape[0]
    N = weight.shape[0]
    a_hi, a_lo = to_operand(x, precision)
    w_hi, w_lo = cache.get(weight, precision)
    if out is None:
        Np = (N + 3) // 4 * 4
        out = torch.empty(R, Np, device=x.device, dtype=torch.float32)[:, :N]
    if bias is not None:
        bias = bias.detach()
        if N % 4 != 0:  # the epilogue reads the bias with 16-byte loads
            padded = torch.zeros((N + 3) // 4 * 4, device=x.device, dtype=torch.float32)
            padded[:N] = bias
            bias = padded
    _capi.gemm(a_hi, w_hi, a_lo=a_lo, w_lo=w_lo, bias=bias, res=res, act=act, out_f32=out)
    return out


def layernorm(x: torch.Tensor, norm: torch.nn.LayerNorm) -> torch.Tensor:
    y = torch.empty_like(x)
    _capi.layernorm(x, norm.weight.detach() if norm.weight is not None else None,
                    norm.bias.detach() if norm.bias is not None else None, norm.eps, y_f32=y)
    return y
