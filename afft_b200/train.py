"""Training-step path (BASELINE config 5: SA-Fuser forward + backward, gradients all-reduced with NCCL by DDP).

The reference trains by running autograd through nn.Linear / nn.LayerNorm / nn.GELU / softmax attention
(train.py:234-262, common/runner.py:258-267).  Here every one of those library calls - forward AND backward - is
a hand-written sm_100a kernel behind the C ABI:

  LinearFn     forward  y = x W^T + b          afft_gemm (tcgen05, bf16 operands, fp32 accumulate)
               dgrad    dx = dy W              afft_gemm on the transposed weight
               wgrad    dW = dy^T x            afft_gemm on transposed activations (contraction over rows)
               dbias                           afft_colsum
  LayerNormFn  afft_layernorm / afft_layernorm_bwd
  GeluFn       afft_gelu_fwd / afft_gelu_bwd   (erf for the fuser MLP, tanh for GPT-2)
  AttentionFn  afft_attention (fp32 q|k|v, probabilities saved) / afft_attention_bwd

PyTorch supplies the autograd graph, residual adds, concatenation, dropout masks and the optimizer (plumbing);
torch.nn.parallel.DistributedDataParallel supplies the bucketed NCCL all-reduce overlapped with backward.
Scope of this first version: ModalTokenCMFuser (SA-Fuser) + GPT-2 + heads; activations between kernels are fp32;
attention-probability dropout is not applied (every other dropout / DropPath site is).
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

from . import _capi

_ST = _capi.current_stream_ptr


def _lib():
    return _capi.lib()


def _pad8(n: int) -> int:
    return (n + 7) // 8 * 8


def to_bf16(x: torch.Tensor) -> torch.Tensor:
    """fp32 [R, C] -> bf16 [R, C] (library conversion kernel); the row pitch is padded to a multiple of 8 elements
    (TMA needs 16-byte pitches) and the pad columns are zero."""
    x = x.contiguous()
    R, C = x.shape
    Cp = _pad8(C)
    buf = torch.zeros(R, Cp, device=x.device, dtype=torch.bfloat16) if Cp != C else \
        torch.empty(R, Cp, device=x.device, dtype=torch.bfloat16)
    _capi.check(_lib().afft_convert_bf16(x.data_ptr(), C, R, C, buf.data_ptr(), None, Cp, 0, _ST(x.device)))
    return buf[:, :C]


def to_bf16_t(x: torch.Tensor) -> torch.Tensor:
    """fp32 [R, C] -> bf16 [C, pad8(R)] (transposed; pad columns are zero so they add nothing to a contraction)."""
    x = x.contiguous()
    R, C = x.shape
    Rp = _pad8(R)
    out = torch.zeros(C, Rp, device=x.device, dtype=torch.bfloat16) if Rp != R else \
        torch.empty(C, Rp, device=x.device, dtype=torch.bfloat16)
    _capi.check(_lib().afft_convert_bf16(x.data_ptr(), C, R, C, out.data_ptr(), None, Rp, 1, _ST(x.device)))
    return out


def bf16_t(xb: torch.Tensor) -> torch.Tensor:
    """bf16 [R, C] -> bf16 [C, pad8(R)]."""
    R, C = xb.shape
    Rp = _pad8(R)
    out = torch.zeros(C, Rp, device=xb.device, dtype=torch.bfloat16) if Rp != R else \
        torch.empty(C, Rp, device=xb.device, dtype=torch.bfloat16)
    _capi.check(_lib().afft_transpose_bf16(xb.data_ptr(), xb.stride(0), R, C, out.data_ptr(), Rp, _ST(xb.device)))
    return out


class LinearFn(torch.autograd.Function):
    """y = x W^T + b for nn.Linear weights [N, K]; y = x W + b for transformers' Conv1D weights [K, N]."""

    @staticmethod
    def forward(ctx, x, weight, bias, conv1d: bool):
        xb = to_bf16(x)
        w_fwd = to_bf16_t(weight) if conv1d else to_bf16(weight)  # [N, K], K contiguous
        N = w_fwd.shape[0]
        Np = (N + 3) // 4 * 4  # fp32 output pitch: 16-byte multiple
        y = torch.empty(x.shape[0], Np, device=x.device, dtype=torch.float32)
        if bias is not None and Np != N:  # the epilogue reads the bias with 16-byte loads
            bias_p = torch.zeros(Np, device=x.device, dtype=torch.float32)
            bias_p[:N] = bias.detach()
            bias = bias_p[:N]
        _capi.gemm(xb, w_fwd[:, :xb.shape[1]], bias=bias, out_f32=y[:, :N])
        # the bf16 operand is kept for dgrad (its transpose is the dgrad operand): transposing 2-byte elements reads
        # half of what a second conversion of the fp32 weight would
        ctx.save_for_backward(xb, weight, w_fwd)
        ctx.conv1d, ctx.has_bias = conv1d, bias is not None
        return y[:, :N]

    @staticmethod
    def backward(ctx, dy):
        xb, weight, w_fwd = ctx.saved_tensors
        dy = dy.contiguous()
        M, N = dy.shape
        K = xb.shape[1]
        dx = dw = db = None
        dyb = to_bf16(dy)
        if ctx.needs_input_grad[0]:
            # dgrad: dx [M, K] = dy [M, N] . W;  B operand [K, N] with N contiguous
            w_dg = bf16_t(w_fwd[:, :K])  # [K, pad8(N)]: W for nn.Linear, W^T^T = W [K, N] for Conv1D
            dx = torch.empty(M, K, device=dy.device, dtype=torch.float32)
            _capi.gemm(dyb, w_dg[:, :N], out_f32=dx)
        if ctx.needs_input_grad[1]:
            dy_t = to_bf16_t(dy)  # [N, Mp]
            x_t = bf16_t(xb)      # [K, Mp]
            if ctx.conv1d:       # dW [K, N] = x^T . dy
                dw = torch.empty(K, N, device=dy.device, dtype=torch.float32)
                _capi.gemm(x_t, dy_t, out_f32=dw)
            else:                # dW [N, K] = dy^T . x
                dw = torch.empty(N, K, device=dy.device, dtype=torch.float32)
                _capi.gemm(dy_t, x_t, out_f32=dw)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = torch.zeros(N, device=dy.device, dtype=torch.float32)
            _capi.check(_lib().afft_colsum(dy.data_ptr(), N, M, N, db.data_ptr(), _ST(dy.device)))
        return dx, dw, db, None


class LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, eps: float):
        x = x.contiguous()
        y = torch.empty_like(x)
        _capi.layernorm(x, gamma, beta, eps, y_f32=y)
        ctx.save_for_backward(x, gamma)
        ctx.eps, ctx.affine = eps, gamma is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma = ctx.saved_tensors
        dy = dy.contiguous()
        rows, dim = x.shape
        dx = torch.empty_like(x)
        dg = db = None
        if ctx.affine:
            dg = torch.zeros(dim, device=x.device, dtype=torch.float32)
            db = torch.zeros(dim, device=x.device, dtype=torch.float32)
        _capi.check(_lib().afft_layernorm_bwd(x.data_ptr(), dim, _capi.ptr(gamma), ctx.eps, dy.data_ptr(), dim, rows, dim,
                                              dx.data_ptr(), dim, _capi.ptr(dg), _capi.ptr(db), _ST(x.device)))
        return dx, dg, db, None


class GeluFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, kind: int):
        x = x.contiguous()
        y = torch.empty_like(x)
        _capi.check(_lib().afft_gelu_fwd(x.data_ptr(), y.data_ptr(), x.numel(), kind, _ST(x.device)))
        ctx.save_for_backward(x)
        ctx.kind = kind
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        dy = dy.contiguous()
        dx = torch.empty_like(x)
        _capi.check(_lib().afft_gelu_bwd(x.data_ptr(), dy.data_ptr(), dx.data_ptr(), x.numel(), ctx.kind, _ST(x.device)))
        return dx, None


class AttentionFn(torch.autograd.Function):
    """Multi-head attention over short sequences on fp32 q|k|v rows [n_seq * L, 3 * H * hd]; mask as afft_attention."""

    @staticmethod
    def forward(ctx, qkv, n_seq: int, L: int, H: int, hd: int, mask: int, T: int):
        qkv = qkv.contiguous()
        D = H * hd
        hi = torch.empty(n_seq * L, D, device=qkv.device, dtype=torch.bfloat16)
        lo = torch.empty_like(hi)
        probs = torch.empty(n_seq, H, L, L, device=qkv.device, dtype=torch.float32)
        _capi.attention(qkv, n_seq, L, H, hd, mask=mask, T=T, out_hi=hi, out_lo=lo, probs=probs, p_outer=H * L * L)
        ctx.save_for_backward(qkv, probs)
        ctx.dims = (n_seq, L, H, hd)
        return hi.float() + lo.float(), probs

    @staticmethod
    def backward(ctx, d_out, _d_probs):
        qkv, probs = ctx.saved_tensors
        n_seq, L, H, hd = ctx.dims
        d_out = d_out.contiguous()
        dqkv = torch.empty_like(qkv)
        _capi.check(_lib().afft_attention_bwd(qkv.data_ptr(), qkv.shape[1], probs.data_ptr(), d_out.data_ptr(), d_out.shape[1],
                                              dqkv.data_ptr(), n_seq, L, H, hd, hd ** -0.5, _ST(qkv.device)))
        return dqkv, None, None, None, None, None, None


def _linear(x, lin, conv1d=False):
    return LinearFn.apply(x, lin.weight, getattr(lin, "bias", None), conv1d)


def _ln(x, norm):
    return LayerNormFn.apply(x, norm.weight, norm.bias, norm.eps)


def _drop_path(x, rate: float, training: bool, rows_per_sample: int):
    """Stochastic depth per sample (reference models/transformerblock.py:96-104); x rows are grouped per sample."""
    if rate == 0.0 or not training:
        return x
    keep = 1.0 - rate
    n = x.shape[0] // rows_per_sample
    mask = (keep + torch.rand(n, 1, 1, device=x.device, dtype=x.dtype)).floor_()
    return (x.view(n, rows_per_sample, -1) / keep * mask).view_as(x)


def forward_train(fp, feats: Dict[str, torch.Tensor]) -> Dict[str, Dict[str, torch.Tensor]]:
    """CMFPEarly.forward (reference models/future_prediction.py:257-291) in training mode, differentiable."""
    from .models.fusion import ModalTokenCMFuser
    fuser = fp.fuser
    if not isinstance(fuser, ModalTokenCMFuser):
        raise NotImplementedError("the training-step path is built for the SA-Fuser (ModalTokenCMFuser) only so far")
    if fuser.cross_attn:
        raise NotImplementedError("cross_attn=True is not supported in the training-step path")
    training = fp.training
    order = [m for m in fp.modal_feature_order if m in feats]
    first = feats[order[0]]
    B, T = first.shape[0], first.shape[1]
    D = fp.latent_dim
    toks = []
    for m in order:  # feature mapping (models/feature_mapping.py:59-63)
        x = feats[m].reshape(B * T, -1).float()
        lin = fp.mapping[m].mapping[0]
        toks.append(x if isinstance(lin, torch.nn.Identity) else _linear(x, lin))
    n = len(order) + 1
    if not fuser.frame_level_token:
        tok = fuser.modal_token.expand(B * T, -1, -1).reshape(B * T, D)
    else:
        tok = fuser.modal_token.expand(B, -1, -1).reshape(B * T, D)
    h = torch.stack([tok] + toks, dim=1)  # (B*T, n, D): models/fusion.py:338-349
    if fuser.modality_embedding is not None:
        h = h + fuser.modality_embedding
    h = F.dropout(h, fuser.embd_drop.p, training).reshape(B * T * n, D)
    H1 = fuser.num_heads
    attns = []
    for blk in fuser.blocks:  # models/transformerblock.py:131-135
        dp = getattr(blk.drop_path, "drop_prob", 0.0) or 0.0
        y = _ln(h, blk.norm1)
        a, p = AttentionFn.apply(_linear(y, blk.attn.qkv), B * T, n, H1, D // H1, 0, 1)
        a = F.dropout(_linear(a, blk.attn.proj), blk.attn.proj_drop.p, training)
        h = h + _drop_path(a, dp, training, n)
        y = _ln(h, blk.norm2)
        f = GeluFn.apply(_linear(y, blk.mlp.mlp[0]), _capi.ACT_GELU_ERF)
        f = F.dropout(_linear(f, blk.mlp.mlp[2]), blk.mlp.mlp[3].p, training)
        h = h + _drop_path(f, dp, training, n)
        attns.append(p.view(B, T, H1, n, n))
    x = _ln(h, fuser.norm)
    z = x.view(B * T, n, D)[:, 0].reshape(B, T, D)  # models/fusion.py:363-364

    gpt = fp.future_predictor.gpt_model
    G, H2 = gpt.n_embd, gpt.n_head
    g = _linear(z.reshape(B * T, D), fp.dim_encoder).view(B, T, G) + gpt.wpe.weight[:T]
    g = F.dropout(g, gpt.drop.p, training).reshape(B * T, G)
    for blk in gpt.h:  # transformers GPT2Block
        y = _ln(g, blk.ln_1)
        a, _ = AttentionFn.apply(_linear(y, blk.attn.c_attn, conv1d=True), B, T, H2, G // H2, 1, T)
        g = g + F.dropout(_linear(a, blk.attn.c_proj, conv1d=True), blk.attn.resid_dropout.p, training)
        y = _ln(g, blk.ln_2)
        f = GeluFn.apply(_linear(y, blk.mlp.c_fc, conv1d=True), _capi.ACT_GELU_TANH)
        g = g + F.dropout(_linear(f, blk.mlp.c_proj, conv1d=True), blk.mlp.dropout.p, training)
    g = _ln(g, gpt.ln_f)
    z_hat = _linear(g, fp.dim_decoder).view(B, T, D)

    past_futures = torch.cat([z[:, :1], z_hat[:, :T - 1]], dim=1)  # models/future_prediction.py:172-176
    future = z_hat[:, T - 1:]
    out = {"orig_past": {"all-fused": z}, "future": {"all-fused": future}, "all-fused": {"all-fused": z[:, T - 1:]},
           "past_futures": {"all-fused": past_futures}}
    for cls, c in fp.num_classes.items():
        head = fp.classifiers[cls]["all-fused"]
        for prefix, src in (("past_", past_futures), ("", future)):
            s = F.dropout(src.reshape(-1, D), head[0].p, training)
            out[f"{prefix}logits/{cls}"] = {"all-fused": _linear(s, head[1]).reshape(B, -1, c)}
    out["attentions"] = {"all-fused": {"modality_attns": torch.stack(attns).transpose(0, 1).detach(), "temporal_attns": {}}}
    return out


def reference_losses(outputs, target: torch.Tensor, target_subclips: torch.Tensor, cls: str = "action") -> Dict[str, torch.Tensor]:
    """The three losses of reference common/runner.py:112-168 (BasicLossAccuracy, hard labels, loss weights 1):
    CE on the future logits, CE on the past logits (ignore_index -1), MSE(past_futures[:, 1:], orig_past[:, 1:])."""
    logits = outputs[f"logits/{cls}"]["all-fused"]
    past = outputs[f"past_logits/{cls}"]["all-fused"]
    losses = {
        f"cls_{cls}": F.cross_entropy(logits.reshape(-1, logits.shape[-1]), target.reshape(-1), ignore_index=-1,
                                      reduction="none").mean(),
        f"past_cls_{cls}": F.cross_entropy(past.reshape(-1, past.shape[-1]), target_subclips.reshape(-1), ignore_index=-1,
                                           reduction="none").mean(),
        "past_reg": F.mse_loss(outputs["past_futures"]["all-fused"][:, 1:], outputs["orig_past"]["all-fused"][:, 1:]),
    }
    losses["total"] = losses[f"cls_{cls}"] + losses[f"past_cls_{cls}"] + losses["past_reg"]
    return losses
