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
  MlpFn        Linear -> GELU -> Linear as one node: the activation (and its backward) is produced directly as the next
               GEMM's bf16 operands by afft_convert_dual_gelu, never in fp32
  AttnProjFn   attention -> output projection as one node: the kernel's bf16 heads feed the projection GEMM directly

PyTorch supplies the autograd graph, residual adds, concatenation, dropout masks and the optimizer (plumbing);
torch.nn.parallel.DistributedDataParallel supplies the bucketed NCCL all-reduce overlapped with backward.
Scope: CMFPEarly with every fuser (SA / SA without token / T-SA / CA) and every feature mapping + GPT-2 + heads;
activations between kernels are fp32; every stochastic regulariser of the reference is applied (attention-probability
dropout inside the attention kernel).  N > 1: afft_b200.dist.GradBuckets all-reduces the gradients per layer group from
inside the backward pass.
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


def to_bf16_dual(x: torch.Tensor, colsum: torch.Tensor = None):
    """fp32 [R, C] -> (bf16 [R, C], bf16 [C, pad8(R)]) in one pass (+ colsum[c] += sum_r x[r, c]): both orientations a
    GEMM needs of an activation / gradient (pad columns are zero)."""
    x = x.contiguous()
    R, C = x.shape
    Cp, Rp = _pad8(C), _pad8(R)
    hi = torch.zeros(R, Cp, device=x.device, dtype=torch.bfloat16) if Cp != C else \
        torch.empty(R, Cp, device=x.device, dtype=torch.bfloat16)
    tr = torch.zeros(C, Rp, device=x.device, dtype=torch.bfloat16) if Rp != R else \
        torch.empty(C, Rp, device=x.device, dtype=torch.bfloat16)
    _capi.check(_lib().afft_convert_dual(x.data_ptr(), C, R, C, hi.data_ptr(), Cp, tr.data_ptr(), Rp, _capi.ptr(colsum),
                                         _ST(x.device)))
    return hi[:, :C], tr


def bf16_t(xb: torch.Tensor) -> torch.Tensor:
    """bf16 [R, C] -> bf16 [C, pad8(R)]."""
    R, C = xb.shape
    Rp = _pad8(R)
    out = torch.zeros(C, Rp, device=xb.device, dtype=torch.bfloat16) if Rp != R else \
        torch.empty(C, Rp, device=xb.device, dtype=torch.bfloat16)
    _capi.check(_lib().afft_transpose_bf16(xb.data_ptr(), xb.stride(0), R, C, out.data_ptr(), Rp, _ST(xb.device)))
    return out


# ------------------------------------------------------------------------------------------------
# Flat training state: parameters, gradients, momentum and the bf16 operand image in four flat buffers
# ------------------------------------------------------------------------------------------------
_STATE = None  # the active TrainState (consulted by the autograd Functions below) or None


class TrainState:
    """Owns the flat buffers of one model's training step (reference train.py:315-368 builds torch.optim.SGD and wraps
    the model in DistributedDataParallel; common/runner.py:258-267 runs forward / backward / step).

      flat_p    fp32 parameters; every ``p.data`` is a view (layout = backward-completion order, slices 16-byte aligned)
      buckets   afft_b200.dist.GradBuckets: fp32 gradients in the same layout, ``p.grad`` are views, one NCCL all-reduce
                per layer group issued from inside backward
      flat_m    momentum
      flat_w16  bf16 image of flat_p, written by the optimizer kernel: the GEMM operands of the next step (nn.Linear
                weights are used as they lie for the forward GEMM, Conv1D weights as they lie for dgrad; the other
                orientation is one 2-byte transpose per weight and step instead of an fp32 conversion plus a transpose)

    While a state is active (``with state:``), the native backward kernels write weight, bias and LayerNorm gradients
    straight into the ``.grad`` views (first write of a step overwrites, later ones accumulate) instead of returning
    tensors for autograd to add, and ``step()`` is ONE kernel (afft_sgd_nesterov) over the flat buffers."""

    def __init__(self, head, lr: float, momentum: float = 0.9, weight_decay: float = 0.0, nesterov: bool = True,
                 comm_dtype=torch.float32, n_buckets: int = 0):
        from . import dist as adist
        self.lr, self.momentum, self.weight_decay, self.nesterov = lr, momentum, weight_decay, nesterov
        groups = merge_groups(grad_groups(head), n_buckets)
        self.buckets = adist.GradBuckets(groups, comm_dtype=comm_dtype)
        flat_g = self.buckets.flat
        self.flat_p = torch.empty_like(flat_g)
        self.flat_p.zero_()
        self.flat_m = torch.zeros_like(flat_g)
        self.flat_w16 = torch.empty(flat_g.numel(), device=flat_g.device, dtype=torch.bfloat16)
        self.w16 = {}
        self._params = {}
        with torch.no_grad():
            for g in self.buckets.groups:
                for p in g:
                    off, n = self.buckets.offsets[id(p)]
                    view = self.flat_p[off:off + n].view_as(p)
                    view.copy_(p.data)
                    p.data = view
                    self.w16[id(p)] = self.flat_w16[off:off + n].view_as(p)
                    self._params[id(p)] = p
        self.refresh_w16()
        self._uses = {}        # id(param) -> forward uses not yet met by a backward this step
        self._written = set()  # params whose .grad view has been written by a native kernel this step
        self._overwritten = set()   # params whose gradient a wgrad GEMM OVERWRITES on its first use (learned in step 1)
        self._accum_grads = None    # .grad views that are accumulated into (+=) and therefore cleared at step start

    def refresh_w16(self):
        """bf16 image of the current parameters (after construction / load_state_dict; the optimizer keeps it current)."""
        n = self.flat_p.numel()
        _capi.check(_lib().afft_convert_bf16(self.flat_p.data_ptr(), n, 1, n, self.flat_w16.data_ptr(), None, n, 0,
                                             _ST(self.flat_p.device)))

    def __enter__(self):
        global _STATE
        self._prev, _STATE = _STATE, self
        return self

    def __exit__(self, *exc):
        global _STATE
        _STATE = self._prev
        return False

    # ---- per-step protocol: zero() -> forward -> backward -> finish() -> step() ----
    def zero(self):
        """Start of a step.  GEMM weight gradients are overwritten by their first wgrad of the step, so from the second
        step on only the accumulated-into gradients (biases, LayerNorm, embeddings, tokens) are cleared - 1.5 GB of
        memset saved per step.  Which parameters are overwritten is learned during the first step."""
        if self._accum_grads is None:
            self.buckets.flat.zero_()
        elif self._accum_grads:
            torch._foreach_zero_(self._accum_grads)
        self.buckets._pending = [len(g) for g in self.buckets.groups]
        self.buckets._works = []
        self._uses.clear()
        self._written.clear()

    def note_use(self, p):
        self._uses[id(p)] = self._uses.get(id(p), 0) + 1

    def grad_target(self, p, overwrites: bool = False):
        """(.grad view, accumulate?) for a native kernel about to write this parameter's gradient.  overwrites: the
        kernel replaces the view's contents on the first write of a step (wgrad GEMM) instead of adding to them."""
        acc = id(p) in self._written
        self._written.add(id(p))
        if overwrites and self._accum_grads is None:
            self._overwritten.add(id(p))
        return p.grad, acc

    def grad_done(self, p):
        """One backward use of p finished; when it was the last one of the step, its bucket counts it."""
        left = self._uses.get(id(p), 1) - 1
        self._uses[id(p)] = left
        if left == 0:
            self.buckets.notify(p)

    def finish(self):
        """After backward: GEMM weights no kernel wrote this step (unused in the graph) are cleared, outstanding
        all-reduces are waited for."""
        if self._accum_grads is None:  # end of the first step: everything not overwritten by a GEMM is cleared per step
            self._accum_grads = [p.grad for pid, p in self._params.items() if pid not in self._overwritten]
        for pid in self._overwritten:
            if pid not in self._written:
                self._params[pid].grad.zero_()
        self.buckets.finish()

    def step(self):
        n = self.flat_p.numel()
        _capi.check(_lib().afft_sgd_nesterov(self.flat_p.data_ptr(), self.buckets.flat.data_ptr(), self.flat_m.data_ptr(),
                                             self.flat_w16.data_ptr(), n, self.lr, self.momentum, self.weight_decay,
                                             int(self.nesterov), _ST(self.flat_p.device)))


def _state_for(*params):
    st = _STATE
    if st is None:
        return None
    return st if all(p is None or id(p) in st.w16 for p in params) else None


def gelu_bf16_dual(u: torch.Tensor, kind: int, d_act: torch.Tensor = None, colsum: torch.Tensor = None, want_t: bool = True):
    """fp32 pre-activation u [R, C] -> bf16 gelu(u) (d_act None) or bf16 d_act * gelu'(u), as [R, C] and [C, pad8(R)], plus
    the column sums of the same values: afft_convert_dual_gelu - the activation never exists in fp32."""
    R, C = u.shape
    Cp, Rp = _pad8(C), _pad8(R)
    hi = torch.zeros(R, Cp, device=u.device, dtype=torch.bfloat16) if Cp != C else \
        torch.empty(R, Cp, device=u.device, dtype=torch.bfloat16)
    tr = None
    if want_t:
        tr = torch.zeros(C, Rp, device=u.device, dtype=torch.bfloat16) if Rp != R else \
            torch.empty(C, Rp, device=u.device, dtype=torch.bfloat16)
    _capi.check(_lib().afft_convert_dual_gelu(u.data_ptr(), u.stride(0), R, C, hi.data_ptr(), Cp, _capi.ptr(tr), Rp,
                                              _capi.ptr(colsum), _capi.ptr(d_act), d_act.stride(0) if d_act is not None else 0,
                                              kind, _ST(u.device)))
    return hi[:, :C], tr


# ---- the pieces of a Linear's forward / backward, shared by LinearFn and the fused nodes below ----
def _weight_operand(weight, bias, conv1d: bool, st):
    """bf16 [N, K] (K contiguous) forward operand of an nn.Linear [N, K] / Conv1D [K, N] weight."""
    if st is not None:  # the optimizer's bf16 image: no conversion
        w16 = st.w16[id(weight)]
        st.note_use(weight)
        if bias is not None:
            st.note_use(bias)
        return bf16_t(w16) if conv1d else w16
    return to_bf16_t(weight) if conv1d else to_bf16(weight)


def _gemm_bias(xb, w_fwd, bias):
    """fp32 y [M, N] = xb . w_fwd^T + bias (output pitch padded to 16 bytes)."""
    N = w_fwd.shape[0]
    Np = (N + 3) // 4 * 4
    y = torch.empty(xb.shape[0], Np, device=xb.device, dtype=torch.float32)
    if bias is not None and Np != N:  # the epilogue reads the bias with 16-byte loads
        bias_p = torch.zeros(Np, device=xb.device, dtype=torch.float32)
        bias_p[:N] = bias.detach()
        bias = bias_p[:N]
    _capi.gemm(xb, w_fwd[:, :xb.shape[1]], bias=bias, out_f32=y[:, :N])
    return y[:, :N]


def _bias_grad_target(bias, st, N: int, device, needed: bool):
    """(buffer the colsum kernel accumulates into, tensor to hand back to autograd or None)."""
    if bias is None or not needed:
        return None, None
    if st is not None:
        return st.grad_target(bias)[0], None  # cleared at step start; the kernel accumulates
    t = torch.zeros(N, device=device, dtype=torch.float32)
    return t, t


def _dgrad(dyb, weight, w_fwd, K: int, conv1d: bool, st):
    """dx [M, K] = dy [M, N] . W;  B operand [K, N] with N contiguous."""
    M, N = dyb.shape
    if st is not None and conv1d and N % 8 == 0:
        w_dg = st.w16[id(weight)]  # Conv1D weights are stored [K, N]: the optimizer's bf16 image is the operand
    else:
        w_dg = bf16_t(w_fwd[:, :K])  # [K, pad8(N)]: W for nn.Linear, W^T^T = W [K, N] for Conv1D
    dx = torch.empty(M, K, device=dyb.device, dtype=torch.float32)
    _capi.gemm(dyb, w_dg[:, :N], out_f32=dx)
    return dx


def _wgrad(dy_t, x_t, weight, conv1d: bool, st):
    """dW from the transposed bf16 operands dy^T [N, Mp] and x^T [K, Mp].  With a TrainState the GEMM writes the .grad view
    (no temporary, no autograd accumulation pass) and None is returned; otherwise the fp32 gradient is returned."""
    direct = st is not None and weight.shape[1] % 4 == 0
    if direct:
        dw_out, acc = st.grad_target(weight, overwrites=True)
    else:
        dw_out, acc = torch.empty(weight.shape, device=dy_t.device, dtype=torch.float32), False
    if conv1d:       # dW [K, N] = x^T . dy
        _capi.gemm(x_t, dy_t, out_f32=dw_out, res=dw_out if acc else None)
    else:            # dW [N, K] = dy^T . x
        _capi.gemm(dy_t, x_t, out_f32=dw_out, res=dw_out if acc else None)
    if direct:
        st.grad_done(weight)
        return None
    return dw_out  # autograd accumulates it and the bucket's hook counts the parameter


class LinearFn(torch.autograd.Function):
    """y = x W^T + b for nn.Linear weights [N, K]; y = x W + b for transformers' Conv1D weights [K, N]."""

    @staticmethod
    def forward(ctx, x, weight, bias, conv1d: bool):
        need_wgrad = weight.requires_grad and torch.is_grad_enabled()
        if need_wgrad:
            xb, x_t = to_bf16_dual(x)  # the wgrad operand (x^T) comes out of the same pass over x
        else:
            xb, x_t = to_bf16(x), None
        st = _state_for(weight, bias)
        ctx.st = st
        w_fwd = _weight_operand(weight, bias, conv1d, st)
        y = _gemm_bias(xb, w_fwd, bias)
        # the bf16 operand is kept for dgrad (its transpose is the dgrad operand): transposing 2-byte elements reads
        # half of what a second conversion of the fp32 weight would
        ctx.save_for_backward(x_t if x_t is not None else xb, weight, w_fwd, bias)
        ctx.x_is_t, ctx.K = x_t is not None, xb.shape[1]
        ctx.conv1d, ctx.has_bias = conv1d, bias is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        xs, weight, w_fwd, bias = ctx.saved_tensors
        st = ctx.st
        dy = dy.contiguous()
        M, N = dy.shape
        dx = dw = None
        # dy in both orientations (dgrad contracts over N, wgrad over the rows) and the bias gradient: one pass over dy
        db_out, db = _bias_grad_target(bias, st, N, dy.device, ctx.has_bias and ctx.needs_input_grad[2])
        if ctx.needs_input_grad[1]:
            dyb, dy_t = to_bf16_dual(dy, db_out)
        else:
            dyb, dy_t = to_bf16(dy), None
            if db_out is not None:
                _capi.check(_lib().afft_colsum(dy.data_ptr(), N, M, N, db_out.data_ptr(), _ST(dy.device)))
        if db_out is not None and st is not None:
            st.grad_done(bias)
        if ctx.needs_input_grad[0]:
            dx = _dgrad(dyb, weight, w_fwd, ctx.K, ctx.conv1d, st)
        if ctx.needs_input_grad[1]:
            x_t = xs if ctx.x_is_t else bf16_t(xs)  # [K, Mp]
            dw = _wgrad(dy_t, x_t, weight, ctx.conv1d, st)
        return dx, dw, db, None


class MlpFn(torch.autograd.Function):
    """Linear -> GELU -> Linear (reference models/transformerblock.py:39-56 Mlp; GPT-2's c_fc / act / c_proj) as one
    node: the activation is produced directly as the second GEMM's bf16 operands (afft_convert_dual_gelu) and its backward
    directly as the first Linear's backward operands + bias gradient, so neither gelu(u) nor d/du exists in fp32 - two
    fp32 [rows, 4D] round trips less per block and direction than LinearFn / GeluFn / LinearFn."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, kind: int, conv1d: bool):
        st = _state_for(w1, b1, w2, b2)
        ctx.st = st
        xb, x_t = to_bf16_dual(x)
        w1f = _weight_operand(w1, b1, conv1d, st)
        w2f = _weight_operand(w2, b2, conv1d, st)
        u = _gemm_bias(xb, w1f, b1)
        ab, a_t = gelu_bf16_dual(u, kind)
        y = _gemm_bias(ab, w2f, b2)
        ctx.save_for_backward(x_t, u, a_t, w1, w1f, b1, w2, w2f, b2)
        ctx.kind, ctx.conv1d, ctx.K = kind, conv1d, xb.shape[1]
        return y

    @staticmethod
    def backward(ctx, dy):
        x_t, u, a_t, w1, w1f, b1, w2, w2f, b2 = ctx.saved_tensors
        st, conv1d = ctx.st, ctx.conv1d
        dy = dy.contiguous()
        Fdim = u.shape[1]
        db2_out, db2 = _bias_grad_target(b2, st, dy.shape[1], dy.device, b2 is not None)
        dyb, dy_t = to_bf16_dual(dy, db2_out)
        if db2_out is not None and st is not None:
            st.grad_done(b2)
        dw2 = _wgrad(dy_t, a_t, w2, conv1d, st)
        da = _dgrad(dyb, w2, w2f, Fdim, conv1d, st)  # gradient w.r.t. gelu(u), fp32 [rows, F]
        db1_out, db1 = _bias_grad_target(b1, st, Fdim, dy.device, b1 is not None)
        dub, du_t = gelu_bf16_dual(u, ctx.kind, d_act=da, colsum=db1_out)
        if db1_out is not None and st is not None:
            st.grad_done(b1)
        dw1 = _wgrad(du_t, x_t, w1, conv1d, st)
        dx = _dgrad(dub, w1, w1f, ctx.K, conv1d, st) if ctx.needs_input_grad[0] else None
        return dx, dw1, db1, dw2, db2, None, None


class LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, eps: float):
        x = x.contiguous()
        y = torch.empty_like(x)
        _capi.layernorm(x, gamma, beta, eps, y_f32=y)
        ctx.save_for_backward(x, gamma, beta)
        ctx.eps, ctx.affine = eps, gamma is not None
        ctx.st = _state_for(gamma, beta) if gamma is not None else None
        if ctx.st is not None:
            ctx.st.note_use(gamma)
            ctx.st.note_use(beta)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma, beta = ctx.saved_tensors
        st = ctx.st
        dy = dy.contiguous()
        rows, dim = x.shape
        dx = torch.empty_like(x)
        dg = db = None
        if st is not None:  # the kernel accumulates (+=) into the .grad views (cleared at step start)
            dg_out, db_out = st.grad_target(gamma)[0], st.grad_target(beta)[0]
        elif ctx.affine:
            dg_out = dg = torch.zeros(dim, device=x.device, dtype=torch.float32)
            db_out = db = torch.zeros(dim, device=x.device, dtype=torch.float32)
        else:
            dg_out = db_out = None
        _capi.check(_lib().afft_layernorm_bwd(x.data_ptr(), dim, _capi.ptr(gamma), ctx.eps, dy.data_ptr(), dim, rows, dim,
                                              dx.data_ptr(), dim, _capi.ptr(dg_out), _capi.ptr(db_out), _ST(x.device)))
        if st is not None:
            st.grad_done(gamma)
            st.grad_done(beta)
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
    """Multi-head attention over short sequences on fp32 q|k|v rows [n_seq * L, 3 * H * hd]; mask as afft_attention.
    ``p_drop`` > 0 (training): attention-probability dropout (reference models/transformerblock.py:31,71; GPT-2
    attn_pdrop) - the keep / (1 - p) factors are drawn here and applied to the softmax inside the kernel before P.V.
    Returns (merged heads [n_seq * L, H * hd] fp32, probabilities AFTER dropout [n_seq, H, L, L], as the reference)."""

    @staticmethod
    def forward(ctx, qkv, n_seq: int, L: int, H: int, hd: int, mask: int, T: int, p_drop: float = 0.0):
        qkv = qkv.contiguous()
        D = H * hd
        hi = torch.empty(n_seq * L, D, device=qkv.device, dtype=torch.bfloat16)
        lo = torch.empty_like(hi)
        probs = torch.empty(n_seq, H, L, L, device=qkv.device, dtype=torch.float32)
        drop = None
        if p_drop > 0.0:
            drop = (torch.rand(n_seq, H, L, L, device=qkv.device) >= p_drop).to(torch.float32).mul_(1.0 / (1.0 - p_drop))
        _capi.attention(qkv, n_seq, L, H, hd, mask=mask, T=T, out_hi=hi, out_lo=lo, probs=probs, p_outer=H * L * L,
                        drop_mask=drop)
        ctx.save_for_backward(qkv, probs, drop)
        ctx.dims = (n_seq, L, H, hd)
        return hi.float() + lo.float(), (probs if drop is None else probs * drop)

    @staticmethod
    def backward(ctx, d_out, _d_probs):
        qkv, probs, drop = ctx.saved_tensors
        n_seq, L, H, hd = ctx.dims
        d_out = d_out.contiguous()
        dqkv = torch.empty_like(qkv)
        _capi.check(_lib().afft_attention_bwd(qkv.data_ptr(), qkv.shape[1], probs.data_ptr(), d_out.data_ptr(), d_out.shape[1],
                                              dqkv.data_ptr(), n_seq, L, H, hd, hd ** -0.5, _capi.ptr(drop), _ST(qkv.device)))
        return dqkv, None, None, None, None, None, None, None


class AttnProjFn(torch.autograd.Function):
    """AttentionFn followed by the output projection (reference models/transformerblock.py:30-35; GPT-2 c_proj) as one
    node: the kernel's bf16 merged heads are the projection GEMM's operand as they are (no fp32 copy of the attention output,
    no conversion back), and the projection's dgrad output is the attention backward's input."""

    @staticmethod
    def forward(ctx, qkv, weight, bias, conv1d: bool, n_seq: int, L: int, H: int, hd: int, mask: int, T: int, p_drop: float):
        qkv = qkv.contiguous()
        D = H * hd
        st = _state_for(weight, bias)
        ctx.st = st
        hi = torch.empty(n_seq * L, D, device=qkv.device, dtype=torch.bfloat16)
        probs = torch.empty(n_seq, H, L, L, device=qkv.device, dtype=torch.float32)
        drop = None
        if p_drop > 0.0:
            drop = (torch.rand(n_seq, H, L, L, device=qkv.device) >= p_drop).to(torch.float32).mul_(1.0 / (1.0 - p_drop))
        _capi.attention(qkv, n_seq, L, H, hd, mask=mask, T=T, out_hi=hi, probs=probs, p_outer=H * L * L, drop_mask=drop)
        w_fwd = _weight_operand(weight, bias, conv1d, st)
        y = _gemm_bias(hi, w_fwd, bias)
        ctx.save_for_backward(qkv, probs, drop, bf16_t(hi), weight, w_fwd, bias)
        ctx.dims, ctx.conv1d = (n_seq, L, H, hd), conv1d
        return y, (probs if drop is None else probs * drop)

    @staticmethod
    def backward(ctx, dy, _d_probs):
        qkv, probs, drop, x_t, weight, w_fwd, bias = ctx.saved_tensors
        st = ctx.st
        n_seq, L, H, hd = ctx.dims
        dy = dy.contiguous()
        db_out, db = _bias_grad_target(bias, st, dy.shape[1], dy.device, bias is not None)
        dyb, dy_t = to_bf16_dual(dy, db_out)
        if db_out is not None and st is not None:
            st.grad_done(bias)
        dw = _wgrad(dy_t, x_t, weight, ctx.conv1d, st)
        d_attn = _dgrad(dyb, weight, w_fwd, H * hd, ctx.conv1d, st)
        dqkv = torch.empty_like(qkv)
        _capi.check(_lib().afft_attention_bwd(qkv.data_ptr(), qkv.shape[1], probs.data_ptr(), d_attn.data_ptr(), d_attn.shape[1],
                                              dqkv.data_ptr(), n_seq, L, H, hd, hd ** -0.5, _capi.ptr(drop), _ST(qkv.device)))
        return dqkv, dw, db, None, None, None, None, None, None, None, None


def _fusable(*params) -> bool:
    return torch.is_grad_enabled() and all(p is None or p.requires_grad for p in params)


def _mlp(x, fc1, fc2, kind: int, conv1d: bool = False):
    """fc2(gelu(fc1(x))): one fused node when every parameter trains, the three separate ones otherwise."""
    b1, b2 = getattr(fc1, "bias", None), getattr(fc2, "bias", None)
    if _fusable(fc1.weight, b1, fc2.weight, b2):
        return MlpFn.apply(x, fc1.weight, b1, fc2.weight, b2, kind, conv1d)
    return _linear(GeluFn.apply(_linear(x, fc1, conv1d), kind), fc2, conv1d)


def _attn_proj(qkv, proj, n_seq, L, H, hd, mask, T, p_drop, conv1d: bool = False):
    """proj(attention(qkv)) -> (projection output, probabilities after dropout)."""
    bias = getattr(proj, "bias", None)
    if _fusable(proj.weight, bias):
        return AttnProjFn.apply(qkv, proj.weight, bias, conv1d, n_seq, L, H, hd, mask, T, p_drop)
    a, p = AttentionFn.apply(qkv, n_seq, L, H, hd, mask, T, p_drop)
    return _linear(a, proj, conv1d), p


def _linear(x, lin, conv1d=False):
    return LinearFn.apply(x, lin.weight, getattr(lin, "bias", None), conv1d)


def _ln(x, norm):
    return LayerNormFn.apply(x, norm.weight, norm.bias, norm.eps)


def _residual(h, f, p_drop: float, dp_rate: float, training: bool, rows_per_sample: int):
    """h + drop_path(dropout(f)) (reference models/transformerblock.py:131-135 with :96-104): the dropout is one fused
    kernel, stochastic depth and the residual add are one addcmul with a per-sample scale."""
    f = F.dropout(f, p_drop, training)
    if dp_rate == 0.0 or not training:
        return h + f
    keep = 1.0 - dp_rate
    n = f.shape[0] // rows_per_sample
    scale = (keep + torch.rand(n, 1, 1, device=f.device, dtype=f.dtype)).floor_().div_(keep)
    return torch.addcmul(h.view(n, rows_per_sample, -1), f.view(n, rows_per_sample, -1), scale).view_as(h)


def _drop_path(x, rate: float, training: bool, rows_per_sample: int):
    """Stochastic depth per sample (reference models/transformerblock.py:96-104); x rows are grouped per sample."""
    if rate == 0.0 or not training:
        return x
    keep = 1.0 - rate
    n = x.shape[0] // rows_per_sample
    mask = (keep + torch.rand(n, 1, 1, device=x.device, dtype=x.dtype)).floor_()
    return (x.view(n, rows_per_sample, -1) / keep * mask).view_as(x)


def _rate(module, attr="p"):
    return float(getattr(module, attr, 0.0) or 0.0)


def _apply_mapping(mp, x, training: bool):
    """reference models/feature_mapping.py: Linear (:54-78), GatedLinear (:36-51), NonLinear (:91-107) on [rows, C_m]."""
    from .models import feature_mapping as fm
    if isinstance(mp, fm.Linear):
        lin = mp.mapping[0]
        y = x if isinstance(lin, torch.nn.Identity) else _linear(x, lin)
    elif isinstance(mp, fm.GatedLinear):
        u = _linear(x, mp.mapping[0])
        y = u * torch.sigmoid(_linear(u, mp.mapping[1].fc))  # ContextGating: cat + glu (:21-33)
    elif isinstance(mp, fm.NonLinear):
        u = _linear(x, mp.mapping[0])
        y = torch.relu(u) if mp.activation == "relu" else (GeluFn.apply(u, _capi.ACT_GELU_ERF) if mp.activation == "gelu" else u)
    else:
        raise NotImplementedError(f"training step: unknown mapping {type(mp).__name__}")
    if mp.use_layernorm:
        y = _ln(y, mp.mapping[-1])
    return y


def _self_attention(h, attn, n_seq, L, mask, T, training, proj_drop=True):
    """reference models/transformerblock.py:19-36 on rows [n_seq * L, D]; returns (proj output [after proj_drop], probs)."""
    H = attn.num_heads
    D = h.shape[1]
    a, p = _attn_proj(_linear(h, attn.qkv), attn.proj, n_seq, L, H, D // H, mask, T,
                      _rate(attn.attn_drop) if training else 0.0)
    return (F.dropout(a, _rate(attn.proj_drop), training) if proj_drop else a), p


def _block(h, blk, n_seq, L, mask, T, training):
    """reference models/transformerblock.py:131-135 (Block); DropPath per sequence of L rows."""
    dp = getattr(blk.drop_path, "drop_prob", 0.0) or 0.0
    a, p = _self_attention(_ln(h, blk.norm1), blk.attn, n_seq, L, mask, T, training, proj_drop=False)
    h = _residual(h, a, _rate(blk.attn.proj_drop), dp, training, L)
    f = _mlp(_ln(h, blk.norm2), blk.mlp.mlp[0], blk.mlp.mlp[2], _capi.ACT_GELU_ERF)
    return _residual(h, f, _rate(blk.mlp.mlp[3]), dp, training, L), p


def _fuse_sa(fuser, toks, B, T, D, training, with_token: bool):
    """ModalTokenCMFuser (reference models/fusion.py:319-365) / CMFuser (:86-118): sequences of n tokens per (b, t)."""
    if with_token:
        tok = (fuser.modal_token.expand(B * T, -1, -1) if not fuser.frame_level_token
               else fuser.modal_token.expand(B, -1, -1)).reshape(B * T, D)
        toks = [tok] + toks
    n = len(toks)
    h = torch.stack(toks, dim=1)  # (B*T, n, D)
    if getattr(fuser, "modality_embedding", None) is not None:
        h = h + fuser.modality_embedding
    h = F.dropout(h, _rate(fuser.embd_drop), training).reshape(B * T * n, D)
    mask = 3 if fuser.cross_attn else 0
    H = fuser.num_heads
    attns = []
    for blk in fuser.blocks:
        h, p = _block(h, blk, B * T, n, mask, 1, training)
        attns.append(p.view(B, T, H, n, n))
    x = _ln(h, fuser.norm).view(B * T, n, D)
    z = x[:, 0] if with_token else x.mean(dim=1)
    return z.reshape(B, T, D), torch.stack(attns).transpose(0, 1).detach()


def _fuse_tsa(fuser, toks, B, T, D, training):
    """TemporalCMFuser (reference models/fusion.py:159-215): one sequence of n*T tokens per clip, block-causal mask."""
    seq = [t.view(B, T, D) for t in toks]
    if fuser.frame_level_token:
        seq = [fuser.modal_token.expand(B, -1, -1)] + seq
    n = len(seq)
    h = torch.cat(seq, dim=1)  # (B, n*T, D), token index = m*T + t
    h = h + fuser.position_embeddings.weight[:T].repeat(n, 1)
    if fuser.modality_embedding is not None:
        h = h + fuser.modality_embedding.repeat_interleave(T, dim=0)
    h = F.dropout(h, _rate(fuser.embd_drop), training).reshape(B * n * T, D)
    attns = []
    for blk in fuser.blocks:
        h, p = _block(h, blk, B, n * T, 2, T, training)
        attns.append(p)
    x = _ln(h, fuser.norm).view(B, n, T, D)
    z = x[:, 0] if fuser.frame_level_token else x.mean(dim=1)
    return z.reshape(B, T, D), torch.stack(attns).transpose(0, 1).detach()


def _fuse_ca(fuser, toks, B, T, D, training):
    """TemporalCrossAttentFuser (reference models/fusion.py:243-270) with DecoderBlock (transformerblock.py:157-162)."""
    pos = fuser.position_embeddings.weight[:T]
    seq = [F.dropout(t.view(B, T, D) + pos, _rate(fuser.embd_drop), training).reshape(B * T, D) for t in toks]
    x, mems = seq[0], seq[1:]
    for i, blk in enumerate(fuser.blocks):
        dp = getattr(blk.drop_path, "drop_prob", 0.0) or 0.0
        a, _ = _self_attention(_ln(x, blk.norm_self), blk.attn, B, T, 1, T, training)
        x = x + _drop_path(a, dp, training, T)
        ca = blk.cross_attn
        H = ca.num_heads
        mem = _ln(mems[i], blk.norm_kv)
        qkv = torch.cat([_linear(_ln(x, blk.norm_q), ca.w_q), _linear(mem, ca.w_k), _linear(mem, ca.w_v)], dim=1)
        c, _ = _attn_proj(qkv, ca.proj, B, T, H, D // H, 1, T, _rate(ca.attn_drop) if training else 0.0)
        c = F.dropout(c, _rate(ca.proj_drop), training)
        x = x + _drop_path(c, dp, training, T)
        f = _mlp(_ln(x, blk.norm_mlp), blk.mlp.mlp[0], blk.mlp.mlp[2], _capi.ACT_GELU_ERF)
        f = F.dropout(f, _rate(blk.mlp.mlp[3]), training)
        x = x + _drop_path(f, dp, training, T)
    return _ln(x, fuser.norm).view(B, T, D), torch.zeros(B)  # the reference's dummy attention (:269)


def forward_train(fp, feats: Dict[str, torch.Tensor]) -> Dict[str, Dict[str, torch.Tensor]]:
    """CMFPEarly.forward (reference models/future_prediction.py:257-291) in training mode, differentiable, for every
    fuser (SA / SA without token / T-SA / CA) and every feature mapping.  All stochastic regularisers of the reference
    are applied: embedding / projection / MLP dropout, attention-probability dropout (in the attention kernel), DropPath,
    classifier dropout, GPT-2 embd / attn / resid dropout."""
    from .models import fusion
    fuser = fp.fuser
    if fp.fp_output_len != 1:
        raise NotImplementedError("training step: fp_output_len > 1 (autoregressive roll-out) is built for inference only")
    training = fp.training
    order = [m for m in fp.modal_feature_order if m in feats]
    first = feats[order[0]]
    B, T = first.shape[0], first.shape[1]
    D = fp.latent_dim
    toks = []
    for m in order:  # feature mapping (models/future_prediction.py:133-142)
        toks.append(_apply_mapping(fp.mapping[m], feats[m].reshape(B * T, -1).float(), training))
    if isinstance(fuser, fusion.ModalTokenCMFuser):
        z, attn = _fuse_sa(fuser, toks, B, T, D, training, with_token=True)
    elif isinstance(fuser, fusion.CMFuser):
        z, attn = _fuse_sa(fuser, toks, B, T, D, training, with_token=False)
    elif isinstance(fuser, fusion.TemporalCMFuser):
        z, attn = _fuse_tsa(fuser, toks, B, T, D, training)
    elif isinstance(fuser, fusion.TemporalCrossAttentFuser):
        z, attn = _fuse_ca(fuser, toks, B, T, D, training)
    else:
        raise NotImplementedError(f"training step: unknown fuser {type(fuser).__name__}")

    gpt = fp.future_predictor.gpt_model
    G, H2 = gpt.n_embd, gpt.n_head
    identity = isinstance(fp.dim_encoder, torch.nn.Identity)  # common_dim == fp_inter_dim (future_prediction.py:245-255)
    g = (z if identity else _linear(z.reshape(B * T, D), fp.dim_encoder).view(B, T, G)) + gpt.wpe.weight[:T]
    g = F.dropout(g, gpt.drop.p, training).reshape(B * T, G)
    for blk in gpt.h:  # transformers GPT2Block
        y = _ln(g, blk.ln_1)
        a, _ = _attn_proj(_linear(y, blk.attn.c_attn, conv1d=True), blk.attn.c_proj, B, T, H2, G // H2, 1, T,
                          _rate(blk.attn.attn_dropout) if training else 0.0, conv1d=True)
        g = g + F.dropout(a, blk.attn.resid_dropout.p, training)
        y = _ln(g, blk.ln_2)
        f = _mlp(y, blk.mlp.c_fc, blk.mlp.c_proj, _capi.ACT_GELU_TANH, conv1d=True)
        g = g + F.dropout(f, blk.mlp.dropout.p, training)
    g = _ln(g, gpt.ln_f)
    z_hat = (g if identity else _linear(g, fp.dim_decoder)).view(B, T, D)

    past_futures = torch.cat([z[:, :1], z_hat[:, :T - 1]], dim=1)  # models/future_prediction.py:172-176
    future = z_hat[:, T - 1:]
    out = {"orig_past": {"all-fused": z}, "future": {"all-fused": future}, "all-fused": {"all-fused": z[:, T - 1:]},
           "past_futures": {"all-fused": past_futures}}
    for cls, c in fp.num_classes.items():
        head = fp.classifiers[cls]["all-fused"]
        for prefix, src in (("past_", past_futures), ("", future)):
            s = F.dropout(src.reshape(-1, D), head[0].p, training)
            out[f"{prefix}logits/{cls}"] = {"all-fused": _linear(s, head[1]).reshape(B, -1, c)}
    out["attentions"] = {"all-fused": {"modality_attns": attn, "temporal_attns": {}}}
    return out


def grad_groups(fp) -> list:
    """Parameter groups of a CMFPEarly head in the order the backward pass finishes them (last layers first): the bucket
    layout of ``afft_b200.dist.GradBuckets`` (reference train.py:366-368 leaves this to DistributedDataParallel)."""
    groups = [[p for cls in fp.classifiers.values() for p in cls.parameters()] + list(fp.dim_decoder.parameters())]  # Identity: none
    gpt = fp.future_predictor.gpt_model
    groups[0] += list(gpt.ln_f.parameters())
    for blk in reversed(gpt.h):
        groups.append(list(blk.parameters()))
    tail = list(gpt.wpe.parameters()) + list(fp.dim_encoder.parameters())
    fuser = fp.fuser
    tail += list(fuser.norm.parameters())
    groups.append(tail)
    for blk in reversed(fuser.blocks):
        groups.append(list(blk.parameters()))
    seen = {id(p) for g in groups for p in g}
    rest = [p for p in fp.parameters() if id(p) not in seen]  # tokens, embeddings, mappings: finished last
    if rest:
        groups.append(rest)
    return groups


def merge_groups(groups: list, n_buckets: int) -> list:
    """Merge consecutive layer groups into n_buckets all-reduce buckets of similar size (0: one bucket per group).
    Fewer, larger collectives cost less launch latency per step; more, smaller ones start earlier in the backward pass."""
    if n_buckets <= 0 or n_buckets >= len(groups):
        return groups
    sizes = [sum(p.numel() for p in g) for g in groups]
    target = sum(sizes) / n_buckets
    out, cur, acc = [], [], 0
    for g, n in zip(groups, sizes):
        cur += g
        acc += n
        if acc >= target and len(out) < n_buckets - 1:
            out.append(cur)
            cur, acc = [], 0
    if cur:
        out.append(cur)
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
