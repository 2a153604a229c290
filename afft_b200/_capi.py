"""ctypes binding of the afft_b200 C ABI (include/afft_b200.h).

This module is the only place Python touches the native library.  There is no fallback: if the
shared library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_LIB_NAME = "libafft_b200.so"
_LIB_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_lib")
# AFFT_B200_LIB selects another build of the library (A/B timing of kernel variants); default: the in-tree build
LIB_PATH = os.environ.get("AFFT_B200_LIB") or os.path.join(_LIB_DIR, _LIB_NAME)

AFFT_OK = 0
AFFT_MAX_MODS = 8
AFFT_MAX_CLS = 4
AFFT_NAME_LEN = 32
ABI_VERSION = 8

ACT_NONE, ACT_GELU_ERF, ACT_GELU_TANH, ACT_RELU, ACT_GATE = 0, 1, 2, 3, 4
PREC_BF16, PREC_BF16X3, PREC_FP16 = 0, 1, 2
DT_BF16, DT_F32, DT_FP16 = 0, 1, 2
# precision names of the Python API -> AFFT_PREC_* (include/afft_b200.h)
PRECISIONS = {"bf16": PREC_BF16, "strict": PREC_BF16X3, "fp16": PREC_FP16}


def resolve_precision(strict=False, precision=None) -> str:
    """The one precision name ('bf16' | 'fp16' | 'strict') behind the (strict=, precision=) constructor arguments."""
    if precision is None:
        return "strict" if strict else "bf16"
    if precision not in PRECISIONS:
        raise ValueError(f"precision must be one of {sorted(PRECISIONS)}, got {precision!r}")
    if strict and precision != "strict":
        raise ValueError("strict=True contradicts precision=%r" % (precision,))
    return precision


def operand_dtype(precision: str):
    """torch dtype of the 16-bit GEMM operands of a precision."""
    return torch.float16 if precision == "fp16" else torch.bfloat16
FUSER_SA, FUSER_SA_NOTOKEN, FUSER_TSA, FUSER_CA, FUSER_NONE = 0, 1, 2, 3, 4
STAGE_ALL, STAGE_FUSER, STAGE_GPT = 0, 1, 4


class AfftError(RuntimeError):
    """A call into libafft_b200 returned a non-zero status."""


class GemmDesc(C.Structure):
    _fields_ = [
        ("a_hi", C.c_void_p), ("a_lo", C.c_void_p), ("lda", C.c_int64),
        ("w_hi", C.c_void_p), ("w_lo", C.c_void_p), ("ldw", C.c_int64),
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32), ("precision", C.c_int32),
        ("bias", C.c_void_p), ("res", C.c_void_p), ("ld_res", C.c_int64),
        ("res_mod", C.c_int32), ("act", C.c_int32),
        ("out_f32", C.c_void_p), ("ld_f32", C.c_int64),
        ("out_hi", C.c_void_p), ("out_lo", C.c_void_p), ("ld_bf16", C.c_int64),
        ("row_group", C.c_int32), ("row_stride", C.c_int32), ("row_off", C.c_int32),
        ("force_block_n", C.c_int32),
    ]


class LayerNormDesc(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("ldx", C.c_int64),
        ("in_group", C.c_int32), ("in_stride", C.c_int32), ("n_avg", C.c_int32), ("avg_stride", C.c_int32),
        ("gamma", C.c_void_p), ("beta", C.c_void_p), ("eps", C.c_float),
        ("rows", C.c_int32), ("dim", C.c_int32),
        ("y_f32", C.c_void_p), ("y_hi", C.c_void_p), ("y_lo", C.c_void_p), ("ldy", C.c_int64),
        ("aux_mod", C.c_int32), ("aux_stride", C.c_int32), ("aux_rem", C.c_int32),
        ("aux_f32", C.c_void_p), ("aux_hi", C.c_void_p), ("aux_lo", C.c_void_p), ("ld_aux", C.c_int64),
        ("out_fp16", C.c_int32),
    ]


class AttentionDesc(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p),
        ("ldq", C.c_int64), ("ldk", C.c_int64), ("ldv", C.c_int64),
        ("in_dtype", C.c_int32),
        ("n_seq", C.c_int32), ("L", C.c_int32), ("H", C.c_int32), ("head_dim", C.c_int32),
        ("scale", C.c_float), ("mask", C.c_int32), ("T", C.c_int32),
        ("out_hi", C.c_void_p), ("out_lo", C.c_void_p), ("ldo", C.c_int64),
        ("probs", C.c_void_p), ("p_outer", C.c_int64), ("p_inner_stride", C.c_int64), ("p_inner", C.c_int32),
        ("drop_mask", C.c_void_p),
    ]


class Config(C.Structure):
    _fields_ = [
        ("fuser_kind", C.c_int32), ("T", C.c_int32), ("n_mod", C.c_int32),
        ("mod_name", (C.c_char * AFFT_NAME_LEN) * AFFT_MAX_MODS),
        ("mod_dim", C.c_int32 * AFFT_MAX_MODS),
        ("dim", C.c_int32), ("fuser_depth", C.c_int32), ("fuser_heads", C.c_int32),
        ("modal_encoding", C.c_int32), ("frame_level_token", C.c_int32), ("cross_attn", C.c_int32),
        ("norm_elementwise", C.c_int32),
        ("gpt_dim", C.c_int32), ("gpt_layers", C.c_int32), ("gpt_heads", C.c_int32),
        ("n_cls", C.c_int32),
        ("cls_name", (C.c_char * AFFT_NAME_LEN) * AFFT_MAX_CLS),
        ("cls_dim", C.c_int32 * AFFT_MAX_CLS),
        ("precision", C.c_int32), ("max_batch", C.c_int32), ("device", C.c_int32), ("fp_output_len", C.c_int32),
        ("stages", C.c_int32),
    ]


class IO(C.Structure):
    _fields_ = [
        ("feat", C.c_void_p * AFFT_MAX_MODS),
        ("orig_past", C.c_void_p), ("past_futures", C.c_void_p),
        ("logits", C.c_void_p * AFFT_MAX_CLS), ("ld_logits", C.c_int64 * AFFT_MAX_CLS),
        ("fuser_attn", C.c_void_p), ("gpt_attn", C.c_void_p),
    ]


class ProfileRec(C.Structure):
    _fields_ = [("cat", C.c_int32), ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32), ("ms", C.c_float)]


AFFT_MAX_PROFILE_RECS = 256


class Profile(C.Structure):
    _fields_ = [("n", C.c_int32), ("recs", ProfileRec * AFFT_MAX_PROFILE_RECS)]


# every symbol include/afft_b200.h declares
EXPORTED_SYMBOLS = [
    "afft_abi_version", "afft_last_error", "afft_gemm", "afft_set_gemm_epilogue", "afft_set_gemm_skinny", "afft_convert_bf16",
    "afft_convert_operand", "afft_layernorm", "afft_attention", "afft_create", "afft_destroy", "afft_handle_error",
    "afft_workspace_bytes", "afft_weight_bytes", "afft_set_weight", "afft_missing_weights", "afft_forward",
    "afft_last_launch_count", "afft_profile_enable", "afft_profile_read", "afft_set_max_ksplit", "afft_plan_ksplit",
    "afft_marginalize_topk", "afft_score_fusion", "afft_transpose_bf16", "afft_layernorm_bwd", "afft_gelu_fwd",
    "afft_gelu_bwd", "afft_colsum", "afft_attention_bwd", "afft_sgd_nesterov", "afft_convert_dual",
    "afft_convert_dual_gelu", "afft_workspace_bytes_for", "afft_create_in",
]

_lib: Optional[C.CDLL] = None


def lib() -> C.CDLL:
    """Load libafft_b200.so (built by __graft_entry__.build()).  Raises if it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AfftError(
            f"{LIB_PATH} not found: the CUDA extension is not built. Run `python -c 'import __graft_entry__ as g; "
            f"g.build()'` (or `make -C afft_b200/csrc`). There is no CPU/PyTorch fallback for the AFFT hot path.")
    l = C.CDLL(LIB_PATH)
    l.afft_abi_version.restype = C.c_int
    l.afft_last_error.restype = C.c_char_p
    l.afft_gemm.argtypes = [C.POINTER(GemmDesc), C.c_void_p]
    l.afft_convert_bf16.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64,
                                    C.c_int32, C.c_void_p]
    l.afft_convert_operand.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64,
                                       C.c_int32, C.c_int32, C.c_void_p]
    l.afft_convert_operand.restype = C.c_int
    l.afft_set_gemm_epilogue.argtypes = [C.c_int32]
    l.afft_set_gemm_epilogue.restype = C.c_int
    l.afft_set_gemm_skinny.argtypes = [C.c_int32]
    l.afft_set_gemm_skinny.restype = C.c_int
    l.afft_layernorm.argtypes = [C.POINTER(LayerNormDesc), C.c_void_p]
    l.afft_attention.argtypes = [C.POINTER(AttentionDesc), C.c_void_p]
    l.afft_create.argtypes = [C.POINTER(Config), C.POINTER(C.c_void_p)]
    l.afft_workspace_bytes_for.argtypes = [C.POINTER(Config), C.POINTER(C.c_size_t)]
    l.afft_workspace_bytes_for.restype = C.c_int
    l.afft_create_in.argtypes = [C.POINTER(Config), C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(C.c_void_p)]
    l.afft_create_in.restype = C.c_int
    l.afft_destroy.argtypes = [C.c_void_p]
    l.afft_destroy.restype = None
    l.afft_handle_error.argtypes = [C.c_void_p]
    l.afft_handle_error.restype = C.c_char_p
    l.afft_workspace_bytes.argtypes = [C.c_void_p]
    l.afft_workspace_bytes.restype = C.c_size_t
    l.afft_weight_bytes.argtypes = [C.c_void_p]
    l.afft_weight_bytes.restype = C.c_size_t
    l.afft_set_weight.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int32, C.POINTER(C.c_int64), C.c_void_p]
    l.afft_missing_weights.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
    l.afft_forward.argtypes = [C.c_void_p, C.c_int32, C.POINTER(IO), C.c_void_p]
    l.afft_last_launch_count.argtypes = [C.c_void_p]
    l.afft_marginalize_topk.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                                        C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
    l.afft_marginalize_topk.restype = C.c_int
    l.afft_score_fusion.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.POINTER(C.c_void_p), C.c_int64, C.c_int32, C.c_int32,
                                    C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    l.afft_score_fusion.restype = C.c_int
    l.afft_transpose_bf16.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p]
    l.afft_layernorm_bwd.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_float, C.c_void_p, C.c_int64, C.c_int32, C.c_int32,
                                     C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    l.afft_gelu_fwd.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]
    l.afft_gelu_bwd.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]
    l.afft_colsum.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    l.afft_convert_dual.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                    C.c_void_p, C.c_void_p]
    l.afft_convert_dual.restype = C.c_int
    l.afft_convert_dual_gelu.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p,
                                         C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]
    l.afft_convert_dual_gelu.restype = C.c_int
    l.afft_sgd_nesterov.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_float, C.c_float,
                                    C.c_int32, C.c_void_p]
    l.afft_sgd_nesterov.restype = C.c_int
    l.afft_attention_bwd.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_int32,
                                     C.c_int32, C.c_int32, C.c_float, C.c_void_p, C.c_void_p]
    for _n in ("afft_transpose_bf16", "afft_layernorm_bwd", "afft_gelu_fwd", "afft_gelu_bwd", "afft_colsum", "afft_attention_bwd"):
        getattr(l, _n).restype = C.c_int
    l.afft_profile_enable.argtypes = [C.c_void_p, C.c_int32]
    l.afft_set_max_ksplit.argtypes = [C.c_void_p, C.c_int32]
    l.afft_set_max_ksplit.restype = C.c_int
    l.afft_plan_ksplit.argtypes = [C.c_int32] * 5
    l.afft_plan_ksplit.restype = C.c_int
    l.afft_profile_read.argtypes = [C.c_void_p, C.POINTER(Profile)]
    for name in ("afft_gemm", "afft_convert_bf16", "afft_layernorm", "afft_attention", "afft_create",
                 "afft_set_weight", "afft_missing_weights", "afft_forward", "afft_last_launch_count",
                 "afft_profile_enable", "afft_profile_read"):
        getattr(l, name).restype = C.c_int
    if l.afft_abi_version() != ABI_VERSION:
        raise AfftError(f"ABI mismatch: library {l.afft_abi_version()} vs binding {ABI_VERSION}; rebuild the extension")
    _lib = l
    return l


def check(rc: int, handle: Optional[int] = None) -> None:
    if rc != AFFT_OK:
        l = lib()
        msg = l.afft_handle_error(handle) if handle else l.afft_last_error()
        raise AfftError(f"afft_b200 call failed (status {rc}): {msg.decode(errors='replace') if msg else ''}")


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    if not t.is_cuda:
        raise AfftError("afft_b200 operators take CUDA tensors only (no CPU fallback)")
    return t.data_ptr()


def current_stream_ptr(device=None) -> int:
    return torch.cuda.current_stream(device).cuda_stream


# ------------------------------------------------------------------------------------------------
# thin operator wrappers (used by tests and by tools/; the model path goes through afft_forward)
# ------------------------------------------------------------------------------------------------
def split_bf16(x: torch.Tensor):
    """hi/lo bf16 pair of an fp32 tensor via the library's conversion kernel."""
    assert x.dtype == torch.float32 and x.dim() == 2 and x.is_contiguous()
    hi = torch.empty_like(x, dtype=torch.bfloat16)
    lo = torch.empty_like(x, dtype=torch.bfloat16)
    check(lib().afft_convert_bf16(ptr(x), x.shape[1], x.shape[0], x.shape[1], ptr(hi), ptr(lo), x.shape[1], 0,
                                  current_stream_ptr(x.device)))
    return hi, lo


def gemm(a, w, *, a_lo=None, w_lo=None, bias=None, res=None, res_mod=0, act=ACT_NONE, out_f32=None, out_hi=None,
         out_lo=None, row_map=(0, 0, 0), force_block_n=0, M=None):
    """C = epilogue(A . W^T).  a [M,K] bf16 (row stride a.stride(0)), w [N,K] bf16."""
    d = GemmDesc()
    d.a_hi, d.a_lo, d.lda = ptr(a), ptr(a_lo), a.stride(0)
    d.w_hi, d.w_lo, d.ldw = ptr(w), ptr(w_lo), w.stride(0)
    d.M, d.N, d.K = (M if M is not None else a.shape[0]), w.shape[0], w.shape[1]
    d.precision = PREC_BF16X3 if a_lo is not None else (PREC_FP16 if a.dtype == torch.float16 else PREC_BF16)
    d.bias = ptr(bias)
    d.res, d.ld_res, d.res_mod = ptr(res), (res.stride(0) if res is not None else 0), res_mod
    d.act = act
    d.out_f32, d.ld_f32 = ptr(out_f32), (out_f32.stride(0) if out_f32 is not None else 0)
    d.out_hi, d.out_lo = ptr(out_hi), ptr(out_lo)
    d.ld_bf16 = out_hi.stride(0) if out_hi is not None else 0
    d.row_group, d.row_stride, d.row_off = row_map
    d.force_block_n = force_block_n
    check(lib().afft_gemm(C.byref(d), current_stream_ptr(a.device)))


def layernorm(x, gamma, beta, eps, *, rows=None, ldx=None, y_f32=None, y_hi=None, y_lo=None, in_map=(0, 0),
              avg=(0, 0), aux=(0, 0), aux_f32=None, aux_hi=None, aux_lo=None):
    d = LayerNormDesc()
    d.x, d.ldx = ptr(x), (ldx if ldx is not None else x.stride(0))
    d.in_group, d.in_stride = in_map
    d.n_avg, d.avg_stride = avg
    d.gamma, d.beta, d.eps = ptr(gamma), ptr(beta), eps
    d.rows, d.dim = (rows if rows is not None else x.shape[0]), x.shape[-1]
    first = next(t for t in (y_f32, y_hi, y_lo) if t is not None)
    d.y_f32, d.y_hi, d.y_lo, d.ldy = ptr(y_f32), ptr(y_hi), ptr(y_lo), first.stride(0)
    d.aux_mod, d.aux_stride = aux[0], aux[1]
    d.aux_rem = aux[2] if len(aux) > 2 else 0
    d.aux_f32, d.aux_hi, d.aux_lo = ptr(aux_f32), ptr(aux_hi), ptr(aux_lo)
    d.out_fp16 = int(any(t is not None and t.dtype == torch.float16 for t in (y_hi, aux_hi)))
    fa = next((t for t in (aux_f32, aux_hi, aux_lo) if t is not None), None)
    d.ld_aux = fa.stride(0) if fa is not None else 0
    check(lib().afft_layernorm(C.byref(d), current_stream_ptr(x.device)))


def attention(qkv, n_seq, L, H, head_dim, *, mask=0, T=1, out_hi, out_lo=None, probs=None, p_outer=0,
              p_inner_stride=0, p_inner=1, scale=None, drop_mask=None):
    """qkv [n_seq*L, 3*H*head_dim] (q | k | v), bf16 or fp32."""
    D = H * head_dim
    es = qkv.element_size()
    d = AttentionDesc()
    base = qkv.data_ptr()
    d.q, d.k, d.v = base, base + D * es, base + 2 * D * es
    d.ldq = d.ldk = d.ldv = qkv.stride(0)
    d.in_dtype = {torch.float32: DT_F32, torch.float16: DT_FP16}.get(qkv.dtype, DT_BF16)
    d.n_seq, d.L, d.H, d.head_dim = n_seq, L, H, head_dim
    d.scale = scale if scale is not None else head_dim ** -0.5
    d.mask, d.T = mask, T
    d.out_hi, d.out_lo, d.ldo = ptr(out_hi), ptr(out_lo), out_hi.stride(0)
    d.probs, d.p_outer, d.p_inner_stride, d.p_inner = ptr(probs), p_outer, p_inner_stride, p_inner
    d.drop_mask = ptr(drop_mask)
    check(lib().afft_attention(C.byref(d), current_stream_ptr(qkv.device)))


def score_fusion(attn_logits, logits=None, n_cols=0, *, attn=None, out=None):
    """p = softmax(attn_logits[:, :M]); out[:, :n_cols] = sum_i p[:, i] * logits[i][:, :n_cols] (all fp32, 2-D).
    With logits=None only the softmax is written to ``attn`` [rows, M]."""
    M = attn.shape[1] if attn is not None else len(logits)
    rows = attn_logits.shape[0]
    arr = (C.c_void_p * 8)()
    ld_l = 0
    if logits is not None:
        for i, t in enumerate(logits):
            arr[i] = ptr(t)
        ld_l = logits[0].stride(0)
    check(lib().afft_score_fusion(ptr(attn_logits), attn_logits.stride(0), M, arr if logits is not None else None, ld_l, rows,
                                  n_cols, ptr(attn), ptr(out), out.stride(0) if out is not None else 0,
                                  current_stream_ptr(attn_logits.device)))
