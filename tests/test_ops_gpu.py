"""GPU: each stateless operator of the C ABI against a plain torch reference of the same op
(float64 matmul / F.layer_norm / softmax attention).  Tolerances are stated per case."""
import pytest
import torch

from afft_b200 import _capi as capi

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    torch.backends.cuda.matmul.allow_tf32 = False
    return torch.device("cuda:0")


def _randn(gen, dev, *shape, scale=1.0):
    return (torch.randn(*shape, generator=gen) * scale).to(dev)


def _mm(a, w):
    return (a.double() @ w.double().t()).float()


# bf16 operands are exact inputs; the only error is fp32 accumulation order: ~1e-5 at K=1024, scale |out|~1
@pytest.mark.parametrize("M,N,K,bn", [(128, 128, 64, 128), (128, 256, 64, 256), (1000, 1024, 1024, 0), (90, 3072, 1024, 0),
                                      (18, 2048, 1024, 0), (576, 1024, 352, 0), (300, 3806, 1024, 0), (131, 106, 1024, 0),
                                      (1, 1024, 1024, 0), (4608, 2048, 8192, 128), (2304, 8192, 2048, 256)])
def test_gemm_plain(dev, M, N, K, bn):
    g = torch.Generator().manual_seed(M * 7 + N)
    a = _randn(g, dev, M, K).bfloat16()
    w = _randn(g, dev, N, K, scale=0.05).bfloat16()
    ld = (N + 3) // 4 * 4
    out = torch.full((M, ld), float("nan"), device=dev)
    capi.gemm(a, w, out_f32=out, force_block_n=bn)
    ref = _mm(a, w)
    assert not torch.isnan(out[:, :N]).any()
    assert (out[:, :N] - ref).abs().max().item() < 2e-4 * max(1.0, (K / 1024) ** 0.5) * 3
    if ld > N:
        assert torch.isnan(out[:, N:]).all()  # padding columns are never written


def test_gemm_epilogues(dev):
    g = torch.Generator().manual_seed(5)
    M, N, K = 777, 1024, 1024
    a = _randn(g, dev, M, K).bfloat16()
    w = _randn(g, dev, N, K, scale=0.05).bfloat16()
    bias, res = _randn(g, dev, N), _randn(g, dev, M, N)
    ref0 = _mm(a, w)
    F = torch.nn.functional
    out_f = torch.zeros(M, N, device=dev)
    capi.gemm(a, w, bias=bias, act=capi.ACT_GELU_ERF, out_f32=out_f)
    assert (out_f - F.gelu(ref0 + bias)).abs().max().item() < 1e-4  # erf approx 1.5e-7 + accumulation order
    capi.gemm(a, w, bias=bias, act=capi.ACT_GELU_TANH, out_f32=out_f)
    assert (out_f - F.gelu(ref0 + bias, approximate="tanh")).abs().max().item() < 1e-4
    out_b = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
    capi.gemm(a, w, bias=bias, act=capi.ACT_GELU_ERF, out_hi=out_b)
    assert torch.equal(out_b, F.gelu(ref0 + bias).bfloat16()) or \
        (out_b.float() - F.gelu(ref0 + bias)).abs().max().item() < 4e-2  # one bf16 ulp at |x|<8
    h = res.clone()
    capi.gemm(a, w, bias=bias, res=h, out_f32=h)  # in-place residual stream update
    assert (h - (ref0 + bias + res)).abs().max().item() < 1e-4
    T = 7
    pos = _randn(g, dev, T, N)
    outm = torch.zeros(M // T * (T + 1) + T + 1, N, device=dev)
    capi.gemm(a, w, res=pos, res_mod=T, out_f32=outm, row_map=(T, T + 1, 1))
    r = torch.arange(M, device=dev)
    assert (outm[(r // T) * (T + 1) + r % T + 1] - (ref0 + pos[r % T])).abs().max().item() < 1e-4
    slots = 5
    hbuf = torch.zeros(M * slots, N, device=dev)
    capi.gemm(a, w, out_f32=hbuf.view(M, slots * N)[:, 2 * N:3 * N])
    assert (hbuf.view(M, slots, N)[:, 2] - ref0).abs().max().item() < 1e-4
    assert hbuf.view(M, slots, N)[:, [0, 1, 3, 4]].abs().max().item() == 0.0
    Nc = 3806
    wc = _randn(g, dev, Nc, K, scale=0.05).bfloat16()
    bc = torch.zeros(Nc + 16, device=dev)[:Nc]
    bc.copy_(_randn(g, dev, Nc))
    oc = torch.full((M, 3808), 7.0, device=dev)
    capi.gemm(a, wc, bias=bc, out_f32=oc)
    assert (oc[:, :Nc] - (_mm(a, wc) + bc)).abs().max().item() < 1e-4
    assert (oc[:, Nc:] == 7.0).all()


@pytest.mark.parametrize("M,N,K,bn", [(256, 256, 1024, 128), (777, 1024, 1024, 256), (300, 3806, 1024, 0), (576, 1024, 352, 0)])
def test_gemm_strict_bf16x3(dev, M, N, K, bn):
    """hi.hi + hi.lo + lo.hi: fp32 inputs reproduced to ~2^-16 relative per product."""
    g = torch.Generator().manual_seed(M + N)
    a32, w32 = _randn(g, dev, M, K), _randn(g, dev, N, K, scale=0.05)
    a_hi, a_lo = capi.split_bf16(a32)
    w_hi, w_lo = capi.split_bf16(w32)
    assert torch.equal(a_hi, a32.bfloat16())
    assert (a_hi.float() + a_lo.float() - a32).abs().max().item() < 2e-5 * a32.abs().max().item()
    ld = (N + 3) // 4 * 4
    out = torch.zeros(M, ld, device=dev)
    capi.gemm(a_hi, w_hi, a_lo=a_lo, w_lo=w_lo, out_f32=out, force_block_n=bn)
    ref = _mm(a32, w32)
    assert (out[:, :N] - ref).abs().max().item() < 2e-4
    oh = torch.zeros(M, ld, device=dev, dtype=torch.bfloat16)
    ol = torch.zeros(M, ld, device=dev, dtype=torch.bfloat16)
    capi.gemm(a_hi, w_hi, a_lo=a_lo, w_lo=w_lo, out_hi=oh, out_lo=ol, force_block_n=bn)
    assert ((oh.float() + ol.float())[:, :N] - ref).abs().max().item() < 2e-4


def test_gemm_relu_gate_and_skinny_n(dev):
    """ReLU and ContextGating (residual * sigmoid) epilogues of the ablation mappings / MATT, and MATT's N = 4 output."""
    g = torch.Generator().manual_seed(11)
    M, N, K = 333, 1024, 352
    a = _randn(g, dev, M, K).bfloat16()
    w = _randn(g, dev, N, K, scale=0.05).bfloat16()
    bias, u = _randn(g, dev, N), _randn(g, dev, M, N)
    ref0 = _mm(a, w) + bias
    out = torch.zeros(M, N, device=dev)
    capi.gemm(a, w, bias=bias, act=capi.ACT_RELU, out_f32=out)
    assert (out - torch.relu(ref0)).abs().max().item() < 1e-4
    capi.gemm(a, w, bias=bias, act=capi.ACT_GATE, res=u, out_f32=out)
    assert (out - u * torch.sigmoid(ref0)).abs().max().item() < 1e-4
    with pytest.raises(capi.AfftError):
        capi.gemm(a, w, bias=bias, act=capi.ACT_GATE, out_f32=out)  # the gated operand is required
    w4 = _randn(g, dev, 4, K, scale=0.05).bfloat16()
    b4 = _randn(g, dev, 4)
    out4 = torch.zeros(M, 4, device=dev)
    capi.gemm(a, w4, bias=b4, out_f32=out4)
    assert (out4 - (_mm(a, w4) + b4)).abs().max().item() < 1e-4


def test_gemm_rejects_bad_arguments(dev):
    a = torch.zeros(128, 64, device=dev, dtype=torch.bfloat16)
    w = torch.zeros(128, 64, device=dev, dtype=torch.bfloat16)
    with pytest.raises(capi.AfftError):
        capi.gemm(a, w)  # no output
    with pytest.raises(capi.AfftError):
        capi.gemm(a, w, out_f32=torch.zeros(128, 130, device=dev)[:, 1:129])  # misaligned pitch/pointer
    with pytest.raises(capi.AfftError):  # TMA needs 16-B multiple operand pitches
        capi.gemm(torch.zeros(128, 68, device=dev, dtype=torch.bfloat16)[:, :64], w, out_f32=torch.zeros(128, 128, device=dev))
    torch.cuda.synchronize()


@pytest.mark.parametrize("dim,eps", [(1024, 1e-6), (2048, 1e-5), (512, 1e-6)])
def test_layernorm(dev, dim, eps):
    g = torch.Generator().manual_seed(dim)
    rows = 1237
    x = _randn(g, dev, rows, dim) * 3 + 0.5
    gm, bt = _randn(g, dev, dim), _randn(g, dev, dim)
    yf = torch.zeros(rows, dim, device=dev)
    yh = torch.zeros(rows, dim, device=dev, dtype=torch.bfloat16)
    yl = torch.zeros(rows, dim, device=dev, dtype=torch.bfloat16)
    capi.layernorm(x, gm, bt, eps, y_f32=yf, y_hi=yh, y_lo=yl)
    ref = torch.nn.functional.layer_norm(x.double(), (dim,), gm.double(), bt.double(), eps).float()
    assert (yf - ref).abs().max().item() < 2e-5
    assert torch.equal(yh, yf.bfloat16())
    assert (yh.float() + yl.float() - ref).abs().max().item() < 2e-4
    capi.layernorm(x, None, None, eps, y_f32=yf)  # elementwise_affine=False
    assert (yf - torch.nn.functional.layer_norm(x, (dim,), None, None, eps)).abs().max().item() < 2e-5


def test_layernorm_row_maps(dev):
    g = torch.Generator().manual_seed(3)
    B, T, n, dim = 5, 6, 4, 1024
    x = _randn(g, dev, B * T * n, dim)
    gm, bt = _randn(g, dev, dim), _randn(g, dev, dim)
    F = torch.nn.functional
    yf = torch.zeros(B * T, dim, device=dev)
    aux = torch.zeros(B * (T + 1), dim, device=dev)
    capi.layernorm(x, gm, bt, 1e-6, rows=B * T, ldx=n * dim, y_f32=yf, aux=(T, T + 1), aux_f32=aux)
    ref = F.layer_norm(x.view(B * T, n, dim)[:, 0], (dim,), gm, bt, 1e-6)
    assert (yf - ref).abs().max().item() < 2e-5
    assert (aux.view(B, T + 1, dim)[:, 0] - ref.view(B, T, dim)[:, 0]).abs().max().item() < 2e-5
    assert aux.view(B, T + 1, dim)[:, 1:].abs().max().item() == 0
    capi.layernorm(x, gm, bt, 1e-6, rows=B * T, ldx=dim, y_f32=yf, in_map=(1, n), avg=(n, 1))
    assert (yf - F.layer_norm(x.view(B * T, n, dim), (dim,), gm, bt, 1e-6).mean(1)).abs().max().item() < 2e-5
    # T-SA layout: first T tokens of each clip's n*T-token sequence
    capi.layernorm(x, gm, bt, 1e-6, rows=B * T, ldx=dim, y_f32=yf, in_map=(T, n * T))
    assert (yf - F.layer_norm(x.view(B, n * T, dim)[:, :T].reshape(B * T, dim), (dim,), gm, bt, 1e-6)).abs().max().item() < 2e-5


def _ref_attn(qkv, n_seq, L, H, hd, mask, T):
    dev = qkv.device
    x = qkv.float().view(n_seq, L, 3, H, hd).permute(2, 0, 3, 1, 4)
    q, k, v = x[0], x[1], x[2]
    s = (q @ k.transpose(-1, -2)) * hd ** -0.5
    i = torch.arange(L, device=dev)[:, None]
    j = torch.arange(L, device=dev)[None, :]
    if mask == 1:
        s = s.masked_fill(j > i, float("-inf"))
    elif mask == 2:
        s = s.masked_fill((j % T) > (i % T), float("-inf"))
    elif mask == 3:
        s = s.masked_fill(j == i, float("-inf"))
    p = s.softmax(-1)
    return (p @ v).transpose(1, 2).reshape(n_seq * L, H * hd), p


@pytest.mark.parametrize("n_seq,L,H,hd,mask,T,dt", [
    (36, 5, 4, 256, 0, 1, torch.bfloat16), (7, 18, 4, 512, 1, 18, torch.bfloat16), (3, 50, 4, 256, 2, 10, torch.bfloat16),
    (36, 5, 4, 256, 3, 1, torch.float32), (7, 18, 4, 512, 1, 18, torch.float32), (3, 50, 4, 256, 2, 10, torch.float32),
    (2, 1, 4, 256, 0, 1, torch.bfloat16), (2, 64, 4, 256, 1, 64, torch.bfloat16), (5, 10, 4, 256, 1, 10, torch.bfloat16),
    (5, 32, 4, 512, 1, 32, torch.bfloat16), (5, 16, 4, 256, 0, 16, torch.bfloat16), (3, 7, 4, 512, 1, 7, torch.bfloat16),
    (37, 3, 4, 256, 0, 1, torch.bfloat16), (37, 4, 4, 256, 0, 1, torch.bfloat16), (38, 5, 4, 256, 3, 1, torch.bfloat16),
    (7, 6, 4, 256, 0, 1, torch.bfloat16), (9, 2, 4, 256, 0, 1, torch.bfloat16), (36, 5, 8, 256, 0, 1, torch.bfloat16),
    (3, 40, 4, 256, 2, 8, torch.bfloat16), (2, 64, 4, 256, 2, 16, torch.bfloat16), (2, 33, 4, 256, 0, 33, torch.bfloat16),
    (2, 40, 4, 512, 1, 40, torch.bfloat16)])
def test_attention(dev, n_seq, L, H, hd, mask, T, dt):
    g = torch.Generator().manual_seed(L * 31 + hd)
    D = H * hd
    qkv = _randn(g, dev, n_seq * L, 3 * D).to(dt)
    oh = torch.zeros(n_seq * L, D, device=dev, dtype=torch.bfloat16)
    ol = torch.zeros(n_seq * L, D, device=dev, dtype=torch.bfloat16)
    probs = torch.zeros(n_seq, H, L, L, device=dev)
    capi.attention(qkv, n_seq, L, H, hd, mask=mask, T=T, out_hi=oh, out_lo=ol, probs=probs, p_outer=H * L * L)
    ro, rp = _ref_attn(qkv, n_seq, L, H, hd, mask, T)
    # bf16 inputs with 6 < L <= 32 run on the tensor-core kernel: P is re-quantised to bf16 for P.V and the
    # output has no lo part (2^-9 relative on |v| ~ 3); everything else is fp32 math + hi/lo outputs.
    mma_path = dt == torch.bfloat16 and ((6 < L <= 32 and mask in (0, 1, 2)) or (32 < L <= 64 and hd == 256 and mask in (0, 1, 2))
                                         or (2 <= L <= 6 and hd == 256 and mask in (0, 3)))
    if mma_path:
        assert (oh.float() - ro).abs().max().item() < 3e-2
    else:
        assert (oh.float() + ol.float() - ro).abs().max().item() < 2e-4
    assert (probs - rp).abs().max().item() < 2e-5
    assert (probs.sum(-1) - 1).abs().max().item() < 1e-5
    with pytest.raises(capi.AfftError):
        capi.attention(qkv, n_seq, 65, H, hd, out_hi=oh)


# ------------------------------------------------------------------------------------------------
# fp16-operand mode (AFFT_PREC_FP16): same kernels, fp16 A / W / 16-bit outputs, fp32 accumulation
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,K,bn", [(128, 128, 64, 128), (1000, 1024, 1024, 0), (90, 3072, 1024, 0), (576, 1024, 352, 0),
                                      (300, 3806, 1024, 0), (1, 1024, 1024, 0), (2304, 8192, 2048, 256)])
def test_gemm_fp16_operands(dev, M, N, K, bn):
    """fp16 operands are exact inputs of the tensor core: the only error is fp32 accumulation order."""
    g = torch.Generator().manual_seed(M * 7 + N + 1)
    a = _randn(g, dev, M, K).half()
    w = _randn(g, dev, N, K, scale=0.05).half()
    ld = (N + 3) // 4 * 4
    out = torch.full((M, ld), float("nan"), device=dev)
    capi.gemm(a, w, out_f32=out, force_block_n=bn)
    ref = _mm(a, w)
    assert not torch.isnan(out[:, :N]).any()
    assert (out[:, :N] - ref).abs().max().item() < 2e-4 * max(1.0, (K / 1024) ** 0.5) * 3


def test_gemm_fp16_epilogues_and_saturation(dev):
    g = torch.Generator().manual_seed(55)
    M, N, K = 777, 1024, 1024
    a = _randn(g, dev, M, K).half()
    w = _randn(g, dev, N, K, scale=0.05).half()
    bias, res = _randn(g, dev, N), _randn(g, dev, M, N)
    ref0 = _mm(a, w)
    F = torch.nn.functional
    out_h = torch.zeros(M, N, device=dev, dtype=torch.float16)
    capi.gemm(a, w, bias=bias, act=capi.ACT_GELU_ERF, out_hi=out_h)  # FC1 of the fuser MLP: sigmoid form of erf-GELU
    ref = F.gelu(ref0 + bias)
    assert (out_h.float() - ref).abs().max().item() < 4e-3  # one fp16 ulp at |x| < 8 (3.9e-3) + 2.5e-5 approximation
    assert (out_h.float() - ref).abs().mean().item() < 2e-4
    capi.gemm(a, w, bias=bias, act=capi.ACT_GELU_TANH, out_hi=out_h)
    assert (out_h.float() - F.gelu(ref0 + bias, approximate="tanh")).abs().max().item() < 4e-3
    out_f = torch.zeros(M, N, device=dev)
    capi.gemm(a, w, bias=bias, res=res, out_f32=out_f, out_hi=out_h)
    assert (out_f - (ref0 + bias + res)).abs().max().item() < 1e-4
    assert torch.equal(out_h, out_f.half())
    # saturation instead of inf: a huge bias drives the fp16 output beyond 65504
    big = torch.full((N,), 1e6, device=dev)
    capi.gemm(a, w, bias=big, out_hi=out_h)
    assert torch.isfinite(out_h).all() and (out_h == 65504).all()
    x = torch.tensor([[1e9, -1e9, 70000.0, 1.0, 0.0, -3.0, 65504.0, 6e-8]], device=dev)
    h = torch.zeros(1, 8, device=dev, dtype=torch.float16)
    capi.check(capi.lib().afft_convert_operand(x.data_ptr(), 8, 1, 8, h.data_ptr(), None, 8, 0, capi.PREC_FP16,
                                               capi.current_stream_ptr(dev)))
    assert torch.equal(h, x.clamp(-65504, 65504).half())


def test_layernorm_fp16_output(dev):
    g = torch.Generator().manual_seed(77)
    rows, dim = 1237, 1024
    x = _randn(g, dev, rows, dim) * 3 + 0.5
    gm, bt = _randn(g, dev, dim), _randn(g, dev, dim)
    yf = torch.zeros(rows, dim, device=dev)
    yh = torch.zeros(rows, dim, device=dev, dtype=torch.float16)
    capi.layernorm(x, gm, bt, 1e-6, y_f32=yf, y_hi=yh)
    assert torch.equal(yh, yf.half())
    yh2 = torch.zeros_like(yh)
    capi.layernorm(x, gm, bt, 1e-6, y_hi=yh2)  # the hot (16-bit output only) instantiation
    assert torch.equal(yh2, yh)


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("n_seq,L,H,hd,mask,T", [(36, 5, 4, 256, 0, 1), (38, 5, 4, 256, 3, 1), (7, 18, 4, 512, 1, 18),
                                                 (3, 50, 4, 256, 2, 10), (5, 10, 4, 256, 1, 10), (37, 4, 4, 256, 0, 1),
                                                 (2, 64, 4, 256, 1, 64), (2, 1, 4, 256, 0, 1)])
def test_attention_tensor_core_paths(dev, n_seq, L, H, hd, mask, T, dt):
    """The mma.sync attention kernels the forward uses (16-bit inputs, no lo output), bf16 and fp16 operands."""
    g = torch.Generator().manual_seed(L * 31 + hd + 1)
    D = H * hd
    qkv = _randn(g, dev, n_seq * L, 3 * D).to(dt)
    oh = torch.zeros(n_seq * L, D, device=dev, dtype=dt)
    probs = torch.zeros(n_seq, H, L, L, device=dev)
    capi.attention(qkv, n_seq, L, H, hd, mask=mask, T=T, out_hi=oh, probs=probs, p_outer=H * L * L)
    ro, rp = _ref_attn(qkv, n_seq, L, H, hd, mask, T)
    # P is re-quantised to the operand format for P.V and the output is rounded to it: 2^-9 (bf16) / 2^-12 (fp16) on |v| ~ 3
    assert (oh.float() - ro).abs().max().item() < (3e-2 if dt == torch.bfloat16 else 4e-3)
    assert (probs - rp).abs().max().item() < 2e-5
    assert (probs.sum(-1) - 1).abs().max().item() < 1e-5


def test_hostops_dense_pads_odd_widths(dev):
    """MATT with dim=None has widths that are not multiples of 8 (3424 -> 856 -> 428): operands are zero-padded."""
    from afft_b200 import hostops
    g = torch.Generator().manual_seed(9)
    x, w, b = _randn(g, dev, 37, 428), _randn(g, dev, 107, 428, scale=0.05), _randn(g, dev, 107)
    for precision, tol in (("bf16", 1e-4), ("fp16", 1e-4), ("strict", 2e-4)):
        y = hostops.dense(x, w, b, hostops.WeightCache(), precision=precision)
        if precision == "strict":
            ref = _mm(x, w) + b
        else:
            dt = torch.bfloat16 if precision == "bf16" else torch.float16
            ref = _mm(x.to(dt), w.to(dt)) + b
        assert y.shape == (37, 107)
        assert (y - ref).abs().max().item() < tol, precision


def test_score_fusion_many_rows(dev):
    """rows beyond the 65535 gridDim.y limit (CMFPScoreFusion at B >= 5958, T = 10)."""
    g = torch.Generator().manual_seed(4)
    rows, M, C = 70001, 3, 12
    scores = _randn(g, dev, rows, 4)
    logits = [_randn(g, dev, rows, C) for _ in range(M)]
    attn = torch.empty(rows, M, device=dev)
    out = torch.empty(rows, C, device=dev)
    capi.score_fusion(scores, logits, C, attn=attn, out=out)
    p = torch.softmax(scores[:, :M].double(), dim=-1)
    ref = sum(p[:, i:i + 1] * logits[i].double() for i in range(M))
    assert (attn.double() - p).abs().max().item() < 1e-6
    assert (out.double() - ref).abs().max().item() < 1e-5


def test_gemm_tma_staged_epilogue_matches_default(dev):
    """The opt-in TMA-staged epilogue (afft_set_gemm_epilogue(1): residual fetched by cp.async.bulk.tensor, outputs stored
    by cp.async.bulk.tensor) computes the same per-element arithmetic as the default one: results are bit-identical."""
    g = torch.Generator().manual_seed(21)
    F = torch.nn.functional
    lib = capi.lib()
    try:
        for M, N, K in ((777, 1024, 1024), (1300, 3072, 512), (2304, 2048, 2048)):
            a = _randn(g, dev, M, K).bfloat16()
            w = _randn(g, dev, N, K, scale=0.05).bfloat16()
            bias, res = _randn(g, dev, N), _randn(g, dev, M, N)
            outs = {}
            for v2 in (0, 1):
                capi.check(lib.afft_set_gemm_epilogue(v2))
                h = res.clone()
                capi.gemm(a, w, bias=bias, res=h, out_f32=h)                       # in-place residual stream
                hb = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
                h2 = res.clone()
                capi.gemm(a, w, bias=bias, res=h2, out_f32=h2, out_hi=hb)          # + 16-bit copy
                qb = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
                capi.gemm(a, w, out_hi=qb)                                         # qkv-style
                gb = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
                capi.gemm(a, w, bias=bias, act=capi.ACT_GELU_ERF, out_hi=gb)       # FC1-style
                f = torch.zeros(M, N, device=dev)
                capi.gemm(a, w, bias=bias, out_f32=f)                              # classifier-style
                ah, wh = a.half(), w.half()
                fh = torch.zeros(M, N, device=dev, dtype=torch.float16)
                capi.gemm(ah, wh, bias=bias, act=capi.ACT_GELU_TANH, out_hi=fh)    # fp16 operands
                outs[v2] = (h, h2, hb, qb, gb, f, fh)
            for t0, t1 in zip(outs[0], outs[1]):
                assert torch.equal(t0, t1)
            assert (outs[1][0] - (_mm(a, w) + bias + res)).abs().max().item() < 1e-4
    finally:
        capi.check(lib.afft_set_gemm_epilogue(0))


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("M,N,K", [(1, 1024, 1024), (16, 8, 64), (18, 2048, 8192), (18, 8192, 2048), (18, 3806, 1024),
                                   (33, 1024, 352), (36, 6144, 2048), (64, 1001, 1000), (90, 3072, 1024), (90, 1024, 4096),
                                   (96, 4096, 1024), (5, 4, 2048)])
def test_gemm_skinny_kernel(dev, M, N, K, dt):
    """M <= 96 rows run the weight-streaming mma.sync kernel (csrc/gemm_skinny.cuh): against float64 on the same 16-bit
    operands, and against the tcgen05 kernels on the same problem (afft_set_gemm_skinny(0))."""
    g = torch.Generator().manual_seed(M * 13 + N + K)
    a = _randn(g, dev, M, K).to(dt)
    w = _randn(g, dev, N, K, scale=0.05).to(dt)
    ld = (N + 3) // 4 * 4
    ref = _mm(a, w)
    tol = 2e-4 * max(1.0, (K / 1024) ** 0.5) * 3
    outs = []
    try:
        for skinny in (1, 0):
            capi.check(capi.lib().afft_set_gemm_skinny(skinny))
            out = torch.full((M, ld), float("nan"), device=dev)
            capi.gemm(a, w, out_f32=out)
            assert not torch.isnan(out[:, :N]).any()
            assert (out[:, :N] - ref).abs().max().item() < tol
            if ld > N:
                assert torch.isnan(out[:, N:]).all()  # padding columns are never written
            outs.append(out[:, :N].clone())
    finally:
        capi.check(capi.lib().afft_set_gemm_skinny(1))
    assert (outs[0] - outs[1]).abs().max().item() < tol


def test_gemm_skinny_epilogues(dev):
    """Every epilogue feature on the skinny kernel: bias, GELU (erf / tanh), ReLU, gate, residual in place, residual by
    row modulus + output row map, strided output slots, 16-bit outputs in both formats, an odd N with padded pitch."""
    g = torch.Generator().manual_seed(21)
    F = torch.nn.functional
    M, N, K = 90, 1024, 1024
    a = _randn(g, dev, M, K).bfloat16()
    w = _randn(g, dev, N, K, scale=0.05).bfloat16()
    bias, res = _randn(g, dev, N), _randn(g, dev, M, N)
    ref0 = _mm(a, w)
    out = torch.zeros(M, N, device=dev)
    capi.gemm(a, w, bias=bias, act=capi.ACT_GELU_ERF, out_f32=out)
    assert (out - F.gelu(ref0 + bias)).abs().max().item() < 1e-4
    capi.gemm(a, w, bias=bias, act=capi.ACT_GELU_TANH, out_f32=out)
    assert (out - F.gelu(ref0 + bias, approximate="tanh")).abs().max().item() < 1e-4
    capi.gemm(a, w, bias=bias, act=capi.ACT_RELU, out_f32=out)
    assert (out - torch.relu(ref0 + bias)).abs().max().item() < 1e-4
    capi.gemm(a, w, bias=bias, act=capi.ACT_GATE, res=res, out_f32=out)
    assert (out - res * torch.sigmoid(ref0 + bias)).abs().max().item() < 1e-4
    h = res.clone()
    capi.gemm(a, w, bias=bias, res=h, out_f32=h)  # in-place residual stream update
    assert (h - (ref0 + bias + res)).abs().max().item() < 1e-4
    ob = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
    of = torch.zeros(M, N, device=dev)
    capi.gemm(a, w, bias=bias, res=res, out_f32=of, out_hi=ob)
    assert torch.equal(ob, of.bfloat16())
    T = 18
    pos = _randn(g, dev, T, N)
    outm = torch.zeros(M // T * (T + 1) + T + 1, N, device=dev)
    capi.gemm(a, w, res=pos, res_mod=T, out_f32=outm, row_map=(T, T + 1, 1))
    r = torch.arange(M, device=dev)
    assert (outm[(r // T) * (T + 1) + r % T + 1] - (ref0 + pos[r % T])).abs().max().item() < 1e-4
    slots = 5
    hbuf = torch.zeros(M * slots, N, device=dev)
    capi.gemm(a, w, out_f32=hbuf.view(M, slots * N)[:, 2 * N:3 * N])
    assert (hbuf.view(M, slots, N)[:, 2] - ref0).abs().max().item() < 1e-4
    assert hbuf.view(M, slots, N)[:, [0, 1, 3, 4]].abs().max().item() == 0.0
    # fp16 operands and output, saturation to the finite range
    a16, w16 = a.half(), w.half()
    big = torch.full((N,), 1e6, device=dev)
    o16 = torch.zeros(M, N, device=dev, dtype=torch.float16)
    capi.gemm(a16, w16, bias=big, out_hi=o16)
    assert torch.isfinite(o16).all() and (o16 == 65504).all()
    capi.gemm(a16, w16, bias=bias, act=capi.ACT_GELU_ERF, out_hi=o16)
    assert (o16.float() - F.gelu(_mm(a16, w16) + bias)).abs().max().item() < 4e-3
    # classifier width (N = 3806, pitch 3808), 18 rows
    Nc = 3806
    wc = _randn(g, dev, Nc, K, scale=0.05).bfloat16()
    bc = torch.zeros(Nc + 16, device=dev)[:Nc]
    bc.copy_(_randn(g, dev, Nc))
    oc = torch.full((18, 3808), 7.0, device=dev)
    capi.gemm(a[:18], wc, bias=bc, out_f32=oc)
    assert (oc[:, :Nc] - (_mm(a[:18], wc) + bc)).abs().max().item() < 1e-4
    assert (oc[:, Nc:] == 7.0).all()
