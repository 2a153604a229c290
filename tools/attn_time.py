"""Isolated timing of the attention kernels through the C ABI (CUDA events).  usage: python tools/attn_time.py
Set AFFT_B200_LIB to time another build (A/B on the same box)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from afft_b200 import _capi as capi  # noqa: E402

dev = torch.device("cuda:0")


def timeit(fn, reps=40):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps // 4):
            fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / (reps // 4) * 1e3)
    return best



B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
g = torch.Generator().manual_seed(0)
name = os.path.basename(capi.LIB_PATH)
# SA-Fuser: 5 tokens per timestep, 4 heads x 256, B*18 sequences, probabilities written (headline config)
n_seq, L, H, hd = B * 18, 5, 4, 256
qkv = torch.randn(n_seq * L, 3 * H * hd, generator=g).to(dev).bfloat16()
out = torch.empty(n_seq * L, H * hd, device=dev, dtype=torch.bfloat16)
probs = torch.empty(n_seq, H, L, L, device=dev)
t = timeit(lambda: capi.attention(qkv, n_seq, L, H, hd, out_hi=out, probs=probs, p_outer=H * L * L, p_inner=1))
byts = qkv.numel() * 2 + out.numel() * 2 + probs.numel() * 4
print(f"{name} SA-Fuser tokens attention  n_seq={n_seq} L=5 H=4 hd=256: {t:.1f} us, {byts / t / 1e3:.0f} GB/s")
# GPT-2: 18 steps causal, 4 heads x 512, B sequences
n_seq, L, H, hd = B, 18, 4, 512
qkv = torch.randn(n_seq * L, 3 * H * hd, generator=g).to(dev).bfloat16()
out = torch.empty(n_seq * L, H * hd, device=dev, dtype=torch.bfloat16)
t = timeit(lambda: capi.attention(qkv, n_seq, L, H, hd, mask=1, out_hi=out))
byts = qkv.numel() * 2 + out.numel() * 2
print(f"{name} GPT-2 causal attention     n_seq={n_seq} L=18 H=4 hd=512: {t:.1f} us, {byts / t / 1e3:.0f} GB/s")
# T-SA-Fuser: 50 tokens block-causal, 4 heads x 256
n_seq, L, H, hd = B, 50, 4, 256
qkv = torch.randn(n_seq * L, 3 * H * hd, generator=g).to(dev).bfloat16()
out = torch.empty(n_seq * L, H * hd, device=dev, dtype=torch.bfloat16)
probs = torch.empty(n_seq, H, L, L, device=dev)
t = timeit(lambda: capi.attention(qkv, n_seq, L, H, hd, mask=2, T=10, out_hi=out, probs=probs, p_outer=H * L * L, p_inner=1))
byts = qkv.numel() * 2 + out.numel() * 2 + probs.numel() * 4
print(f"{name} T-SA block-causal attention n_seq={n_seq} L=50 H=4 hd=256: {t:.1f} us, {byts / t / 1e3:.0f} GB/s")

# LayerNorm (bf16 output, affine): fuser shape and GPT-2 shape
for rows, dim in ((B * 90, 1024), (B * 18, 2048)):
    x = torch.randn(rows, dim, generator=g).to(dev)
    gam, bet = torch.ones(dim, device=dev), torch.zeros(dim, device=dev)
    y = torch.empty(rows, dim, device=dev, dtype=torch.bfloat16)
    t = timeit(lambda: capi.layernorm(x, gam, bet, 1e-5, y_hi=y))
    print(f"{name} LayerNorm rows={rows} dim={dim} (AFFT_LN_THREADS={os.environ.get('AFFT_LN_THREADS', '256')}): {t:.1f} us, "
          f"{rows * dim * 6 / t / 1e3:.0f} GB/s")
