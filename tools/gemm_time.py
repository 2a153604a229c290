"""Isolated timing of one GEMM shape through the C ABI (CUDA events, 5 warm-ups + 40 timed launches, operands
larger than L2 are re-streamed every launch).  usage: python tools/gemm_time.py M N K epi[none|gelu|res|f32] [reps]
Set AFFT_B200_LIB to time another build of the library (A/B of kernel variants on the same box)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from afft_b200 import _capi as capi  # noqa: E402

M, N, K = (int(x) for x in sys.argv[1:4])
epi = sys.argv[4] if len(sys.argv) > 4 else "none"
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 40
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
a = torch.randn(M, K, generator=g).to(dev).bfloat16()
w = (torch.randn(N, K, generator=g) * 0.05).to(dev).bfloat16()
bias = torch.randn(N, generator=g).to(dev)
Np = (N + 7) // 8 * 8  # 16-byte row pitches for the epilogue's vector accesses
out_b = torch.zeros(M, Np, device=dev, dtype=torch.bfloat16)[:, :N]
out_f = torch.zeros(M, Np, device=dev)[:, :N]
bias = torch.cat([bias, torch.zeros(Np - N, device=dev)])[:N]
kw = {}
if epi == "gelu":
    kw.update(bias=bias, act=capi.ACT_GELU_ERF, out_hi=out_b)
elif epi == "res":
    kw.update(bias=bias, res=out_f, out_f32=out_f)
elif epi == "f32":
    kw.update(out_f32=out_f)
else:
    kw.update(out_hi=out_b)
for _ in range(5):
    capi.gemm(a, w, **kw)
torch.cuda.synchronize()
best = 1e9
tot = 0.0
for _ in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps // 4):
        capi.gemm(a, w, **kw)
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / (reps // 4) * 1e3
    best = min(best, t)
    tot += t
print(f"{os.path.basename(capi.LIB_PATH)} M={M} N={N} K={K} epi={epi}: avg {tot / 4:.1f} us, best {best:.1f} us, "
      f"{2.0 * M * N * K / (best * 1e-6) / 1e12:.0f} TFLOP/s (best)")
