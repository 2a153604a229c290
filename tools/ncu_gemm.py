"""Run one GEMM configuration a few times (for `ncu -k regex:gemm_bf16`).
usage: python tools/ncu_gemm.py M N K epi[none|gelu|res|resbf] [bn] [strict]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from afft_b200 import _capi as capi  # noqa: E402

M, N, K = (int(x) for x in sys.argv[1:4])
epi = sys.argv[4] if len(sys.argv) > 4 else "none"
bn = int(sys.argv[5]) if len(sys.argv) > 5 else 0
strict = len(sys.argv) > 6 and sys.argv[6] == "strict"
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
a = torch.randn(M, K, generator=g).to(dev).bfloat16()
w = (torch.randn(N, K, generator=g) * 0.05).to(dev).bfloat16()
bias = torch.randn(N, generator=g).to(dev)
out_b = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
out_f = torch.zeros(M, N, device=dev)
kw = {}
if strict:
    kw.update(a_lo=torch.zeros_like(a), w_lo=torch.zeros_like(w))
if epi == "gelu":
    kw.update(bias=bias, act=capi.ACT_GELU_ERF, out_hi=out_b)
elif epi == "res":
    kw.update(bias=bias, res=out_f, out_f32=out_f)
elif epi == "f32":
    kw.update(out_f32=out_f)
else:
    kw.update(out_hi=out_b)
for _ in range(5):
    capi.gemm(a, w, force_block_n=bn, **kw)
torch.cuda.synchronize()
print("done")
