"""Batch sweep of the forward path (BASELINE config 2: 'inference on 1 B200, batch sweep'):
ms per forward and clips/s for plain stream launches and for CUDA-graph replay.
usage: python tools/batch_sweep.py [config] [bf16|fp16|strict] [b1,b2,...]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from afft_b200 import _capi, configs  # noqa: E402
from afft_b200.models import BaseModel  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "ek100_sa_tsn_wo_audio"
precision = sys.argv[2] if len(sys.argv) > 2 else "bf16"
BATCHES = tuple(int(b) for b in sys.argv[3].split(",")) if len(sys.argv) > 3 else (1, 8, 32, 64, 128, 256, 512, 1024)
cfg, T, ncls, _ = configs.named_config(name)
flops = configs.gemm_flops_per_clip(cfg, T, ncls)
dev = torch.device("cuda:0")
torch.manual_seed(0)
BMAX = 1024
model = BaseModel(cfg, ncls, {}, precision=precision, max_batch=BMAX).to(dev).eval()
order = [m for m in cfg["modal_feature_order"] if m in cfg["modal_dims"]]
kw = dict(mixup_fn=None, target=None, target_subclips=None, target_subclips_ignore_index=None)
feats = {m: torch.randn(BMAX, T, cfg["modal_dims"][m], 1, 1, 1, device=dev) for m in order}
with torch.no_grad():
    model({m: t[:2] for m, t in feats.items()}, **kw)
eng = next(iter(model.future_predictor._engines.values()))
D, C = cfg["common_dim"], list(ncls.values())[0]
ldc = (C + 3) // 4 * 4
n_tok, H1 = eng.n_slots, eng.fuser_heads
orig = torch.empty(BMAX, T, D, device=dev)
pf = torch.empty(BMAX, T + eng.fp_output_len, D, device=dev)
logits = torch.empty(BMAX, T + eng.fp_output_len, ldc, device=dev)
attn = torch.empty(BMAX, eng.fuser_depth, T, H1, n_tok, n_tok, device=dev)
io = _capi.IO()
for i, m in enumerate(order):
    io.feat[i] = feats[m].data_ptr()
io.orig_past, io.past_futures = orig.data_ptr(), pf.data_ptr()
io.logits[0], io.ld_logits[0] = logits.data_ptr(), ldc
io.fuser_attn = attn.data_ptr()


def timeit(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


rows = []
for B, ks in [(b, k) for b in BATCHES for k in ((1, 4) if b <= 128 else (4,))]:
    eng.set_max_ksplit(ks)  # 1 = split-K off (for comparison); 4 = library default
    iters = 50 if B <= 128 else 20
    ms_plain = timeit(lambda: eng.forward_into(io, B), iters)
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        eng.forward_into(io, B)
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g):
        eng.forward_into(io, B)
    ms_graph = timeit(g.replay, iters)
    rec = {"config": name, "B": B, "max_ksplit": ks, "ms_plain": round(ms_plain, 4), "ms_graph": round(ms_graph, 4),
           "clips_per_s_plain": round(B / ms_plain * 1e3, 1), "clips_per_s_graph": round(B / ms_graph * 1e3, 1),
           "tflops_graph": round(B * flops / ms_graph / 1e9, 1), "launches": eng.launch_count()}
    rows.append(rec)
    print(json.dumps(rec), flush=True)
