"""Where does the host spend its time in the end-to-end loop (bench.py e2e)?  Prints ms per call of each host-side piece
(no device synchronisation inside the timed pieces) and the device time per step.  usage: python tools/e2e_host_profile.py [B]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from afft_b200 import configs  # noqa: E402
from afft_b200.models import BaseModel  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cfg, T, ncls, _ = configs.named_config("ek100_sa_tsn")
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = BaseModel(cfg, ncls, {}, max_batch=B).to(dev).eval()
order = [m for m in cfg["modal_feature_order"] if m in cfg["modal_dims"]]
KW = dict(mixup_fn=None, target=None, target_subclips=None, target_subclips_ignore_index=None)
host = [{m: torch.randn(B, T, cfg["modal_dims"][m], 1, 1, 1).pin_memory() for m in order} for _ in range(2)]
C = list(ncls.values())[0]
host_out = [torch.empty(B, C).pin_memory() for _ in range(2)]
copy_stream = torch.cuda.Stream(dev)
main = torch.cuda.current_stream(dev)
acc = {"prefetch": 0.0, "model": 0.0, "d2h": 0.0, "wait": 0.0}


def prefetch(i):
    with torch.cuda.stream(copy_stream):
        d = {m: t.to(dev, non_blocking=True) for m, t in host[i % 2].items()}
        ev = torch.cuda.Event()
        ev.record(copy_stream)
    return d, ev


def loop(n, measure):
    nxt = prefetch(0)
    done = [None, None]
    for i in range(n):
        d, ev = nxt
        main.wait_event(ev)
        t0 = time.perf_counter()
        if i + 1 < n:
            nxt = prefetch(i + 1)
        t1 = time.perf_counter()
        with torch.no_grad():
            o, _ = model(d, **KW)
        t2 = time.perf_counter()
        for t in d.values():
            t.record_stream(main)
        host_out[i % 2].copy_(o["logits/action"]["all-fused"][:, 0, :], non_blocking=True)
        done[i % 2] = torch.cuda.Event()
        done[i % 2].record(main)
        t3 = time.perf_counter()
        if i > 0:
            done[(i - 1) % 2].synchronize()
        t4 = time.perf_counter()
        if measure:
            acc["prefetch"] += t1 - t0
            acc["model"] += t2 - t1
            acc["d2h"] += t3 - t2
            acc["wait"] += t4 - t3
    done[(n - 1) % 2].synchronize()


loop(5, False)
torch.cuda.synchronize()
N = 40
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
t0 = time.perf_counter()
loop(N, True)
e1.record()
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / N * 1e3
print(f"device ms/step {e0.elapsed_time(e1) / N:.3f}  wall ms/step {wall:.3f}")
print("host ms per step: " + ", ".join(f"{k} {v / N * 1e3:.3f}" for k, v in acc.items()))
# H2D alone
torch.cuda.synchronize()
e0.record(copy_stream)
for i in range(10):
    with torch.cuda.stream(copy_stream):
        d = {m: t.to(dev, non_blocking=True) for m, t in host[i % 2].items()}
e1.record(copy_stream)
torch.cuda.synchronize()
byts = sum(t.numel() * 4 for t in host[0].values())
print(f"H2D alone: {e0.elapsed_time(e1) / 10:.3f} ms per batch of {byts / 1e6:.1f} MB = {byts / (e0.elapsed_time(e1) / 10) / 1e6:.1f} GB/s")
