"""CPU: the N>1 host logic (contiguous clip sharding, barrier, max-over-ranks timing) with world_size 2 on gloo."""
import os
import socket
import subprocess
import sys

from afft_b200 import dist as adist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json
sys.path.insert(0, os.environ["AFFT_ROOT"])
import torch
from afft_b200 import dist as adist
rank, local_rank, world = adist.init(backend="gloo")
assert world == 2
lo, hi = adist.shard_bounds(37, rank, world)
adist.barrier()
t = adist.max_over_ranks(1.0 + rank, device="cpu")          # slowest rank wins
n = adist.sum_over_ranks(hi - lo, device="cpu")              # all clips covered exactly once
# the sharded "forward": each rank processes its contiguous slice; results must tile the batch
x = torch.arange(37, dtype=torch.float64)
part = (x[lo:hi] * 2).sum().item()
total = adist.sum_over_ranks(part, device="cpu")
print(json.dumps({"rank": rank, "lo": lo, "hi": hi, "tmax": t, "n": n, "total": total}), flush=True)
adist.shutdown()
'''


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 7, 32, 37, 256):
        for world in (1, 2, 3, 8):
            spans = [adist.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_world_size_2_gloo(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, AFFT_ROOT=ROOT, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r), LOCAL_RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for r in range(2)]
    outs = []
    for p in procs:
        o, e = p.communicate(timeout=180)
        assert p.returncode == 0, e[-2000:]
        outs.append(o)
    import json
    recs = sorted((json.loads(o.strip().splitlines()[-1]) for o in outs), key=lambda r: r["rank"])
    assert (recs[0]["lo"], recs[0]["hi"], recs[1]["lo"], recs[1]["hi"]) == (0, 19, 19, 37)
    assert recs[0]["tmax"] == recs[1]["tmax"] == 2.0
    assert recs[0]["n"] == 37 and recs[0]["total"] == float(sum(range(37)) * 2)


BUCKET_WORKER = r'''
import os, sys, json
sys.path.insert(0, os.environ["AFFT_ROOT"])
import torch
from afft_b200 import dist as adist
rank, local_rank, world = adist.init(backend="gloo")
torch.manual_seed(0)
net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4))
unused = torch.nn.Parameter(torch.ones(3))                       # never reached by backward: reduced in finish()
groups = [list(net[4].parameters()), list(net[2].parameters()), list(net[0].parameters()) + [unused]]  # backward order
res = {}
for name, dt in (("fp32", torch.float32), ("bf16", torch.bfloat16)):
    gb = adist.GradBuckets(groups, comm_dtype=dt)
    order = []
    orig = gb._launch
    gb._launch = lambda gi, orig=orig: (order.append(gi), orig(gi))[1]
    for step in range(2):                                        # second step: counters re-armed by zero()
        gb.zero()
        x = torch.randn(5, 8, generator=torch.Generator().manual_seed(100 + rank + 10 * step))
        net(x).pow(2).sum().backward()
        launched_in_backward = list(order)
        gb.finish()
        mine = torch.cat([p.grad.flatten() for g in groups for p in g if p is not unused]).clone()  # views into gb.flat
        mine_unused = unused.grad.clone()
        order.clear()
    # reference: the same gradients computed for both ranks' inputs on one process, averaged
    refs = []
    for r in range(world):
        for p in net.parameters():
            p.grad = None
        x = torch.randn(5, 8, generator=torch.Generator().manual_seed(100 + r + 10 * 1))
        net(x).pow(2).sum().backward()
        refs.append(torch.cat([p.grad.flatten() for g in groups for p in g if p is not unused]))
    ref = sum(refs) / world
    n = ref.numel()
    assert all(p.grad.data_ptr() % 16 == 0 for g in groups for p in g)  # slices are 16-byte aligned (TMA operands)
    res[name] = {"err": float((mine - ref).abs().max()), "scale": float(ref.abs().max()), "unused": float(mine_unused.abs().max()),
                 "launched_in_backward": launched_in_backward, "bytes": gb.bytes_per_step()}
    gb.remove_hooks()
print(json.dumps({"rank": rank, **res}), flush=True)
adist.shutdown()
'''


def test_grad_buckets_overlap_order_and_average_world_size_2_gloo(tmp_path):
    """Training step, N > 1: the bucketed gradient all-reduce issued from inside backward (last layers first), fp32 and
    bf16 transport, against gradients computed for both ranks' batches in one process."""
    import json
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    script = tmp_path / "bucket_worker.py"
    script.write_text(BUCKET_WORKER)
    env = dict(os.environ, AFFT_ROOT=ROOT, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r), LOCAL_RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for r in range(2)]
    outs = []
    for p in procs:
        o, e = p.communicate(timeout=180)
        assert p.returncode == 0, e[-2000:]
        outs.append(json.loads(o.strip().splitlines()[-1]))
    for rec in outs:
        assert rec["fp32"]["err"] < 1e-6 * max(1.0, rec["fp32"]["scale"])
        assert rec["bf16"]["err"] < 1e-2 * max(1.0, rec["bf16"]["scale"])  # bf16 transport: 2^-9 relative
        assert rec["fp32"]["unused"] == 0.0
        # groups 0 and 1 (last layers) were launched while backward was still running; group 2 holds an unused
        # parameter, so it is reduced by finish()
        assert rec["fp32"]["launched_in_backward"] == [0, 1]
        assert rec["fp32"]["bytes"] == 2 * rec["bf16"]["bytes"] > 0
