#!/usr/bin/env python
"""Benchmark of the AFFT fusion-and-anticipation forward path (BASELINE.json metric: SA-Fuser EK100 forward
clips/sec on B200, fraction of the bf16 tensor-core roofline, next to the host-CPU path).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU cores

A "step" is one forward of the hot path over one batch of synthetic clips per GPU (weak scaling: the
per-GPU batch is fixed, clips are sharded data-parallel, no data-path collective).  One JSON line is printed
by rank 0.  See DESIGN.md section "Measurement" for how each field is obtained.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from afft_b200 import configs, synthetic  # noqa: E402
from afft_b200 import dist as adist  # noqa: E402

METRIC = "SA-Fuser EK100 forward clips/sec"
UNIT = "clips/s"
KW = dict(mixup_fn=None, target=None, target_subclips=None, target_subclips_ignore_index=None)


_REAL_STDOUT = None


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version line to fd 1 when
    NCCL_DEBUG=VERSION), so everything but the result is sent to stderr: fd 1 is pointed at fd 2 for the whole run and
    the result goes to the saved descriptor."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"burst": p["bf16_tflops"], "sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                "hbm_gbs": p["hbm_gbs"], "source": "measured (MEASURED_PEAKS.json)"}
    return {"burst": 1590.0, "sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU through NVML while a timed region runs."""

    def __init__(self, index: int, period_s: float = 0.01):
        self.index, self.period = index, period_s
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            # CUDA_VISIBLE_DEVICES remapping: resolve by UUID of the torch device
            uuid = str(torch.cuda.get_device_properties(index).uuid)
            h = None
            for cand in (f"GPU-{uuid}", uuid):
                try:
                    h = pynvml.nvmlDeviceGetHandleByUUID(cand.encode() if isinstance(cand, str) else cand)
                    break
                except Exception:  # noqa: BLE001
                    continue
            self._h = h if h is not None else pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self._nvml = None

    def _loop(self):
        nv = self._nvml
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
                 "hw_power_brake": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80)}
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                mask = get_reasons(self._h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(self.period)

    def __enter__(self):
        if self._nvml is not None:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm (oracle port) on the host cores
# ------------------------------------------------------------------------------------------------
def _cpu_name():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def cpu_forward_timer(cfg, T, ncls, batch, steps, warmup, budget_s=240.0, chunk=0):
    """Times oracle.forward (the CPU restatement of the reference path, pinned to the reference module; issued as the
    ATen library calls the module makes) on all host cores.  One step = `batch` clips, processed in chunks of `chunk`
    clips (0: the whole batch at once).  Returns (clips/s, cores, description, steps, seconds)."""
    from afft_b200.models import BaseModel
    from oracle import afft_oracle  # bench.py's cpu_baseline / --impl reference legs may execute the oracle
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    afft_oracle.ATEN_OPS = True  # issue the same ATen library calls as the reference module (fair CPU timing)
    torch.manual_seed(0)
    model = BaseModel(cfg, ncls, {})  # random-init weights of the architecture (CPU tensors; no native call)
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    feats = synthetic.synthetic_features(cfg["modal_dims"], batch, T, seed=1000)
    chunk = chunk if 0 < chunk < batch else batch

    def one_step():
        for b0 in range(0, batch, chunk):
            afft_oracle.forward(sd, cfg, ncls, {m: f[b0:b0 + chunk] for m, f in feats.items()})

    with torch.no_grad():
        t0 = time.perf_counter()
        one_step()
        first = time.perf_counter() - t0
        if first * (steps + warmup) > budget_s:  # keep the whole run bounded
            steps = max(1, int(budget_s / first) - warmup)
        for _ in range(max(0, warmup - 1)):
            one_step()
        t0 = time.perf_counter()
        for _ in range(steps):
            one_step()
        dt = time.perf_counter() - t0
    desc = (f"{steps} steps of {batch} clips" + (f" in chunks of {chunk}" if chunk != batch else "") +
            f", fp32 torch CPU ops (oracle port issuing the reference module's ATen calls), {cores} threads on {_cpu_name()}")
    return batch * steps / dt, cores, desc, steps, dt


def cpu_best_of(cfg, T, ncls, batches=(8, 32, 128), budget_s=45.0):
    """cpu_baseline: best clips/s over the batch sizes SURVEY section 8d names (1 warm-up + up to 3 timed forwards each)."""
    best, table = None, {}
    for b in batches:
        v, cores, desc, _, _ = cpu_forward_timer(cfg, T, ncls, b, 3, 1, budget_s=budget_s / len(batches))
        table[str(b)] = round(v, 2)
        if best is None or v > best[0]:
            best = (v, cores, desc, b)
    return best, table


def staged_e2e(args, cfg, T, B, model, dev, main_stream, sampler, n_gpus, rank, C, location="pinned"):
    """The same metric measured from CLIP DESCRIPTORS: (video, start_sec, end_sec) -> native plan (host integers) ->
    gather kernel reading the pinned host row tables over PCIe -> BaseModel -> logits to the host.  Replaces the
    reference's LMDB reader + collate + .to(device) (datasets/reader_fns.py:65-138, test.py:81)."""
    import numpy as np
    from afft_b200 import staging
    dims = cfg["modal_dims"]
    n_videos, n_frames = 12, 1500
    g = torch.Generator().manual_seed(4000 + rank)
    store = staging.FeatureStore(dims, orig_fps_mods=("audio",))
    names = [f"P{v + 1:02d}_{101 + v}" for v in range(n_videos)]
    for m, width in dims.items():
        n = n_frames if m != "audio" else int(n_frames / 30.0 * 50.0) + 2
        frames = np.arange(1, n + 1, dtype=np.int32)
        frames = frames[(frames % 11) != 0]  # every 11th frame absent: the closest-earlier-frame fallback runs
        store.set_modality(m, {v: (frames, torch.randn(len(frames), width, generator=g).numpy()) for v in names},
                           "pinned" if location == "pinned" else str(dev))
    rng = np.random.default_rng(rank)
    batches = []
    for _ in range(4):
        vids = rng.choice(names, size=B).tolist()
        en = rng.uniform(5.0, n_frames / 30.0, size=B)
        batches.append((vids, en - T / 4.0, en))
    stager = staging.FeatureStager(store, T, max_batch=B, device=dev, depth=2, fps=30.0, frame_rate=4.0)
    host_outs = [torch.empty(B, C).pin_memory() for _ in range(2)]

    def loop(n):
        nxt = stager.stage(*batches[0])
        done = [None, None]
        for i in range(n):
            feats, ev, slot = nxt
            main_stream.wait_event(ev)
            if i + 1 < n:
                nxt = stager.stage(*batches[(i + 1) % len(batches)])  # plan + gather of the next batch overlap this forward
            with torch.no_grad():
                o, _ = model(dict(feats), **KW)
            stager.done(slot)
            host_outs[i % 2].copy_(o["logits/action"]["all-fused"][:, 0, :], non_blocking=True)
            done[i % 2] = torch.cuda.Event()
            done[i % 2].record(main_stream)
            if i > 0:
                done[(i - 1) % 2].synchronize()
        done[(n - 1) % 2].synchronize()

    loop(max(2, args.warmup))
    adist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with sampler:
        e0.record()
        loop(args.steps)
        e1.record()
        torch.cuda.synchronize()
    adist.barrier()
    ms = adist.max_over_ranks(e0.elapsed_time(e1), dev)
    table_bytes = sum(t.numel() * 4 for t in store.rows.values())
    return {"value": round(n_gpus * B * args.steps / (ms / 1e3), 1), "unit": UNIT, "ms_per_step": round(ms / args.steps, 4),
            "plan_bytes_h2d_per_step": len(dims) * B * T * 4,
            "feature_bytes_read_over_pcie_per_step": sum(dims.values()) * B * T * 4 if location == "pinned" else 0,
            "d2h_bytes_per_step": B * C * 4,
            "store": f"{'pinned host' if location == 'pinned' else 'HBM-resident'} row tables, {table_bytes / 1e6:.0f} MB, {n_videos} videos",
            "api": "afft_b200.staging.FeatureStager.stage (afft_store_plan + afft_store_gather) -> BaseModel.__call__"}

def run_reference(args):
    """The reference arm: the reference's own CPU implementation of the path (the oracle port - /root/reference does not
    exist on the GPU box - issuing the ATen calls the reference module makes; within 4 % of the module's own speed in
    the build container) on all host cores, on the GPU arm's config: one step = the same B clips per step, processed in
    chunks of the CPU's best batch size (the reference's eval batch is 32; expts/01_SA-Fuser_ek100_val_TSN.txt:6)."""
    rank, _, world = adist.env_world()
    if rank != 0:
        return
    cfg, T, ncls, eval_bs = configs.named_config(args.config)
    B = args.batch
    chunk = args.cpu_batch
    chunk_table = None
    if not chunk:  # pick the CPU's best chunk with one short probe each
        best, chunk_table = cpu_best_of(cfg, T, ncls, batches=(eval_bs, 128), budget_s=20.0)
        chunk = best[3]
    value, cores, desc, steps, dt = cpu_forward_timer(cfg, T, ncls, B, args.steps, args.warmup, budget_s=200.0, chunk=chunk)
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": args.warmup, "ms_per_step": round(1e3 * dt / steps, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, cfg, T, B, "reference algorithm on host CPU"),
        "cpu_baseline": {"value": round(value, 3), "unit": UNIT, "cores": cores, "kind": "port", "sample": desc,
                         "chunk_clips": chunk, "chunk_probe_clips_per_s": chunk_table},
        "e2e": {"value": round(value, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_config(args, cfg, T, batch, how, ncls=None):
    return {"workload": f"{args.config}: SA-Fuser EK100 R-TSN+O+AU+F 4h_18s (expts/01_SA-Fuser_ek100_val_TSN.txt)"
            if args.config == "ek100_sa_tsn" else args.config,
            "clips_per_gpu_per_step": batch, "T": T, "modal_dims": cfg["modal_dims"],
            "gemm_gflop_per_clip": round(configs.gemm_flops_per_clip(cfg, T, ncls or configs.named_config(args.config)[2]) / 1e9, 3),
            "weights": "random init (torch.manual_seed(0))", "parallelism": f"dp{args.gpus} (clips sharded, no collective)",
            "path": how,
            "l2": "no explicit flush: per-step working set (772 MB 16-bit weights + >1 GB activations) exceeds the 126 MB L2; "
                  "4 input buffer sets are rotated"}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
PREC_DESC = {"bf16": "bf16 operands / fp32 accumulate", "fp16": "fp16 operands / fp32 accumulate",
             "strict": "strict bf16x3 (hi.hi + hi.lo + lo.hi) / fp32 accumulate"}


class DeviceRun:
    """One precision mode of the model on this rank: engine, persistent io structs, the lowest-overhead step."""

    def __init__(self, cfg, T, ncls, B, precision, dev, dev_sets, order):
        from afft_b200 import _capi
        from afft_b200.models import BaseModel
        torch.manual_seed(0)
        self.model = BaseModel(cfg, ncls, {}, precision=precision, max_batch=B).to(dev).eval()
        self.head = self.model.future_predictor
        with torch.no_grad():
            self.model(dict(dev_sets[0]), **KW)  # builds the engine, packs the weights
        torch.cuda.synchronize()
        self.eng = eng = next(iter(self.head._engines.values()))
        self.launches_per_fwd = eng.launch_count()
        D = cfg["common_dim"]
        self.C = C = list(ncls.values())[0]
        ldc = (C + 3) // 4 * 4
        n_tok, H1 = eng.n_slots, eng.fuser_heads
        if eng.fuser_kind == _capi.FUSER_TSA:
            attn_buf = torch.empty(B, eng.fuser_depth, H1, n_tok * T, n_tok * T, device=dev)
        elif eng.fuser_kind == _capi.FUSER_CA:
            attn_buf = None
        else:
            attn_buf = torch.empty(B, eng.fuser_depth, T, H1, n_tok, n_tok, device=dev)
        self.bufs = dict(orig=torch.empty(B, T, D, device=dev), pf=torch.empty(B, T + eng.fp_output_len, D, device=dev),
                         logits=torch.empty(B, T + eng.fp_output_len, ldc, device=dev), attn=attn_buf)
        self.ios = []
        for sset in dev_sets:
            io = _capi.IO()
            for i, m in enumerate(order):
                io.feat[i] = sset[m].data_ptr()
            io.orig_past, io.past_futures = self.bufs["orig"].data_ptr(), self.bufs["pf"].data_ptr()
            io.logits[0], io.ld_logits[0] = self.bufs["logits"].data_ptr(), ldc
            io.fuser_attn = attn_buf.data_ptr() if attn_buf is not None else None
            self.ios.append(io)
        self.B = B

    def step(self, i):
        self.eng.forward_into(self.ios[i % len(self.ios)], self.B)

    def timed(self, steps, warmup, dev, sampler=None):
        """K steps between a barrier + synchronize on both sides, CUDA events, max over ranks -> total ms."""
        for i in range(warmup):
            self.step(i)
        torch.cuda.synchronize()
        adist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ctx = sampler if sampler is not None else _Null()
        with ctx:
            torch.cuda.synchronize()
            e0.record()
            for i in range(steps):
                self.step(i)
            e1.record()
            torch.cuda.synchronize()
        adist.barrier()
        return adist.max_over_ranks(e0.elapsed_time(e1), dev)

    def profile(self, psteps=5):
        """Per-launch device time slices of `psteps` steady-state forwards (in-kernel exit timestamps; afft_profile_enable)."""
        self.eng.profile_enable(True)
        agg = {0: [0.0, 0], 1: [0.0, 0], 2: [0.0, 0], 3: [0.0, 0]}
        by_shape, gemm_flops, last = {}, 0.0, None
        for i in range(psteps):
            for j in range(3):  # back to back: the profiled (last) forward starts while its predecessor drains - steady state
                self.step(3 * i + j)
            torch.cuda.synchronize()
            last = self.eng.profile_read()
            for cat, M, N, K, ms in last:
                agg[cat][0] += ms
                agg[cat][1] += 1
                if cat == 0:
                    gemm_flops += 2.0 * M * N * K
                    sh = by_shape.setdefault((M, N, K), [0.0, 0])
                    sh[0] += ms
                    sh[1] += 1
        self.eng.profile_enable(False)
        return agg, by_shape, gemm_flops, last


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def parity_block(runs, cfg, T, ncls, dev, n_clips=1024, chunk=128):
    """>= 1024 seeded clips (seed 123) through every precision mode and through the fp32 CPU oracle:
    max |dlogit| and ordered top-5 identity, stratified by the oracle's 5th-6th logit gap (afft_b200/parity.py)."""
    from afft_b200 import parity
    from oracle import afft_oracle  # checker
    any_run = next(iter(runs.values()))
    sd = {k: v.detach().cpu() for k, v in any_run.model.state_dict().items()}
    feats = synthetic.synthetic_features(cfg["modal_dims"], n_clips, T, seed=123)
    torch.set_num_threads(os.cpu_count() or 1)
    afft_oracle.ATEN_OPS = True
    t0 = time.perf_counter()
    refs = []
    with torch.no_grad():
        for b0 in range(0, n_clips, chunk):
            r = afft_oracle.forward(sd, cfg, ncls, {m: f[b0:b0 + chunk] for m, f in feats.items()}, dtype=torch.float32)
            refs.append(r["logits/action"]["all-fused"][:, 0])
    ref = torch.cat(refs)
    oracle_s = time.perf_counter() - t0
    out = {"clips": n_clips, "reference": "oracle/afft_oracle.py fp32 on the host CPU (pinned to the reference module, tests/golden)",
           "inputs": "synthetic randn features, seed 123; weights torch.manual_seed(0) random init",
           "oracle_seconds": round(oracle_s, 1), "modes": {}}
    for prec, run in runs.items():
        got = []
        for b0 in range(0, n_clips, run.B):
            with torch.no_grad():
                o, _ = run.model({m: f[b0:b0 + run.B].reshape(-1, T, f.shape[-1], 1, 1, 1).to(dev) for m, f in feats.items()}, **KW)
            got.append(o["logits/action"]["all-fused"][:, 0].float().cpu())
        out["modes"][prec] = parity.top5_stats(torch.cat(got), ref)
    return out


def batch_sweep_block(run, flops_per_clip, peaks, batches=(1, 8, 32, 64, 128)):
    """BASELINE config 2 ("batch sweep"): the same engine at smaller batches (32 is the reference's shipped eval batch,
    expts/01_SA-Fuser_ek100_val_TSN.txt:6), plain stream launches and CUDA-graph replay."""
    out = {}
    io = run.ios[0]

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

    for b in batches:
        if b > run.B:
            continue
        ms_plain = timeit(lambda: run.eng.forward_into(io, b), 50)
        ms_graph = None
        try:
            g = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                run.eng.forward_into(io, b)
            torch.cuda.current_stream().wait_stream(side)
            with torch.cuda.graph(g):
                run.eng.forward_into(io, b)
            ms_graph = timeit(g.replay, 50)
        except Exception as exc:  # noqa: BLE001 - the plain number stands on its own
            print(f"[bench] graph capture at B={b} failed: {exc!r}", file=sys.stderr)
            torch.cuda.synchronize()
        best = min(ms_plain, ms_graph) if ms_graph is not None else ms_plain
        out[str(b)] = {"ms_plain": round(ms_plain, 4), "ms_graph_replay": round(ms_graph, 4) if ms_graph is not None else None,
                       "clips_per_s": round(b / best * 1e3, 1),
                       "whole_step_frac": round(b * flops_per_clip / best / 1e9 / peaks["sustained"], 4)}
    return out


def gpu_baseline_block(B, T, dev, steps=10):
    """Stock eager PyTorch on the same GPU (tools/stock_torch_gpu.py: the library calls the reference module makes -
    F.linear / F.layer_norm / F.gelu / softmax attention; shares no code with afft_b200 or oracle/)."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import stock_torch_gpu as st
    W = st.make_weights(dev)
    feats = [torch.randn(B, T, d, device=dev) for d in (1024, 352, 1024, 1024)]
    res = {"what": "stock eager PyTorch (F.linear / F.layer_norm / F.gelu / softmax attention), same GPU, same batch; "
                   "tools/stock_torch_gpu.py", "unit": UNIT, "steps": steps}
    old = torch.backends.cuda.matmul.allow_tf32
    for mode in ("tf32", "bf16_autocast"):
        torch.backends.cuda.matmul.allow_tf32 = mode == "tf32"
        ctx = torch.autocast("cuda", dtype=torch.bfloat16) if mode == "bf16_autocast" else torch.autocast("cuda", enabled=False)
        with torch.no_grad(), ctx:
            for _ in range(3):
                st.forward(W, feats, T=T)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                st.forward(W, feats, T=T)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        res[mode] = {"ms_per_step": round(ms, 3), "value": round(B / ms * 1e3, 1)}
    torch.backends.cuda.matmul.allow_tf32 = old
    del W, feats
    torch.cuda.empty_cache()
    return res


def run_afft(args):
    from afft_b200 import _capi

    rank, local_rank, world = adist.init()
    if world != args.gpus and rank == 0:
        print(f"[bench] note: WORLD_SIZE={world} but --gpus={args.gpus}; using WORLD_SIZE", file=sys.stderr)
    n_gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200); the hot path has no CPU fallback. "
                         "Use --impl reference for the CPU arm.")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    _capi.lib()  # fail loudly if the extension is not built
    cfg, T, ncls, eval_bs = configs.named_config(args.config)
    B = args.batch
    peaks = load_peaks()
    flops_per_clip = configs.gemm_flops_per_clip(cfg, T, ncls)
    order = [m for m in cfg["modal_feature_order"] if m in cfg["modal_dims"]]

    # ---- device-resident inputs: NBUF rotating sets ----
    NBUF = 4
    g = torch.Generator(device=dev).manual_seed(1000 + rank)
    dev_sets = [{m: torch.randn(B, T, cfg["modal_dims"][m], 1, 1, 1, device=dev, generator=g) for m in order}
                for _ in range(NBUF)]
    run = DeviceRun(cfg, T, ncls, B, args.precision, dev, dev_sets, order)
    model, head, eng, C = run.model, run.head, run.eng, run.C

    # ---- value: device-resident inputs, exactly K steps, CUDA events, max over ranks ----
    sampler = ClockSampler(local_rank)
    ms_total = run.timed(args.steps, args.warmup, dev, sampler)
    ms_per_step = ms_total / args.steps
    value = n_gpus * B * args.steps / (ms_total / 1e3)

    # ---- sustained: the same loop repeated for >= 2 s (the power cap settles after ~1 s of dense tensor work) ----
    sus_steps = max(args.steps, int(args.sustain_s * 1e3 / ms_per_step) + 1)
    sus_sampler = ClockSampler(local_rank)
    sus_ms = run.timed(sus_steps, 0, dev, sus_sampler)
    sustained = {"value": round(n_gpus * B * sus_steps / (sus_ms / 1e3), 1), "unit": UNIT, "steps": sus_steps,
                 "seconds": round(sus_ms / 1e3, 3), "ms_per_step": round(sus_ms / sus_steps, 4), "clocks": sus_sampler.summary()}

    # ---- e2e: public API, pinned host inputs, H2D + forward + D2H of the consumed logits every step ----
    host_sets = [{m: torch.randn(B, T, cfg["modal_dims"][m], 1, 1, 1).pin_memory() for m in order} for _ in range(2)]
    h2d_bytes = sum(t.numel() * 4 for t in host_sets[0].values())
    d2h_bytes = B * C * 4
    host_outs = [torch.empty(B, C).pin_memory() for _ in range(2)]
    copy_stream = torch.cuda.Stream(dev)
    main_stream = torch.cuda.current_stream(dev)
    head.return_attentions = True

    def prefetch(i):
        with torch.cuda.stream(copy_stream):
            d = {m: t.to(dev, non_blocking=True) for m, t in host_sets[i % 2].items()}
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return d, ev

    def e2e_loop(n):
        # Two batches in flight: while batch i computes, the host consumes the logits of batch i-1 (its D2H was enqueued
        # right behind its forward) and the copy stream uploads batch i+1.  Every batch's inputs cross H2D and every
        # batch's logits cross D2H inside the timed region; the host waits for each result exactly once.
        nxt = prefetch(0)
        done = [None, None]
        checksum = 0.0
        for i in range(n):
            d, ev = nxt
            main_stream.wait_event(ev)
            if i + 1 < n:
                nxt = prefetch(i + 1)  # overlaps the next batch's H2D with this batch's compute
            with torch.no_grad():
                o, _ = model(d, **KW)  # the call test.py:82 makes
            for t in d.values():
                t.record_stream(main_stream)
            host_outs[i % 2].copy_(o["logits/action"]["all-fused"][:, 0, :], non_blocking=True)  # test.py:86
            done[i % 2] = torch.cuda.Event()
            done[i % 2].record(main_stream)
            if i > 0:
                done[(i - 1) % 2].synchronize()
                checksum += float(host_outs[(i - 1) % 2][0, 0])  # the host reads the previous batch's result
        done[(n - 1) % 2].synchronize()
        checksum += float(host_outs[(n - 1) % 2][0, 0])
        return checksum

    e2e_loop(max(2, args.warmup))
    # Three timed passes of K steps each; the MEDIAN pass is reported and all three are listed.  The host <-> device
    # copies share PCIe and host memory with whatever else runs on the box: single passes were seen 25 % off
    # (8.5 instead of 6.4 ms per step) with the device-resident `value` of the same run unchanged.
    e2e_passes = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        adist.barrier()
        e0.record()
        e2e_loop(args.steps)
        e1.record()
        torch.cuda.synchronize()
        adist.barrier()
        e2e_passes.append(adist.max_over_ranks(e0.elapsed_time(e1), dev))
    e2e_ms = sorted(e2e_passes)[1]
    e2e_value = n_gpus * B * args.steps / (e2e_ms / 1e3)

    # ---- e2e from clip descriptors (row N4): native plan + device gather from a pinned host feature store ----
    staged = None
    if not args.no_staged:
        st_sampler = ClockSampler(local_rank)
        staged = {"pinned_store": staged_e2e(args, cfg, T, B, model, dev, main_stream, st_sampler, n_gpus, rank, C, "pinned"),
                  "hbm_store": staged_e2e(args, cfg, T, B, model, dev, main_stream, st_sampler, n_gpus, rank, C, "hbm")}

    # ---- roofline of the dominant kernel (the tcgen05 GEMM): in-kernel exit timestamps, PDL overlap preserved ----
    PSTEPS = 5
    agg, by_shape, gemm_flops, last = run.profile(PSTEPS)
    if args.verbose and rank == 0:
        names = {0: "gemm", 1: "ln", 2: "attn", 3: "other"}
        print("[launches] " + " ".join(f"{names[c]}:{ms * 1e3:.0f}" for c, _, _, _, ms in last), file=sys.stderr)
    traffic = None
    for tname in ("r02_gemm_traffic.json", "r01_gemm_traffic.json"):
        tpath = os.path.join(ROOT, "profiles", tname)
        if os.path.exists(tpath) and args.config == "ek100_sa_tsn" and B == 256 and args.precision != "strict":
            with open(tpath) as f:  # dram read+write bytes per GEMM launch from the committed ncu capture of this workload
                traffic = round(json.load(f)["traffic_bytes_per_launch_avg"])
            traffic_src = tname
            break
    # The slices partition the device time of the PROFILED forwards; they are normalised to the un-profiled K-step time so
    # that the categories sum to ms_per_step exactly (raw slice sum reported beside as kernel_ms_per_step_profiled).
    kernel_ms_total = sum(v[0] for v in agg.values())
    norm = (ms_per_step * PSTEPS / kernel_ms_total) if kernel_ms_total > 0 else 1.0
    gemm_ms = agg[0][0] * norm
    achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    step_tflops = value / n_gpus * flops_per_clip / 1e12
    sus_tflops = sustained["value"] / n_gpus * flops_per_clip / 1e12

    line = {
        "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": {"bf16": "bf16", "fp16": "fp16", "strict": "bf16x3"}[args.precision], "data": "synthetic",
        "config": workload_config(args, cfg, T, B, "afft_forward (C ABI), " + PREC_DESC[args.precision]),
        "sustained": sustained,
        "e2e": {"value": round(e2e_value, 1), "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": d2h_bytes, "ms_per_step": round(e2e_ms / args.steps, 4),
                "passes_ms_per_step": [round(p / args.steps, 4) for p in e2e_passes], "reported": "median of 3 passes of K steps",
                "api": "afft_b200.models.BaseModel.__call__ (test.py:72-86 pattern), pinned host inputs, double-buffered H2D, logits of batch i-1 read on the host while batch i computes"},
        "e2e_from_clip_descriptors": staged,
        "gpu_launches": run.launches_per_fwd * args.steps,
        "launches_per_step": run.launches_per_fwd,
        "clocks": sampler.summary(),
        "roofline": {
            "bound": "tensor", "kernel": "gemm_bf16_tcgen05_2cta_kernel / gemm_bf16_tcgen05_kernel", "achieved": round(achieved, 1),
            "peak": peaks["sustained"], "unit": "TFLOP/s", "frac": round(achieved / peaks["sustained"], 4),
            "peak_kind": "bf16_tflops_sustained, " + peaks["source"], "frac_of_burst": round(achieved / peaks["burst"], 4),
            "traffic": traffic,
            "traffic_note": (f"STATIC: dram bytes per GEMM launch, avg over the GEMM launches of a forward, from the committed ncu "
                             f"capture profiles/{traffic_src} (not measured in this run)") if traffic is not None else None,
            "how": f"{PSTEPS} profiled forwards, each the last of 3 issued back to back; every kernel records its exit %globaltimer, "
                   "launch i owns (end of i-1, end of i]: the slices sum to the forward's device time and PDL overlap is preserved "
                   "(afft_profile_enable)",
            "algorithmic_flop_per_launch_avg": round(gemm_flops / max(1, agg[0][1])),
            "launches_per_step": agg[0][1] // PSTEPS, "gemm_ms_per_step": round(gemm_ms / PSTEPS, 4),
            "kernel_ms_per_step_profiled": round(kernel_ms_total / PSTEPS, 4),
            "gemm_share_of_kernel_time": round(agg[0][0] / kernel_ms_total, 4) if kernel_ms_total else None,
            "whole_step_tflops": round(step_tflops, 1), "whole_step_frac": round(step_tflops / peaks["sustained"], 4),
            "whole_step_frac_of_burst": round(step_tflops / peaks["burst"], 4),
            "sustained_whole_step_tflops": round(sus_tflops, 1), "sustained_whole_step_frac": round(sus_tflops / peaks["sustained"], 4),
            "other_kernels_ms_per_step": {"layernorm": round(agg[1][0] * norm / PSTEPS, 4), "attention": round(agg[2][0] * norm / PSTEPS, 4),
                                          "assembly_convert": round(agg[3][0] * norm / PSTEPS, 4)},
        },
    }
    if args.verbose and rank == 0:
        for (M, N, K), (ms, n) in sorted(by_shape.items(), key=lambda kv: -kv[1][0]):
            print(f"[gemm] M={M} N={N} K={K} launches/step={n // PSTEPS} ms/launch={ms / n:.4f} "
                  f"TFLOP/s={2.0 * M * N * K / (ms / n) / 1e9:.0f}", file=sys.stderr)

    # ---- the other precision modes on the same box, same inputs (sub-records) ----
    runs = {args.precision: run}
    modes = {args.precision: {"value": round(value, 1), "ms_per_step": round(ms_per_step, 4),
                              "whole_step_frac": round(step_tflops / peaks["sustained"], 4), "headline": True}}
    if not args.no_modes:
        for prec in ("bf16", "fp16", "strict"):
            if prec in runs:
                continue
            r2 = DeviceRun(cfg, T, ncls, B, prec, dev, dev_sets, order)
            ms2 = r2.timed(args.steps, args.warmup, dev)
            v2 = n_gpus * B * args.steps / (ms2 / 1e3)
            sus2_steps = max(args.steps, int(min(args.sustain_s, 1.0) * 1e3 / (ms2 / args.steps)) + 1)
            sus2 = r2.timed(sus2_steps, 0, dev)
            a2, _, gf2, _ = r2.profile(3)
            modes[prec] = {"value": round(v2, 1), "ms_per_step": round(ms2 / args.steps, 4),
                           "whole_step_frac": round(v2 / n_gpus * flops_per_clip / 1e12 / peaks["sustained"], 4),
                           "sustained_value": round(n_gpus * B * sus2_steps / (sus2 / 1e3), 1),
                           "gemm_frac": round(gf2 / (a2[0][0] * 1e-3) / 1e12 / peaks["sustained"], 4) if a2[0][0] > 0 else None,
                           "path": PREC_DESC[prec]}
            runs[prec] = r2
    modes[args.precision].update(sustained_value=sustained["value"], gemm_frac=round(achieved / peaks["sustained"], 4),
                                 path=PREC_DESC[args.precision])
    line["modes"] = modes

    if rank == 0 and n_gpus == 1 and not args.no_cpu_baseline:
        # parity of exactly this build on >= 1024 clips for every mode, the stock-PyTorch GPU bar and the CPU baseline
        line["parity"] = parity_block(runs, cfg, T, ncls, dev, n_clips=args.parity_clips)
        ps = line["parity"]["modes"][args.precision]
        line["parity_sample"] = {"clips": ps["clips"], "max_abs_dlogit_vs_oracle_fp32": ps["max_abs_dlogit"],
                                 "top5_identical_clips": ps["ordered_top5_identical"], "mode": args.precision}
        line["batch_sweep"] = batch_sweep_block(run, flops_per_clip, peaks)
        if args.config == "ek100_sa_tsn":
            line["gpu_baseline"] = gpu_baseline_block(B, T, dev)
        best, table = cpu_best_of(cfg, T, ncls, batches=(8, eval_bs, 128))
        line["cpu_baseline"] = {"value": round(best[0], 3), "unit": UNIT, "cores": best[1], "kind": "port", "sample": best[2],
                                "clips_per_s_by_batch": table,
                                "note": "oracle port issuing the reference module's ATen calls (the reference module itself "
                                        "cannot travel to the GPU box); best of the batch sizes listed"}
    if rank == 0:
        emit(line)
    adist.shutdown()


# ------------------------------------------------------------------------------------------------
# training-step arm (BASELINE config 5): fwd + bwd + NCCL gradient all-reduce (DDP) + SGD-nesterov step
# ------------------------------------------------------------------------------------------------
def run_train(args):
    from afft_b200 import _capi
    from afft_b200 import train as atrain
    from afft_b200.models import BaseModel

    rank, local_rank, world = adist.init()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --mode train needs CUDA devices (B200)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    _capi.lib()
    train_cfg_name = args.config if args.config != "ek100_sa_tsn" else "ek100_sa_swin"
    cfg, T, ncls, _ = configs.named_config(train_cfg_name)
    B = args.batch if args.batch != 256 else 16  # expts/01_SA-Fuser_ek100_train.txt:7: 16 clips per GPU
    C = list(ncls.values())[0]
    torch.manual_seed(0)
    model = BaseModel(cfg, ncls, {}).to(dev).train()
    # Gradients live in ONE flat buffer laid out in backward-completion order (afft_b200.dist.GradBuckets); for N > 1 each
    # layer group's slice is all-reduced over NCCL as soon as backward has finished it, overlapping the remaining dgrad /
    # wgrad work, eagerly or inside the captured graph.  --grad-comm bf16 halves the bytes on NVLink.  --ddp keeps
    # torch's DistributedDataParallel (eager only) for comparison.
    import torch.distributed as tdist
    use_ddp = world > 1 and args.ddp
    ddp = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local_rank]) if use_ddp else model
    buckets = state = None
    if not use_ddp:
        if world > 1:
            for p in model.parameters():
                tdist.broadcast(p.data, src=0)  # what DDP's constructor does
        # flat parameter / gradient / momentum / bf16-operand buffers; native wgrad writes into the gradient views, the
        # optimizer is one kernel that also emits the next step's bf16 weights (afft_b200.train.TrainState)
        state = atrain.TrainState(model.future_predictor, lr=1e-3, momentum=0.9, weight_decay=1e-6, nesterov=True,
                                  comm_dtype=torch.bfloat16 if args.grad_comm == "bf16" else torch.float32,
                                  n_buckets=args.grad_buckets)
        buckets = state.buckets
        state.__enter__()
    # expts/01 :48-52; DDP arm: torch's fused SGD
    opt = torch.optim.SGD(model.parameters(), lr=1e-3, momentum=0.9, nesterov=True, weight_decay=1e-6, fused=True) if use_ddp else None
    g = torch.Generator(device=dev).manual_seed(1000 + rank)
    sets = [{m: torch.randn(B, T, d, 1, 1, 1, device=dev, generator=g) for m, d in cfg["modal_dims"].items()} for _ in range(2)]
    cls_name = list(ncls.keys())[0]
    target = {cls_name: torch.randint(0, C, (B, 1), device=dev, generator=g)}
    target_sub = {cls_name: torch.randint(0, C, (B, T), device=dev, generator=g)}
    # the experiment trains with MixUp on the backbone outputs and smoothed soft labels (expts/01_SA-Fuser_ek100_train.txt:10-12,
    # conf/config.yaml:16-22); afft_b200.runner restates common/mixup.py + common/runner.py with static shapes (capturable)
    from afft_b200 import runner as arunner
    mixup_fn = None
    if not args.no_mixup:
        mixup_fn = arunner.MixUp(alpha=0.1, label_smoothing={"action": 0.4, "verb": 0.01, "noun": 0.03}, num_classes=dict(ncls),
                                 device_lambda=True)

    def fwd_bwd_opt(feats):
        if state is not None:
            state.zero()
        else:
            opt.zero_grad(set_to_none=True)
        loss, _, _ = arunner.training_losses(ddp, feats, target, target_sub, mixup_fn=mixup_fn, mixup_backbone=True)
        loss.backward()  # per-group NCCL all-reduces are issued from inside backward (GradBuckets / DDP)
        if state is not None:
            state.finish()
            state.step()
        else:
            opt.step()
        return loss

    def step(i):
        return fwd_bwd_opt(dict(sets[i % 2]))

    for i in range(max(3, args.warmup)):
        step(i)
    torch.cuda.synchronize()

    # One GPU: the whole step (forward, backward, SGD) is captured into ONE CUDA graph - at 16 clips per GPU the step is
    # ~1900 launches of small kernels and Python-launch-bound.  (With DDP the NCCL bucket hooks stay eager.)
    graphed = False
    if not args.no_graph and not use_ddp:
        try:
            static = {m: torch.empty_like(t) for m, t in sets[0].items()}
            eager_step = step
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for i in range(3):  # warm-up on the capture pool's side stream (PyTorch whole-network capture recipe)
                    for m in static:
                        static[m].copy_(sets[i % 2][m])
                    fwd_bwd_opt(dict(static))
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_loss = fwd_bwd_opt(dict(static))

            def step(i):  # noqa: F811
                for m in static:
                    static[m].copy_(sets[i % 2][m])
                graph.replay()
                return static_loss

            for i in range(3):
                step(i)
            torch.cuda.synchronize()
            graphed = True
        except Exception as exc:  # noqa: BLE001 - capture is an optimisation; the eager step is the same arithmetic
            print(f"[bench] CUDA-graph capture of the training step failed ({exc!r}); running eagerly", file=sys.stderr)
            torch.cuda.synchronize()
            step = eager_step
    adist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local_rank)
    with sampler:
        e0.record()
        for i in range(args.steps):
            loss = step(i)
        e1.record()
        torch.cuda.synchronize()
    adist.barrier()
    ms_total = adist.max_over_ranks(e0.elapsed_time(e1), dev)
    n_params = sum(p.numel() for p in model.parameters())
    flops_per_clip = 3.0 * configs.gemm_flops_per_clip(cfg, T, ncls)
    value = world * B * args.steps / (ms_total / 1e3)
    line = {
        "metric": "SA-Fuser EK100 training step clips/sec (fwd + bwd + gradient all-reduce + SGD)", "value": round(value, 1),
        "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": round(ms_total / args.steps, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic", "mode": "train",
        "config": {"workload": f"{train_cfg_name} training step" + (" (expts/01_SA-Fuser_ek100_train.txt: T=16, 4 modalities, SGD-nesterov)"
                                                                    if train_cfg_name == "ek100_sa_swin" else " (SGD-nesterov)"),
                   "clips_per_gpu_per_step": B, "T": T, "params": n_params,
                   "mixup": ("MixUp(alpha 0.1) on the backbone outputs + label smoothing 0.4, soft-label losses, acc1 / acc5 per step "
                             "(afft_b200.runner = common/mixup.py + common/runner.py)") if mixup_fn is not None else "off (--no-mixup)", "grad_allreduce_bytes": 4 * n_params if world > 1 else 0,
                   "allreduce": ((f"{len(buckets.groups)} per-layer-group NCCL all-reduces ({args.grad_comm} transport) issued from inside "
                                  f"backward (GradBuckets), " + ("captured in the step's CUDA graph" if graphed else "eager"))
                                 if (world > 1 and buckets is not None) else
                                 "torch DDP bucketed NCCL all-reduce overlapped with backward (eager)" if world > 1 else "none (1 GPU)"),
                   "grad_allreduce_bytes_on_wire": buckets.bytes_per_step() if buckets is not None else 4 * n_params * (world > 1),
                   "gemm_gflop_per_clip_fwd_bwd": round(flops_per_clip / 1e9, 2), "final_loss": round(float(loss.detach()), 4)},
        "achieved_tflops": round(value * flops_per_clip / 1e12, 1),
        "cuda_graph": graphed,
        "clocks": sampler.summary(),
    }
    if rank == 0:
        emit(line)
    if graphed and world > 1:
        # Tearing the NCCL communicator down while a graph with captured collectives is alive hung at exit (observed
        # once, N = 2): release the graph, meet the other ranks, and leave without destroy_process_group.
        torch.cuda.synchronize()
        adist.barrier()
        try:
            graph.reset()
        except Exception:  # noqa: BLE001
            pass
        sys.stderr.flush()
        os._exit(0)
    adist.shutdown()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", choices=["infer", "train"], default="infer")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["afft", "reference"], default="afft")
    ap.add_argument("--config", default="ek100_sa_tsn", choices=configs.CONFIG_NAMES)
    ap.add_argument("--batch", type=int, default=256, help="clips per GPU per step")
    ap.add_argument("--cpu-batch", type=int, default=0, help="clips per CPU forward (default: the experiment's eval batch)")
    ap.add_argument("--precision", choices=["bf16", "fp16", "strict"], default="fp16",
                    help="GEMM operand arithmetic of the headline numbers: fp16 (default: the mode that meets the north star's "
                         "roofline AND top-5 clauses together - same kernels and tensor rate as bf16, 8x less operand rounding, "
                         "1-2 %% slower), bf16 (the library's default: widest range) or strict (bf16x3); the other two modes "
                         "are reported in `modes`")
    ap.add_argument("--strict", action="store_true", help="alias of --precision strict")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="train mode: do not capture the step into a CUDA graph")
    ap.add_argument("--no-mixup", action="store_true", help="train mode: hard labels, no MixUp (the experiment uses MixUp)")
    ap.add_argument("--grad-comm", choices=["fp32", "bf16"], default="fp32", help="train mode: gradient all-reduce transport dtype")
    ap.add_argument("--grad-buckets", type=int, default=4,
                    help="train mode: merge the layer groups into this many all-reduce buckets (0: one per group); 4 measured best at N = 8")
    ap.add_argument("--ddp", action="store_true", help="train mode, N > 1: torch DistributedDataParallel (eager) instead of GradBuckets")
    ap.add_argument("--no-staged", action="store_true", help="skip the clip-descriptor (feature store) e2e measurement")
    ap.add_argument("--no-modes", action="store_true", help="skip the sub-records of the other precision modes")
    ap.add_argument("--sustain-s", type=float, default=2.0, help="length of the sustained timing loop in seconds")
    ap.add_argument("--parity-clips", type=int, default=1024, help="clips of the in-run parity block (rank 0, N=1)")
    ap.add_argument("--verbose", action="store_true")
    args = ap.parse_args()
    if args.strict:
        args.precision = "strict"
    args.strict = args.precision == "strict"
    claim_stdout()
    if args.mode == "train":
        run_train(args)
    elif args.impl == "reference":
        args.warmup = max(args.warmup, 1)
        run_reference(args)
    else:
        args.warmup = max(args.warmup, 3)  # timing rule: at least 3 warm-up steps
        run_afft(args)


if __name__ == "__main__":
    main()
