"""Input-side microbenchmark (row N4): plan + gather throughput of afft_b200.staging against the reference-style
Python reader (oracle/feats_oracle.py, one process), on a synthetic store with the headline config's row widths.
usage: python tools/stager_bench.py [B] [n_videos] [frames_per_video]
Prints one JSON line per store location."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from afft_b200 import configs, staging  # noqa: E402


def synthetic_store(dims, n_videos, n_frames, location, seed=0):
    g = torch.Generator().manual_seed(seed)
    store = staging.FeatureStore(dims, orig_fps_mods=("audio",))
    names = [f"P{v // 100 + 1:02d}_{100 + v % 100}" for v in range(n_videos)]
    for m, C in dims.items():
        n = n_frames if m != "audio" else int(n_frames / 30.0 * 50.0) + 2
        per = {}
        for v in names:
            frames = np.arange(1, n + 1, dtype=np.int32)
            frames = frames[(frames % 11) != 0]  # every 11th frame absent: the earlier-frame fallback is exercised
            per[v] = (frames, torch.randn(len(frames), C, generator=g).numpy())
        store.set_modality(m, per, location)
    return store, names


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    n_videos = int(sys.argv[2]) if len(sys.argv) > 2 else 24
    n_frames = int(sys.argv[3]) if len(sys.argv) > 3 else 3000
    cfg, T, ncls, _ = configs.named_config("ek100_sa_tsn")
    dims = cfg["modal_dims"]
    rng = np.random.default_rng(1)
    bytes_per_clip = sum(dims.values()) * T * 4
    for location in ("pinned", "cuda:0"):
        store, names = synthetic_store(dims, n_videos, n_frames, location)
        stager = staging.FeatureStager(store, T, max_batch=B, depth=2)
        vids = rng.choice(names, size=B).tolist()
        en = rng.uniform(5.0, n_frames / 30.0, size=B)
        st = en - T / 4.0
        # host plan alone
        t0 = time.perf_counter()
        for _ in range(20):
            store.plan(vids, st, en, 30.0, T, 4.0, "last_clip")
        plan_ms = (time.perf_counter() - t0) / 20 * 1e3
        # full stage() calls back to back (plan + plan upload + gather), device time of the gather by events
        for _ in range(3):
            f, ev, slot = stager.stage(vids, st, en)
            stager.done(slot)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 30
        t0 = time.perf_counter()
        e0.record(stager.stream)
        for _ in range(iters):
            f, ev, slot = stager.stage(vids, st, en)
            stager.done(slot, stager.stream)
        e1.record(stager.stream)
        torch.cuda.synchronize()
        wall_ms = (time.perf_counter() - t0) / iters * 1e3
        dev_ms = e0.elapsed_time(e1) / iters
        rec = {"store": location, "B": B, "T": T, "bytes_per_clip": bytes_per_clip, "plan_ms_per_batch": round(plan_ms, 3),
               "stage_wall_ms_per_batch": round(wall_ms, 3), "stage_stream_ms_per_batch": round(dev_ms, 3),
               "clips_per_s": round(B / max(wall_ms, dev_ms) * 1e3, 1),
               "gather_GBps_algorithmic(read+write)": round(2 * B * bytes_per_clip / (dev_ms * 1e-3) / 1e9, 1)}
        print(json.dumps(rec), flush=True)
        del stager, store
        torch.cuda.empty_cache()
    # reference-style reader: per-frame dict gets + numpy stacking in Python (what a DataLoader worker executes)
    from oracle import feats_oracle
    small = {m: {} for m in dims}
    g = np.random.default_rng(0)
    for m, C in dims.items():
        n = 3000 if m != "audio" else 5002
        for v in names[:2]:
            for fr in range(1, n + 1):
                if fr % 11:
                    small[m][f"{v}_frame_{fr:010d}.jpg"] = g.standard_normal(C).astype(np.float32)
    vids2 = rng.choice(names[:2], size=32).tolist()
    en2 = rng.uniform(5.0, 100.0, size=32)
    t0 = time.perf_counter()
    feats_oracle.read_batch(small, ("audio",), vids2, (en2 - T / 4.0).tolist(), en2.tolist(), 30.0, T, 4.0, widths=dims)
    dt = time.perf_counter() - t0
    print(json.dumps({"python_reader_port": "oracle/feats_oracle.read_batch (T lookups per clip, not the reference's 135)",
                      "clips_per_s_one_process": round(32 / dt, 1)}))


if __name__ == "__main__":
    main()
