"""Writes tests/golden/feats_reader.npz: outputs of the REFERENCE input pipeline
(``BaseVideoDataset._sample`` -> ``EpicRULSTMFeatsReader`` -> ``read_representations``; datasets/base_video_dataset.py:225-337,
datasets/reader_fns.py:65-157) on a synthetic feature store, and checks oracle/feats_oracle.py against them bit for bit.

Run in the build container only (needs /root/reference):  python tests/golden/make_feats_golden.py
Stand-ins: ``lmdb`` / ``cv2`` / ``hydra.types`` are absent here; the LMDB environment is replaced by a dict-backed
object with the two methods the reader uses (``env.begin()`` as a context manager, ``txn.get(key)``).
"""
import importlib
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import feats_oracle, ref_shim  # noqa: E402

MODS = {"rgb": 16, "objects": 8, "audio": 16, "flow": 16}   # narrow rows: the arithmetic under test is in the indices
ORIG_FPS_MODS = ("audio",)                                   # reader_fns.py:131 ('audio' or 'poses' in the lmdb path)
VIDEOS = {"P01_101": 2400, "P02_07": 900, "P03_123": 300}    # name -> number of 30-fps frames (3-digit: EK100, 2-digit: EK55)
FPS, REQ_FPS = 30.0, 4.0


class FakeTxn:
    def __init__(self, d):
        self.d = d

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def get(self, key):
        return self.d.get(key)


class FakeEnv:
    def __init__(self, d):
        self.d = d

    def begin(self):
        return FakeTxn(self.d)


def build_stores(seed=0):
    """Per modality {key: float32[C]}; ~30 % of the frames are absent, in runs up to 14 long, so that the
    closest-earlier-frame search (radius 9) both succeeds and fails."""
    rng = np.random.default_rng(seed)
    stores = {}
    for mod, C in MODS.items():
        d = {}
        for v, n30 in VIDEOS.items():
            n = n30 if mod not in ORIG_FPS_MODS else int(n30 / FPS * feats_oracle.orig_video_fps(v)) + 2
            present = np.ones(n + 1, dtype=bool)
            i = 1
            while i <= n:
                if rng.random() < 0.08:
                    run = int(rng.integers(1, 15))
                    present[i:i + run] = False
                    i += run
                i += 1
            for f in range(1, n + 1):
                if present[f]:
                    d[f"{v}_frame_{f:010d}.jpg"] = rng.standard_normal(C).astype(np.float32)
        stores[mod] = d
    return stores


def clip_cases(seed=1):
    rng = np.random.default_rng(seed)
    cases = []
    for v, n30 in VIDEOS.items():
        dur = n30 / FPS
        for _ in range(14):
            T = int(rng.choice([10, 16, 18]))
            end = float(rng.uniform(0.3, dur))
            start = end - T / REQ_FPS + float(rng.choice([0.0, 0.0, 0.013, -0.2, 1.1]))
            cases.append((v, start, end, T, "last_clip"))
        cases.append((v, -3.0, 1.0, 18, "last_clip"))          # window clipped at 0: front padding
        cases.append((v, 0.0, 0.2, 10, "last_clip"))            # fewer frames than one step
        cases.append((v, 2.0, 2.0 + 10 / REQ_FPS, 10, "center_clip"))
        cases.append((v, 1.0, 9.0, 10, "first_clip"))
    for t in feats_oracle.nextafter_cases():                # ends exactly on / next to a frame boundary
        cases.append(("P01_101", t - 4.5, t, 18, "last_clip"))
    return [c for c in cases if c[2] > 0.05]


def main():
    ref_shim.install_stubs()
    for name in ("lmdb", "cv2"):
        sys.modules.setdefault(name, types.ModuleType(name))
    ht = types.ModuleType("hydra.types")
    ht.TargetConf = dict
    sys.modules.setdefault("hydra.types", ht)
    sys.modules["hydra"].types = ht
    sys.modules["hydra"].__path__ = []  # make the stub look like a package
    rf = importlib.import_module("datasets.reader_fns")
    try:
        bvd = importlib.import_module("datasets.base_video_dataset")
        sample = bvd.BaseVideoDataset._sample
    except Exception as e:  # noqa: BLE001
        raise SystemExit(f"cannot import the reference dataset module: {e!r}")

    stores = build_stores()
    readers = {}
    for mod in MODS:
        r = rf.EpicRULSTMFeatsReader.__new__(rf.EpicRULSTMFeatsReader)
        torch.nn.Module.__init__(r)
        r.lmdb_path = [f"/fake/{mod}_lmdb"]            # 'audio' in the path selects the orig-fps index (:131)
        r.lmdb_envs = [FakeEnv({k.encode("utf-8"): v.tobytes() for k, v in stores[mod].items()})]
        r.warn_if_using_closeby_frame = False
        readers[mod] = r

    cases = clip_cases()
    out = {"videos": np.array([c[0] for c in cases]), "start": np.array([c[1] for c in cases]),
           "end": np.array([c[2] for c in cases]), "T": np.array([c[3] for c in cases]),
           "strategy": np.array([c[4] for c in cases])}
    n_checked, n_raised = 0, 0
    valid = np.ones(len(cases), dtype=bool)
    for mod in MODS:
        feats = []
        for ci, (v, s, e, T, strat) in enumerate(cases):
            try:
                video, _, _, _, _, _ = sample(None, f"/videos/{v}.MP4", FPS, s, e, None, T, REQ_FPS, strat, readers[mod], None)
            except AssertionError:
                valid[ci] = False  # no frame of the whole window stored: the reference refuses (reader_fns.py:93)
                n_raised += 1
                feats.append(np.zeros((18, MODS[mod]), np.float32))
                continue
            ref = video.reshape(T, -1).numpy()
            ids = feats_oracle.clip_frame_ids(v, s, e, FPS, T, REQ_FPS, strat, orig_fps_index=mod in ORIG_FPS_MODS)
            mine = feats_oracle.gather_clip(stores[mod], v, ids, MODS[mod])
            if not np.array_equal(mine, ref):
                raise SystemExit(f"oracle != reference for {mod} {v} [{s}, {e}] T={T} {strat}")
            n_checked += 1
            feats.append(np.concatenate([ref, np.zeros((18 - T, ref.shape[1]), np.float32)]))  # pad to a common T
        out[f"feat_{mod}"] = np.stack(feats)
    out["valid"] = valid
    print(f"{n_raised} (modality, clip) pairs refused by the reference (window without any stored frame)")
    # the store itself (keys sorted) so the tests rebuild the identical dict without the RNG
    for mod in MODS:
        keys = sorted(stores[mod])
        out[f"keys_{mod}"] = np.array(keys)
        out[f"rows_{mod}"] = np.stack([stores[mod][k] for k in keys])
    # ---- sample_strategy = random_clip: the reference draws from a numpy Generator and from Python's global `random`, once
    # per (clip, modality) call, clips in order and modalities within a clip (BaseVideoDataset._get_video :369-373).
    # Both are seeded here; the oracle / native plan re-draw with the same seeds and must select the same rows.
    import random as pyrandom
    rcases = [c for c in cases if c[4] == "last_clip"][:40]
    rcases = [(v, s - extra, e, T) for (v, s, e, T, _), extra in zip(rcases, np.random.default_rng(5).choice([0.0, 1.5, 4.0], 40))]
    rng_ref = np.random.default_rng(77)
    pyrandom.seed(77)
    rfeat = {m: [] for m in MODS}
    rvalid = np.ones(len(rcases), dtype=bool)
    for ci, (v, s, e, T) in enumerate(rcases):
        for mod in MODS:
            try:
                video, _, _, _, _, _ = sample(None, f"/videos/{v}.MP4", FPS, s, e, None, T, REQ_FPS, "random_clip", readers[mod], rng_ref)
                ref = video.reshape(T, -1).numpy()
            except AssertionError:
                rvalid[ci] = False
                ref = np.zeros((T, MODS[mod]), np.float32)
            rfeat[mod].append(np.concatenate([ref, np.zeros((18 - T, ref.shape[1]), np.float32)]))
    rng_or = np.random.default_rng(77)
    pyrandom.seed(77)
    n_rand = 0
    for ci, (v, s, e, T) in enumerate(rcases):
        for mod in MODS:
            draws = feats_oracle.random_draws(s, e, FPS, T, REQ_FPS, rng_or, pyrandom)
            if not rvalid[ci]:
                continue
            ids = feats_oracle.clip_frame_ids(v, s, e, FPS, T, REQ_FPS, "random_clip", orig_fps_index=mod in ORIG_FPS_MODS, rand=draws)
            mine = feats_oracle.gather_clip(stores[mod], v, ids, MODS[mod])
            if not np.array_equal(mine, rfeat[mod][ci][:T]):
                raise SystemExit(f"oracle != reference (random_clip) for {mod} {v} [{s}, {e}] T={T}")
            n_rand += 1
    out["r_videos"] = np.array([c[0] for c in rcases])
    out["r_start"] = np.array([c[1] for c in rcases])
    out["r_end"] = np.array([c[2] for c in rcases])
    out["r_T"] = np.array([c[3] for c in rcases])
    out["r_valid"] = rvalid
    out["r_seed"] = np.array(77)
    for mod in MODS:
        out[f"r_feat_{mod}"] = np.stack(rfeat[mod])
    print(f"random_clip: oracle == reference on {n_rand} (modality, clip) pairs ({int((~rvalid).sum())} clips refused)")
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "feats_reader.npz")
    np.savez_compressed(path, **out)
    print(f"oracle == reference on {n_checked} (modality, clip) pairs; wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB)")


if __name__ == "__main__":
    main()
