"""Input side: the native batch plan + device gather against the reference reader.

CPU (`-m "not gpu"`): oracle/feats_oracle.py == golden fixture written from the reference classes
(tests/golden/make_feats_golden.py); the native plan (afft_store_plan, host arithmetic only - no compute kernel) returns
exactly the rows the oracle's frame ids select, on the fixture clips and on fresh random clips, including windows
clipped at 0, frame-boundary end points, missing-frame runs longer than the search radius, and error cases.
GPU: the gather kernel reproduces the reference's clip tensors bit for bit from a pinned-host and from an HBM
store, and a staged batch drives BaseModel to the same logits as the same features uploaded by hand.
"""
import os
import re
import subprocess

import numpy as np
import pytest
import torch

from afft_b200 import _capi, staging
from oracle import feats_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FPS, REQ_FPS = 30.0, 4.0
MODS = ["rgb", "objects", "audio", "flow"]


@pytest.fixture(scope="module")
def golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "feats_reader.npz"))
    stores = {m: {k: r for k, r in zip(z[f"keys_{m}"].tolist(), z[f"rows_{m}"])} for m in MODS}
    return z, stores


def _native_store(stores, location):
    dims = {m: len(next(iter(d.values()))) for m, d in stores.items()}
    envs = {m: ((k.encode(), v.tobytes()) for k, v in d.items()) for m, d in stores.items()}
    return staging.FeatureStore.from_key_value(dims, envs, orig_fps_mods=("audio",), location=location), dims


def _rows_from_plan(store, idx, dims):
    out = {}
    for mi, m in enumerate(store.mods):
        table = store.rows[m].cpu().numpy()
        sel = idx[mi].numpy()
        feat = np.zeros(sel.shape + (dims[m],), np.float32)
        feat[sel >= 0] = table[sel[sel >= 0]]
        out[m] = feat
    return out


def test_staging_header_symbols_exported():
    src = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "afft_staging.h")).read(), flags=re.S)
    declared = sorted(set(re.findall(r"\b(afft_[a-z0-9_]+)\s*\(", src)))
    assert declared == sorted(staging.STAGING_SYMBOLS)
    out = subprocess.run(["nm", "-D", "--defined-only", _capi.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert set(declared) <= exported


def test_oracle_matches_reference_fixture(golden):
    z, stores = golden
    n = 0
    for ci, (v, s, e, T, strat) in enumerate(zip(z["videos"].tolist(), z["start"], z["end"], z["T"], z["strategy"].tolist())):
        if not z["valid"][ci]:
            continue
        for m in MODS:
            ids = feats_oracle.clip_frame_ids(v, float(s), float(e), FPS, int(T), REQ_FPS, strat, orig_fps_index=(m == "audio"))
            mine = feats_oracle.gather_clip(stores[m], v, ids, stores[m][next(iter(stores[m]))].shape[0])
            assert np.array_equal(mine, z[f"feat_{m}"][ci][:T]), (m, v, s, e)
            n += 1
    assert n > 250


def test_native_plan_matches_fixture_and_oracle(golden):
    z, stores = golden
    store, dims = _native_store(stores, "host")
    vids, st, en, Ts, strats = z["videos"].tolist(), z["start"], z["end"], z["T"], z["strategy"].tolist()
    for T in sorted(set(Ts.tolist())):
        for strat in sorted(set(strats)):
            every = [i for i in range(len(vids)) if Ts[i] == T and strats[i] == strat]
            # clips the reference reader asserts on (no stored frame anywhere in the window, reader_fns.py:97) are refused
            # by the native plan as well; everything else is planned as one batch
            for i in every:
                if not z["valid"][i]:
                    with pytest.raises(_capi.AfftError, match="no stored frame|no frame id"):
                        store.plan([vids[i]], st[[i]], en[[i]], FPS, int(T), REQ_FPS, strat)
            sel = [i for i in every if z["valid"][i]]
            if not sel:
                continue
            idx, fids = store.plan([vids[i] for i in sel], st[sel], en[sel], FPS, int(T), REQ_FPS, strat, want_frame_ids=True)
            feats = _rows_from_plan(store, idx, dims)
            for bi, ci in enumerate(sel):
                for mi, m in enumerate(MODS):
                    ids = feats_oracle.clip_frame_ids(vids[ci], float(st[ci]), float(en[ci]), FPS, int(T), REQ_FPS, strat,
                                                      orig_fps_index=(m == "audio"))
                    assert np.array_equal(fids[mi, bi].numpy(), ids), (m, vids[ci], st[ci], en[ci])
                    if z["valid"][ci]:
                        assert np.array_equal(feats[m][bi], z[f"feat_{m}"][ci][:T]), (m, vids[ci], st[ci], en[ci])


def test_native_plan_random_clips_large_batch(golden):
    """Fresh clips (not in the fixture), B = 300 so that the threaded path of the planner runs."""
    _, stores = golden
    store, dims = _native_store(stores, "host")
    rng = np.random.default_rng(7)
    vids = rng.choice(["P01_101", "P02_07", "P03_123"], size=300).tolist()
    dur = {"P01_101": 80.0, "P02_07": 30.0, "P03_123": 10.0}
    en = np.array([rng.uniform(0.2, dur[v]) for v in vids])
    st = en - rng.uniform(0.1, 6.0, size=300)
    for T, frame_rate, strat in ((18, 4.0, "last_clip"), (10, 4.0, "center_clip"), (16, None, "first_clip"), (18, 2.5, "last_clip")):
        idx = store.plan(vids, st, en, FPS, T, frame_rate, strat)
        feats = _rows_from_plan(store, idx, dims)
        ref = feats_oracle.read_batch(stores, ("audio",), vids, st.tolist(), en.tolist(), FPS, T, frame_rate, strat, widths=dims)
        for m in MODS:
            assert np.array_equal(feats[m], ref[m]), (m, T, strat)


def test_plan_errors(golden):
    _, stores = golden
    store, _ = _native_store(stores, "host")
    with pytest.raises(_capi.AfftError, match="not in modality"):
        store.plan(["P99_999"], [0.0], [4.0], FPS, 18, REQ_FPS)
    with pytest.raises(_capi.AfftError, match="no frame id"):
        store.plan(["P01_101"], [0.0], [0.01], FPS, 18, REQ_FPS)  # window ends before frame 1 (reader_fns.py:122)
    with pytest.raises(_capi.AfftError):
        store.plan(["P01_101"], [0.0], [4.0], FPS, 0, REQ_FPS)
    with pytest.raises(_capi.AfftError, match="CPU path"):
        staging.FeatureStager(store, 18, 4, device="cpu")
    bad = staging.FeatureStore({"audio": 16}, orig_fps_mods=("audio",))
    bad.set_modality("audio", {"weird_name_1234": (np.array([1, 2]), np.zeros((2, 16), np.float32))}, location="host")
    with pytest.raises(_capi.AfftError, match="Unkown video name format"):  # reader_fns.py:157 (the reference's spelling)
        bad.plan(["weird_name_1234"], [0.0], [4.0], FPS, 18, REQ_FPS)
    with pytest.raises(_capi.AfftError, match="ascending"):
        bad.set_modality("audio", {"P01_101": (np.array([2, 2]), np.zeros((2, 16), np.float32))}, location="host")


@pytest.mark.gpu
@pytest.mark.parametrize("location", ["pinned", "cuda:0"])
def test_gather_matches_reference_fixture(golden, location):
    z, stores = golden
    store, dims = _native_store(stores, location)
    vids, st, en, Ts, strats = z["videos"].tolist(), z["start"], z["end"], z["T"], z["strategy"].tolist()
    sel = [i for i in range(len(vids)) if Ts[i] == 18 and strats[i] == "last_clip"]
    stager = staging.FeatureStager(store, 18, max_batch=len(sel), device="cuda:0", fps=FPS, frame_rate=REQ_FPS)
    for rep in range(3):  # slots are reused
        feats, ev, slot = stager.stage([vids[i] for i in sel], st[sel], en[sel])
        ev.synchronize()
        for m in MODS:
            got = feats[m].reshape(len(sel), 18, dims[m]).cpu().numpy()
            for bi, ci in enumerate(sel):
                if z["valid"][ci]:
                    assert np.array_equal(got[bi], z[f"feat_{m}"][ci]), (m, ci, rep)
        stager.done(slot)


@pytest.mark.gpu
def test_staged_batch_drives_the_model():
    from afft_b200 import configs, synthetic
    from afft_b200.models import BaseModel
    cfg, T, ncls, _ = configs.named_config("egtea_sa")
    dims = cfg["modal_dims"]
    rng = np.random.default_rng(3)
    vids = {f"P0{i}_10{i}": 400 + 50 * i for i in range(1, 4)}
    store = staging.FeatureStore(dims, orig_fps_mods=())
    raw = {}
    for m, C in dims.items():
        per = {}
        for v, n in vids.items():
            frames = np.sort(rng.choice(np.arange(1, n + 1), size=int(n * 0.9), replace=False)).astype(np.int32)
            per[v] = (frames, rng.standard_normal((len(frames), C)).astype(np.float32))
        raw[m] = {f"{v}_frame_{f:010d}.jpg": r for v, (fr, rows) in per.items() for f, r in zip(fr, rows)}
        store.set_modality(m, per, "pinned")
    B = 6
    names = rng.choice(list(vids), size=B).tolist()
    en = np.array([rng.uniform(3.0, vids[v] / FPS) for v in names])
    st = en - T / REQ_FPS
    stager = staging.FeatureStager(store, T, max_batch=8, fps=FPS, frame_rate=REQ_FPS)
    feats, ev, slot = stager.stage(names, st, en)
    torch.cuda.current_stream().wait_event(ev)
    model = BaseModel(cfg, ncls, {})
    model.load_state_dict(synthetic.synthetic_state_dict(model, seed=0))
    model = model.to("cuda:0").eval()
    kw = dict(mixup_fn=None, target=None, target_subclips=None, target_subclips_ignore_index=None)
    with torch.no_grad():
        out, _ = model(dict(feats), **kw)
    stager.done(slot)
    ref = feats_oracle.read_batch(raw, (), names, st.tolist(), en.tolist(), FPS, T, REQ_FPS, widths=dims)
    for m in dims:
        assert np.array_equal(feats[m].reshape(B, T, dims[m]).cpu().numpy(), ref[m]), m
    with torch.no_grad():
        out2, _ = model({m: torch.from_numpy(ref[m]).reshape(B, T, dims[m], 1, 1, 1).cuda() for m in dims}, **kw)
    assert torch.equal(out["logits/action"]["all-fused"], out2["logits/action"]["all-fused"])


def test_plan_shards_over_ranks_like_the_forward(golden):
    """Multi-GPU: clips are independent units; each rank plans (and gathers) its contiguous shard.  The per-rank plans
    concatenate to the single-process plan (no collective on the data path)."""
    from afft_b200 import dist as adist
    z, stores = golden
    store, dims = _native_store(stores, "host")
    rng = np.random.default_rng(11)
    vids = rng.choice(["P01_101", "P02_07", "P03_123"], size=37).tolist()
    en = np.array([rng.uniform(5.0, 10.0) for _ in vids])
    st = en - 4.5
    full = store.plan(vids, st, en, FPS, 18, REQ_FPS)
    for world in (2, 3, 8):
        parts = []
        for r in range(world):
            lo, hi = adist.shard_bounds(len(vids), r, world)
            parts.append(store.plan(vids[lo:hi], st[lo:hi], en[lo:hi], FPS, 18, REQ_FPS))
        assert torch.equal(torch.cat(parts, dim=1), full)


def test_native_frame_ids_match_oracle_over_rates_and_windows(golden):
    """The window / subsampling / padding / original-fps arithmetic at other frame rates than the 30 -> 4 fps of the
    shipped configs (float64 floor / int / round-half-even edge cases): frame ids bit-identical to the oracle."""
    _, stores = golden
    store, _ = _native_store(stores, "host")
    rng = np.random.default_rng(2024)
    vids = rng.choice(["P01_101", "P02_07", "P03_123"], size=64).tolist()
    n_checked = 0
    for fps in (30.0, 29.97, 25.0, 50.0, 59.94005994005994, 24.0):
        for frame_rate in (4.0, 2.5, 5.0, 7.5, None):
            for T, strat in ((18, "last_clip"), (10, "center_clip"), (16, "first_clip"), (1, "last_clip")):
                en = rng.uniform(0.5, 60.0, size=64)
                st = en - rng.choice([T / 4.0, 0.37, 3.0, 11.1], size=64) + rng.choice([0.0, 1e-9, -1e-9, 0.004], size=64)
                # frame-boundary end points: k / fps and its float neighbours
                en[:8] = [k / fps for k in (1, 2, 7, 30, 31, 135, 299, 1200)]
                en[8:16] = np.nextafter(en[:8], 0.0)
                en[16:24] = np.nextafter(en[:8], 1e9)
                keep = []
                for i in range(64):
                    try:
                        ids = [feats_oracle.clip_frame_ids(vids[i], float(st[i]), float(en[i]), fps, T, frame_rate, strat,
                                                            orig_fps_index=(m == "audio")) for m in MODS]
                    except (AssertionError, ValueError):
                        continue  # empty window / no frame id >= 1: the native planner refuses those (covered elsewhere)
                    keep.append((i, ids))
                sel = [i for i, _ in keep]
                _, fids = store.plan([vids[i] for i in sel], st[sel], en[sel], fps, T, frame_rate, strat, want_frame_ids=True,
                                     allow_empty=True)  # windows beyond the stored frames: only the arithmetic is checked
                for bi, (i, ids) in enumerate(keep):
                    for mi in range(len(MODS)):
                        assert np.array_equal(fids[mi, bi].numpy(), ids[mi]), (fps, frame_rate, T, strat, vids[i], st[i], en[i], MODS[mi])
                        n_checked += 1
    assert n_checked > 20000


def test_from_lmdb_uses_the_reference_open_mode_and_path_rule(golden, monkeypatch):
    """FeatureStore.from_lmdb against a stand-in `lmdb` module (the package is not in this image): environments are
    opened readonly / lock=False like reader_fns.py:52, and 'audio' in the path selects the original-fps index."""
    import sys
    import types
    _, stores = golden
    opened = []

    class Txn:
        def __init__(self, d):
            self.d = d

        def cursor(self):
            return iter((k.encode(), v.tobytes()) for k, v in self.d.items())

    class Env:
        def __init__(self, d):
            self.d, self.closed = d, False

        def begin(self):
            return Txn(self.d)

        def close(self):
            self.closed = True

    def fake_open(path, readonly=False, lock=True):
        assert readonly and not lock
        env = Env(stores[path.split("/")[-1].split("_")[0]])
        opened.append(env)
        return env

    fake = types.ModuleType("lmdb")
    fake.open = fake_open
    monkeypatch.setitem(sys.modules, "lmdb", fake)
    dims = {m: len(next(iter(d.values()))) for m, d in stores.items()}
    store = staging.FeatureStore.from_lmdb(dims, {m: f"/data/{m}_lmdb" for m in MODS}, location="host")
    assert store.orig_fps_mods == ("audio",) and all(e.closed for e in opened) and len(opened) == len(MODS)
    ref_store, _ = _native_store(stores, "host")
    args = (["P01_101", "P02_07"], [3.0, 10.0], [7.5, 14.5], FPS, 18, REQ_FPS)
    assert torch.equal(store.plan(*args), ref_store.plan(*args))
    for m in MODS:
        assert torch.equal(store.rows[m], ref_store.rows[m])


def test_random_clip_matches_reference_fixture(golden):
    """sample_strategy='random_clip' (base_video_dataset.py:245-248,282-287): with the numpy Generator and Python's `random`
    seeded as in the fixture run of the reference classes, the native plan selects exactly the rows the reference read -
    clip by clip, one pair of draws per (clip, modality) in the reference's order."""
    import random as pyrandom
    z, stores = golden
    store, dims = _native_store(stores, "host")
    assert store.mods == MODS  # the draws are consumed in modality order
    vids, st, en, Ts = z["r_videos"].tolist(), z["r_start"], z["r_end"], z["r_T"]
    assert bool(z["r_valid"].all())
    seed = int(z["r_seed"])
    # (a) the oracle, clip by clip
    rng = np.random.default_rng(seed)
    pyrandom.seed(seed)
    for ci, v in enumerate(vids):
        for m in MODS:
            d = feats_oracle.random_draws(float(st[ci]), float(en[ci]), FPS, int(Ts[ci]), REQ_FPS, rng, pyrandom)
            ids = feats_oracle.clip_frame_ids(v, float(st[ci]), float(en[ci]), FPS, int(Ts[ci]), REQ_FPS, "random_clip",
                                              orig_fps_index=(m == "audio"), rand=d)
            mine = feats_oracle.gather_clip(stores[m], v, ids, dims[m])
            assert np.array_equal(mine, z[f"r_feat_{m}"][ci][:int(Ts[ci])]), (m, ci)
    # (b) the native plan: batches of one clip in order, so that the draws interleave as in the reference's loop
    rng = np.random.default_rng(seed)
    pyrandom.seed(seed)
    moved = 0
    for ci, v in enumerate(vids):
        T = int(Ts[ci])
        idx, fids = store.plan([v], st[[ci]], en[[ci]], FPS, T, REQ_FPS, "random_clip", want_frame_ids=True, rng=rng)
        feats = _rows_from_plan(store, idx, dims)
        last = store.plan([v], st[[ci]], en[[ci]], FPS, T, REQ_FPS, "last_clip", want_frame_ids=True)[1]
        moved += int(not np.array_equal(fids.numpy(), last.numpy()))
        for m in MODS:
            assert np.array_equal(feats[m][0], z[f"r_feat_{m}"][ci][:T]), (m, ci)
    assert moved > 20  # the jitter really moves the windows
    # (c) a whole batch at once consumes the generators in the same order (clip-major, modality-minor)
    rng = np.random.default_rng(seed)
    pyrandom.seed(seed)
    same_T = [i for i in range(len(vids)) if Ts[i] == 18][:12]
    rng_b = np.random.default_rng(3)
    pyrandom.seed(3)
    idx_b = store.plan([vids[i] for i in same_T], st[same_T], en[same_T], FPS, 18, REQ_FPS, "random_clip", rng=rng_b)
    rng_c = np.random.default_rng(3)
    pyrandom.seed(3)
    for bi, i in enumerate(same_T):
        idx_1 = store.plan([vids[i]], st[[i]], en[[i]], FPS, 18, REQ_FPS, "random_clip", rng=rng_c)
        assert np.array_equal(idx_b[:, bi].numpy(), idx_1[:, 0].numpy())
    with pytest.raises(ValueError):
        store.plan([vids[0]], st[[0]], en[[0]], FPS, 18, REQ_FPS, "random_clip")  # no generator given
