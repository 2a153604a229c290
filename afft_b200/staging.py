"""Input side of the hot path: clip descriptors -> the ``{modality: (B, T, C, 1, 1, 1)}``
feature dict ``BaseModel`` consumes, without the reference's per-frame LMDB gets, numpy stacking, collation and
``.to(device)`` (datasets/reader_fns.py:65-138, datasets/base_video_dataset.py:225-337, test.py:81).

``FeatureStore`` holds one row table per modality - pinned host memory (the GPU reads it over PCIe) or device
memory (``location='cuda'``: the whole store in HBM) - plus the per-video frame index.  ``FeatureStager`` turns a
batch of ``(video, start_sec, end_sec)`` into the model input with one host call (the plan: T row numbers per
modality and clip), one small H2D copy (the plan) and one gather kernel.  All arithmetic is native
(csrc/staging.cu, C ABI in include/afft_staging.h); this module is the binding.
"""
from __future__ import annotations

import ctypes as C
import re
from typing import Dict, Iterable, List, Mapping, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _capi

SAMPLE_STRATEGIES = {"last_clip": 0, "center_clip": 1, "first_clip": 2, "random_clip": 3}  # base_video_dataset.py:28-31
STAGING_SYMBOLS = ["afft_store_create", "afft_store_destroy", "afft_store_error", "afft_store_add_video",
                   "afft_store_set_rows", "afft_store_plan", "afft_store_plan_random", "afft_store_gather",
                   "afft_store_allow_empty_clips"]


def random_clip_draws(start_sec, end_sec, fps: float, T: int, frame_rate: Optional[float], n_mod: int, rng, pyrandom=None):
    """The random numbers of sample_strategy='random_clip' for a batch, drawn exactly as the reference draws them
    (datasets/base_video_dataset.py:236-248,283-286): clip by clip, and within a clip once per modality (``_get_video``
    calls ``_sample`` per modality, :369-373) - first ``rng.integers(bound)`` from the dataset's numpy Generator when the
    window is longer than the clip, then ``random.random()`` from Python's global generator.  Returns
    (start_frame int64 [n_mod, B], offset int32 [n_mod, B])."""
    import random as _random
    pyrandom = pyrandom if pyrandom is not None else _random
    B = len(start_sec)
    sf = np.zeros((n_mod, B), dtype=np.int64)
    off = np.zeros((n_mod, B), dtype=np.int32)
    req_fps = fps if not frame_rate else frame_rate
    frames_to_ext = int(round(T * (fps / req_fps)))
    shift = max(int(round(fps / req_fps / 3)), 1)
    for b in range(B):
        start, end = max(float(start_sec[b]), 0), max(float(end_sec[b]), 0)
        bound = max(int(fps * (end - start)) - frames_to_ext, 0)
        for m in range(n_mod):
            sf[m, b] = int(rng.integers(bound)) if bound > 0 else 0
            off[m, b] = int(round(pyrandom.random() * shift))
    return sf, off
_KEY_RE = re.compile(r"^(.*)_frame_(\d{10})\.jpg$")  # reader_fns.py:133


def _lib():
    l = _capi.lib()
    if not getattr(l, "_staging_bound", False):
        l.afft_store_create.argtypes = [C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_void_p)]
        l.afft_store_destroy.argtypes = [C.c_void_p]
        l.afft_store_destroy.restype = None
        l.afft_store_error.argtypes = [C.c_void_p]
        l.afft_store_error.restype = C.c_char_p
        l.afft_store_add_video.argtypes = [C.c_void_p, C.c_int32, C.c_char_p, C.c_void_p, C.c_int64, C.c_int64]
        l.afft_store_set_rows.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64]
        l.afft_store_plan.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_char_p), C.c_void_p, C.c_void_p, C.c_double,
                                      C.c_int32, C.c_double, C.c_int32, C.c_void_p, C.c_void_p]
        l.afft_store_gather.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.POINTER(C.c_void_p), C.c_void_p]
        l.afft_store_allow_empty_clips.argtypes = [C.c_void_p, C.c_int32]
        l.afft_store_plan_random.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_char_p), C.c_void_p, C.c_void_p, C.c_double,
                                             C.c_int32, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        l.afft_store_plan_random.restype = C.c_int
        for n in ("afft_store_create", "afft_store_add_video", "afft_store_set_rows", "afft_store_plan", "afft_store_gather",
                  "afft_store_allow_empty_clips"):
            getattr(l, n).restype = C.c_int
        l._staging_bound = True
    return l


class FeatureStore:
    """Row tables + frame index of the per-frame features of a set of videos.

    modal_dims: ``{modality: row width}`` in the model's modality naming; ``orig_fps_mods``: modalities whose frames
    are numbered in the original video's frame rate (the reference decides by 'audio' / 'poses' in the LMDB path,
    reader_fns.py:131)."""

    def __init__(self, modal_dims: Mapping[str, int], orig_fps_mods: Iterable[str] = ("audio", "poses")):
        self.lib = _lib()
        self.mods: List[str] = list(modal_dims)
        self.dims: Dict[str, int] = dict(modal_dims)
        self.orig_fps_mods = tuple(m for m in self.mods if m in set(orig_fps_mods))
        widths = (C.c_int32 * len(self.mods))(*[self.dims[m] for m in self.mods])
        flags = (C.c_int32 * len(self.mods))(*[int(m in self.orig_fps_mods) for m in self.mods])
        h = C.c_void_p()
        rc = self.lib.afft_store_create(len(self.mods), widths, flags, C.byref(h))
        if rc != 0:
            raise _capi.AfftError(self.lib.afft_store_error(None).decode())
        self.handle = h
        self.rows: Dict[str, torch.Tensor] = {}

    def close(self):
        if getattr(self, "handle", None):
            self.lib.afft_store_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != 0:
            raise _capi.AfftError(self.lib.afft_store_error(self.handle).decode())

    def set_modality(self, mod: str, videos: Mapping[str, Tuple[np.ndarray, np.ndarray]], location: str = "pinned"):
        """videos: ``{video_name: (frame_ids int ascending [n], rows float32 [n, C])}``.  ``location``: 'pinned'
        (host memory the GPU reads over PCIe), 'cuda' / 'cuda:N' (HBM), or 'host' (pageable: plan-only use on machines
        without a GPU; ``FeatureStager`` refuses it)."""
        m = self.mods.index(mod)
        total = sum(len(f) for f, _ in videos.values())
        table = torch.empty(max(total, 1), self.dims[mod], dtype=torch.float32)
        first = 0
        for name, (frames, rows) in videos.items():
            frames = np.ascontiguousarray(frames, dtype=np.int32)
            rows = np.asarray(rows, dtype=np.float32).reshape(len(frames), self.dims[mod])
            table[first:first + len(frames)] = torch.from_numpy(rows)
            self._check(self.lib.afft_store_add_video(self.handle, m, name.encode(), frames.ctypes.data, len(frames), first))
            first += len(frames)
        if location == "pinned":
            table = table.pin_memory()
        elif location.startswith("cuda"):
            table = table.to(location)
        elif location != "host":
            raise ValueError(location)
        self.rows[mod] = table
        self._check(self.lib.afft_store_set_rows(self.handle, m, table.data_ptr(), total))

    @classmethod
    def from_key_value(cls, modal_dims: Mapping[str, int], envs: Mapping[str, Iterable[Tuple[bytes, bytes]]],
                       orig_fps_mods: Iterable[str] = ("audio", "poses"), location: str = "pinned") -> "FeatureStore":
        """Build from per-modality iterables of ``(key, value)`` byte pairs with the reference's key format
        ``{video}_frame_{id:010d}.jpg`` and float32 row values - what ``lmdb.Environment.begin().cursor()`` yields for the
        RULSTM feature LMDBs (reader_fns.py:52,75,90), or ``dict.items()`` in tests."""
        store = cls(modal_dims, orig_fps_mods)
        for mod, it in envs.items():
            per_video: Dict[str, List[Tuple[int, np.ndarray]]] = {}
            for k, v in it:
                k = k.decode() if isinstance(k, (bytes, bytearray)) else k
                mt = _KEY_RE.match(k.strip())
                if mt is None:
                    continue
                row = np.frombuffer(v, "float32") if isinstance(v, (bytes, bytearray, memoryview)) else np.asarray(v, "float32")
                per_video.setdefault(mt.group(1), []).append((int(mt.group(2)), row))
            vids = {}
            for name, lst in per_video.items():
                lst.sort(key=lambda fr: fr[0])
                vids[name] = (np.array([f for f, _ in lst], dtype=np.int32), np.stack([r for _, r in lst]))
            store.set_modality(mod, vids, location)
        return store

    @classmethod
    def from_lmdb(cls, modal_dims: Mapping[str, int], lmdb_paths: Mapping[str, str],
                  orig_fps_mods: Optional[Iterable[str]] = None, location: str = "pinned") -> "FeatureStore":
        """Ingest the RULSTM feature LMDBs the reference reads (``EpicRULSTMFeatsReader(lmdb_path=...)``,
        reader_fns.py:41-53): one environment per modality, opened read-only without locking like the reference does.
        ``orig_fps_mods`` defaults to the reference's rule: 'audio' or 'poses' in the LMDB path (reader_fns.py:131).
        Needs the ``lmdb`` package (part of the reference's environment.yml); the whole store is read once."""
        try:
            import lmdb
        except ImportError as e:  # pragma: no cover - depends on the environment
            raise _capi.AfftError("FeatureStore.from_lmdb needs the `lmdb` package (reference environment.yml); "
                                  "use from_key_value() with any (key, value) iterable otherwise") from e
        if orig_fps_mods is None:
            orig_fps_mods = [m for m, pth in lmdb_paths.items() if "audio" in str(pth) or "poses" in str(pth)]
        envs, its = [], {}
        for m, pth in lmdb_paths.items():
            env = lmdb.open(str(pth), readonly=True, lock=False)
            envs.append(env)
            txn = env.begin()
            its[m] = txn.cursor()
        try:
            return cls.from_key_value(modal_dims, its, orig_fps_mods=orig_fps_mods, location=location)
        finally:
            for env in envs:
                env.close()

    def plan(self, video_names: Sequence[str], start_sec: Sequence[float], end_sec: Sequence[float], fps: float, T: int,
             frame_rate: Optional[float], strategy: str = "last_clip", out: Optional[torch.Tensor] = None,
             want_frame_ids: bool = False, allow_empty: bool = False, rng=None, pyrandom=None):
        """Row numbers ``int32 [n_mod, B, T]`` (-1 = zero row) of a batch; host arithmetic only.  A clip without any
        stored frame in its window raises, like the reference reader's assertion (reader_fns.py:97), unless
        ``allow_empty`` (the clip is then T zero rows).  strategy='random_clip' (training-time jitter,
        base_video_dataset.py:245-248,282-287) needs ``rng`` - the numpy Generator the reference dataset owns
        (``np.random.default_rng(seed)``, :155) - and draws from Python's global ``random`` (or ``pyrandom``) as the
        reference does: same seeds, same clips, same rows."""
        B = len(video_names)
        if bool(allow_empty) != getattr(self, "_allow_empty", False):
            self._check(self.lib.afft_store_allow_empty_clips(self.handle, int(bool(allow_empty))))
            self._allow_empty = bool(allow_empty)
        names = (C.c_char_p * max(B, 1))(*[v.encode() for v in video_names])
        st = np.ascontiguousarray(start_sec, dtype=np.float64)
        en = np.ascontiguousarray(end_sec, dtype=np.float64)
        idx = out if out is not None else torch.empty(len(self.mods), B, T, dtype=torch.int32)
        assert idx.dtype == torch.int32 and idx.is_contiguous() and idx.numel() >= len(self.mods) * B * T and not idx.is_cuda
        fids = torch.empty(len(self.mods), B, T, dtype=torch.int32) if want_frame_ids else None
        if strategy == "random_clip":
            if rng is None:
                raise ValueError("strategy='random_clip' needs the numpy Generator to draw from (rng=np.random.default_rng(seed))")
            sf, off = random_clip_draws(st, en, float(fps), int(T), frame_rate, len(self.mods), rng, pyrandom)
            self._check(self.lib.afft_store_plan_random(self.handle, B, names, st.ctypes.data, en.ctypes.data, float(fps), int(T),
                                                        float(frame_rate) if frame_rate else 0.0, sf.ctypes.data, off.ctypes.data,
                                                        idx.data_ptr(), fids.data_ptr() if fids is not None else None))
            return (idx, fids) if want_frame_ids else idx
        self._check(self.lib.afft_store_plan(self.handle, B, names, st.ctypes.data, en.ctypes.data, float(fps), int(T),
                                             float(frame_rate) if frame_rate else 0.0, SAMPLE_STRATEGIES[strategy],
                                             idx.data_ptr(), fids.data_ptr() if fids is not None else None))
        return (idx, fids) if want_frame_ids else idx


class FeatureStager:
    """Batches of clips -> device feature dict, double buffered.

    ``stage()`` plans on the host into a pinned buffer, uploads the plan and runs the gather kernel on ``stream``
    (default: a private staging stream), and returns ``(feature_dict, event, slot)``; make the consuming stream wait on
    the event (``torch.cuda.current_stream().wait_event(event)``) before the forward and call ``done(slot)`` after
    enqueuing it, so that the slot's buffers are not refilled before the forward has read them.  ``depth`` batches may
    be in flight."""

    def __init__(self, store: FeatureStore, T: int, max_batch: int, device="cuda:0", depth: int = 2, fps: float = 30.0,
                 frame_rate: Optional[float] = 4.0, strategy: str = "last_clip", rng=None, pyrandom=None):
        """rng / pyrandom: the generators 'random_clip' draws from (the dataset's ``np.random.default_rng(seed)`` and Python's
        ``random`` module by default), see FeatureStore.plan."""
        self.store, self.T, self.max_batch = store, T, max_batch
        self.rng, self.pyrandom = rng, pyrandom
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _capi.AfftError("FeatureStager gathers on the GPU; there is no CPU path")
        for m, t in store.rows.items():
            if not (t.is_cuda or t.is_pinned()):
                raise _capi.AfftError(f"row table of {m} is pageable host memory; use location='pinned' or 'cuda'")
        self.fps, self.frame_rate, self.strategy = fps, frame_rate, strategy
        n_mod = len(store.mods)
        self.depth = depth
        self.plan_host = [torch.empty(n_mod, max_batch, T, dtype=torch.int32).pin_memory() for _ in range(depth)]
        self.plan_dev = [torch.empty(n_mod, max_batch, T, dtype=torch.int32, device=self.device) for _ in range(depth)]
        self.out = [{m: torch.empty(max_batch, T, store.dims[m], device=self.device) for m in store.mods} for _ in range(depth)]
        self.events = [torch.cuda.Event() for _ in range(depth)]
        self.consumed: List[Optional[torch.cuda.Event]] = [None] * depth
        self.stream = torch.cuda.Stream(self.device)
        self._slot = 0

    def stage(self, video_names: Sequence[str], start_sec: Sequence[float], end_sec: Sequence[float], stream=None):
        B = len(video_names)
        if B > self.max_batch:
            raise _capi.AfftError(f"batch {B} > max_batch {self.max_batch}")
        k = self._slot
        self._slot = (k + 1) % self.depth
        self.events[k].synchronize()  # the slot's previous gather has finished reading plan_dev / the consumer was ordered after it
        n_mod, T = len(self.store.mods), self.T
        ph = self.plan_host[k].view(-1)[:n_mod * B * T].view(n_mod, B, T)
        self.store.plan(video_names, start_sec, end_sec, self.fps, T, self.frame_rate, self.strategy, out=ph, rng=self.rng,
                        pyrandom=self.pyrandom)
        stream = stream or self.stream
        if self.consumed[k] is not None:
            stream.wait_event(self.consumed[k])  # the forward that read this slot's tensors
            self.consumed[k] = None
        with torch.cuda.stream(stream):
            pd = self.plan_dev[k].view(-1)[:n_mod * B * T]
            pd.copy_(ph.view(-1), non_blocking=True)
            outs = (C.c_void_p * n_mod)(*[self.out[k][m].data_ptr() for m in self.store.mods])
            rc = self.store.lib.afft_store_gather(self.store.handle, B, T, pd.data_ptr(), outs, stream.cuda_stream)
            self.store._check(rc)
            self.events[k].record(stream)
        feats = {m: self.out[k][m][:B].view(B, T, self.store.dims[m], 1, 1, 1) for m in self.store.mods}
        return feats, self.events[k], k

    def done(self, slot: int, stream=None):
        """The consumer's work on the slot's tensors has been enqueued on ``stream`` (default: current stream)."""
        ev = torch.cuda.Event()
        ev.record(stream or torch.cuda.current_stream(self.device))
        self.consumed[slot] = ev
