"""CPU oracle of the INPUT side of the hot path: which stored feature rows make up a
clip's (T, C) input, and the gather itself.  TEST INFRASTRUCTURE ONLY - imported by tests/, never by the product
(afft_b200/staging.py + csrc/staging.cu run the native implementation).

It restates, in plain Python / numpy, the arithmetic spread over three reference functions:

  * ``BaseVideoDataset._sample``                    datasets/base_video_dataset.py:225-337
        window [start, end] -> [new_start, new_end] ('last' / 'center' / 'first' strategies), the every-k-th-frame
        subsampling from the back (:279-292), front/back padding to ``frames_per_clip`` (:309-335)
  * ``EpicRULSTMFeatsReader._read_rulstm_features``   datasets/reader_fns.py:108-138
        frame ids floor(start*fps)+1 .. floor(end*fps); ids < 1 replaced by the smallest valid id (:121-124);
        audio / poses LMDBs are indexed in the ORIGINAL video's frame rate (:131-133, :140-157)
  * ``EpicRULSTMFeatsReader.read_representations``    datasets/reader_fns.py:65-106
        key ``{video}_frame_{id:010d}.jpg``; a missing frame falls back to the closest EARLIER stored frame within
        9 frames (:76-80), otherwise to a row of zeros (:92-96); at least one frame of the clip must exist (:93)

The reference reads EVERY frame of the window (135 LMDB gets for an 18-step clip at 30 fps) and then keeps every
8th; ``clip_frame_ids`` returns just the kept ids, in the order the model sees them, so an implementation needs T
lookups per modality.  Pinned against the reference classes themselves (with a dict-backed stand-in for the LMDB
environment) by tests/golden/make_feats_golden.py -> tests/golden/feats_reader.npz.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import numpy as np

STRATEGIES = ("last_clip", "center_clip", "first_clip", "random_clip")  # base_video_dataset.py:28-31
SEARCH_RADIUS = 10                          # reader_fns.py:76  range(10)


def _py_round(x: float) -> int:
    return int(round(x))  # Python 3 round(): half to even, like np.rint


def random_draws(start: float, end: float, fps: float, frames_per_clip: int, frame_rate: Optional[float], rng, pyrandom):
    """The two random numbers ``_sample`` consumes for ONE (clip, modality) call with sample_strategy == 'random_clip', in
    the reference's order: ``rng.integers(start_frame)`` (numpy Generator, only when the window is longer than the clip;
    base_video_dataset.py:246-248) and, after the read, ``random.random()`` (Python's global generator, :283-286).
    Returns (start_frame, offset)."""
    start = max(start, 0)
    end = max(end, 0)
    req_fps = fps if frame_rate is None else frame_rate
    nframes = int(fps * (end - start))
    frames_to_ext = _py_round(frames_per_clip * (fps / req_fps))
    start_frame = max(nframes - frames_to_ext, 0)
    if start_frame > 0:
        start_frame = int(rng.integers(start_frame))
    shift = max(_py_round(fps / req_fps / 3), 1)  # "we select frame randomly from the 1/3 of the desired time zone"
    offset = _py_round(pyrandom.random() * shift)
    return start_frame, offset


def window(start: float, end: float, fps: float, frames_per_clip: int, frame_rate: Optional[float], strategy: str,
           rand_start_frame: Optional[int] = None):
    """base_video_dataset.py:236-263 -> (new_start, new_end) in seconds.  'random_clip': rand_start_frame is the draw."""
    start = max(start, 0)
    end = max(end, 0)
    req_fps = fps if frame_rate is None else frame_rate
    nframes = int(fps * (end - start))
    frames_to_ext = _py_round(frames_per_clip * (fps / req_fps))
    if strategy == "center_clip":
        start_frame = max((nframes - frames_to_ext) // 2, 0)
    elif strategy == "last_clip":
        start_frame = max(nframes - frames_to_ext, 0)
    elif strategy == "first_clip":
        start_frame = 0
    elif strategy == "random_clip":
        assert rand_start_frame is not None, "random_clip needs the drawn start frame (random_draws)"
        start_frame = int(rand_start_frame)
    else:
        raise NotImplementedError(strategy)
    new_start = start + max(start_frame / fps, 0)
    new_end = start + max((start_frame + frames_to_ext) / fps, 0)
    new_end = max(min(end, new_end), 0)
    new_start = min(max(new_start, 0), new_end)
    return new_start, new_end


def window_frame_ids(new_start: float, new_end: float, fps: float) -> np.ndarray:
    """reader_fns.py:116-124: every frame id of the window, ascending."""
    start_frame = np.floor(new_start * fps)
    end_frame = np.floor(new_end * fps)
    frames = np.arange(end_frame, start_frame, -1).astype(int)[::-1]
    if frames.size == 0 or frames.max() < 1:
        raise AssertionError("the dataset should not have clips without a frame id >= 1 (reader_fns.py:122)")
    frames[frames < 1] = frames[frames >= 1].min()
    return frames


def kept_positions(n: int, fps: float, frame_rate: Optional[float], frames_per_clip: int, strategy: str,
                   rand_offset: int = 0) -> List[int]:
    """base_video_dataset.py:279-335: positions (into the window's frame list) the model finally sees, padded.
    'random_clip' subsamples from the back like 'last_clip', shifts every kept position by the drawn offset where that
    stays positive (:283-287), and pads / crops like 'last_clip' (:313-316,326-329)."""
    req_fps = fps if frame_rate is None else frame_rate
    step = max(_py_round(fps / req_fps), 1)
    from_back = strategy in ("last_clip", "random_clip")
    if from_back:
        keep = list(range(n))[::-step][::-1]
        if strategy == "random_clip":
            keep = [i - rand_offset if i - rand_offset > 0 else i for i in keep]
    else:
        keep = list(range(n))[::step]
    if len(keep) < frames_per_clip:
        npad = frames_per_clip - len(keep)
        keep = [keep[0]] * npad + keep if from_back else keep + [keep[-1]] * npad
    return keep[-frames_per_clip:] if from_back else keep[:frames_per_clip]


def orig_video_fps(video_name: str) -> float:
    """reader_fns.py:148-157."""
    n = len(video_name.split("_")[-1])
    if n == 3:
        return 50.0
    if n == 2:
        return 59.94005994005994
    raise ValueError(f"Unkown video name format: {video_name}")


def clip_frame_ids(video_name: str, start: float, end: float, fps: float, frames_per_clip: int,
                   frame_rate: Optional[float], strategy: str = "last_clip", orig_fps_index: bool = False,
                   rand: Optional[Sequence[int]] = None) -> np.ndarray:
    """The T frame ids whose features form the clip, in model order.  ``orig_fps_index``: the store is keyed in the
    original video's frame rate (audio, poses).  ``rand``: (start_frame, offset) of ``random_draws`` for 'random_clip'."""
    ns, ne = window(start, end, fps, frames_per_clip, frame_rate, strategy, rand[0] if rand is not None else None)
    frames = window_frame_ids(ns, ne, fps)
    if orig_fps_index:
        frames = np.rint(frames / fps * orig_video_fps(video_name)).astype(int)  # reader_fns.py:143-145
    keep = kept_positions(len(frames), fps, frame_rate, frames_per_clip, strategy, rand[1] if rand is not None else 0)
    return frames[keep]


def lookup(store: Dict[str, np.ndarray], video_name: str, frame_id: int) -> Optional[np.ndarray]:
    """reader_fns.py:71-91 for one frame: the stored row of the frame or of the closest earlier stored frame."""
    for r in range(SEARCH_RADIUS):
        row = store.get(f"{video_name}_frame_{frame_id - r:010d}.jpg")
        if row is not None:
            return row
    return None


def gather_clip(store: Dict[str, np.ndarray], video_name: str, frame_ids: Sequence[int], width: int) -> np.ndarray:
    """(T, width) float32; frames without a stored row within the search radius are zero rows (reader_fns.py:92-96).
    The reference additionally raises when NO frame of the whole window is stored (:93 - it needs one row to learn the
    width); an implementation that reads only the T kept rows cannot reproduce that sanity check and returns zeros."""
    out = np.zeros((len(frame_ids), width), dtype=np.float32)
    for i, f in enumerate(frame_ids):
        row = lookup(store, video_name, int(f))
        if row is not None:
            out[i] = row
    return out


def read_batch(stores: Dict[str, Dict[str, np.ndarray]], orig_fps_mods: Sequence[str], video_names: Sequence[str],
               starts: Sequence[float], ends: Sequence[float], fps: float, frames_per_clip: int,
               frame_rate: Optional[float], strategy: str = "last_clip", widths: Optional[Dict[str, int]] = None) -> Dict[str, np.ndarray]:
    """{modality: (B, T, C_m) float32} - the collated model input before the trailing singleton dims
    (SURVEY Appendix B.0)."""
    out = {}
    for mod, store in stores.items():
        clips = []
        for v, s, e in zip(video_names, starts, ends):
            ids = clip_frame_ids(v, s, e, fps, frames_per_clip, frame_rate, strategy, orig_fps_index=mod in orig_fps_mods)
            width = widths[mod] if widths is not None else len(next(iter(store.values())))
            clips.append(gather_clip(store, v, ids, width))
        out[mod] = np.stack(clips)
    return out


def nextafter_cases():
    """Window bounds that sit exactly on frame boundaries (floor() edge cases) for the parity tests."""
    vals = []
    for k in (1, 7, 30, 31, 135, 136, 1000):
        t = k / 30.0
        vals += [t, math.nextafter(t, 0.0), math.nextafter(t, 1e9)]
    return vals
