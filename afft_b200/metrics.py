"""Anticipation metrics from top-k INDICES (the output of afft_marginalize_topk, row N2) instead of full score
matrices: top-k accuracy and class-mean top-k recall (MT5R), the numbers `test.py` prints.

Reference: common/utils.py:19-56 (`topk_accuracy`, `topk_recall`, from RULSTM) as called by
challenge.py:145-193 (`compute_accuracies_epic`).  The reference ranks with `scores.argsort()[:, ::-1]` on the host
for every class subset again; here the ranking is done once on the GPU (k = 5 indices per clip and task) and the
metrics are integer counting.  Results are identical whenever the top-k scores of a clip are distinct (ties are
broken towards the LOWER class index here, towards the higher one by the reference's reversed argsort).
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence

import numpy as np


def _as_np(x) -> np.ndarray:
    if hasattr(x, "detach"):
        x = x.detach().cpu().numpy()
    return np.asarray(x)


def topk_accuracy(topk, labels, ks: Sequence[int] = (1, 5), selected_class: Optional[int] = None) -> List[float]:
    """topk (n, K) class indices, best first; labels (n,).  common/utils.py:19-42."""
    topk, labels = _as_np(topk), _as_np(labels).reshape(-1)
    if selected_class is not None:
        idx = labels == selected_class
        topk, labels = topk[idx], labels[idx]
    if max(ks) > topk.shape[1]:
        raise ValueError(f"k = {max(ks)} needs at least that many ranked indices per clip (have {topk.shape[1]})")
    tp = topk[:, :max(ks)] == labels.reshape(-1, 1)
    return [float(tp[:, :k].max(1).mean()) for k in ks]


def topk_recall(topk, labels, k: int = 5, classes: Optional[Iterable[int]] = None) -> float:
    """Mean over classes of the top-k accuracy restricted to the clips of that class.  common/utils.py:45-56."""
    topk, labels = _as_np(topk), _as_np(labels).reshape(-1)
    unique = np.unique(labels)
    classes = unique if classes is None else np.intersect1d(np.asarray(list(classes)), unique)
    hit = (topk[:, :k] == labels.reshape(-1, 1)).any(1)
    # per-class mean of `hit`, then the mean over classes - the reference loops over classes and re-ranks each time
    recalls = [hit[labels == c].mean() for c in classes]
    return float(np.sum(recalls) / len(classes))


def epic_metrics(topk_verb, topk_noun, topk_action, verb_labels, noun_labels, action_labels, many_shot=None) -> dict:
    """The dictionary `compute_accuracies_epic` builds (challenge.py:145-193): top-1 / top-5 accuracy and mean top-5
    recall for verb, noun and action, in percent; `many_shot` = (verbs, nouns, actions) class lists for the *_ms keys."""
    out = {}
    for name, tk, lab, ms in (("v", topk_verb, verb_labels, 0), ("n", topk_noun, noun_labels, 1), ("a", topk_action, action_labels, 2)):
        t1, t5 = topk_accuracy(tk, lab, ks=(1, 5))
        out[f"{name}top1"], out[f"{name}top5"] = t1 * 100, t5 * 100
        out[f"{name}mt5r"] = topk_recall(tk, lab, k=5) * 100
        if many_shot is not None:
            out[f"{name}mt5r_ms"] = topk_recall(tk, lab, k=5, classes=many_shot[ms]) * 100
    return out
