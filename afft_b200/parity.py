"""Parity statistics between two sets of logits (no oracle import here: callers pass the reference tensor).

The north star asks for "logits within tolerance and identical top-5 indices" against the fp32 reference
(reference test.py:86 reads ``outputs['logits/action'][...][:, 0, :]``; common/utils.py:19-42 ranks them).  With
random-init weights the gap between the 5th and 6th largest logit of a clip is often below the rounding error of a
16-bit operand format, so agreement is reported as a rate, stratified by that gap.
"""
from __future__ import annotations

from typing import Dict

import torch

GAP_EDGES = (0.0, 1e-4, 1e-3, 1e-2, float("inf"))


def top5_stats(got: torch.Tensor, ref: torch.Tensor, k: int = 5, tie_eps: float = 2e-4) -> Dict:
    """got, ref: [n, C] logits (any device).  Ordered top-k identity per row, max / mean |difference|, and the identity
    rate per bin of the reference's k-th to (k+1)-th logit gap.

    ``near_tie`` counts the clips in which two of the reference's top-(k+1) logits are closer than ``tie_eps``: the fp32
    reference itself orders such a pair differently from run to run (its matmul blocking changes the sums by ~1e-6, and
    by more across batch sizes), so only the remaining clips have a well-defined ordered top-k; ``mismatch_clear`` counts
    the ordered mismatches among those."""
    got = got.detach().float().cpu()
    ref = ref.detach().float().cpu()
    assert got.shape == ref.shape and got.dim() == 2
    n = got.shape[0]
    diff = (got - ref).abs()
    rt = ref.topk(k + 1, dim=-1)
    same = (got.topk(k, dim=-1).indices == rt.indices[:, :k]).all(-1)
    same_set = (got.topk(k, dim=-1).indices.sort(-1).values == rt.indices[:, :k].sort(-1).values).all(-1)
    gap = rt.values[:, k - 1] - rt.values[:, k]
    min_adj = (rt.values[:, :-1] - rt.values[:, 1:]).min(dim=-1).values  # closest pair among the reference's top-(k+1)
    near_tie = min_adj < tie_eps
    bins = []
    for lo, hi in zip(GAP_EDGES[:-1], GAP_EDGES[1:]):
        m = (gap >= lo) & (gap < hi)
        cnt = int(m.sum())
        bins.append({"gap_ge": lo, "gap_lt": (hi if hi != float("inf") else None), "clips": cnt,
                     "ordered_top5_identical": int((same & m).sum())})
    return {"clips": n, "max_abs_dlogit": round(float(diff.max()), 7), "mean_abs_dlogit": round(float(diff.mean()), 8),
            "ordered_top5_identical": int(same.sum()), "ordered_top5_identity_rate": round(float(same.float().mean()), 5),
            "top5_set_identity_rate": round(float(same_set.float().mean()), 5),
            "top1_identity_rate": round(float((got.argmax(-1) == ref.argmax(-1)).float().mean()), 5),
            "near_tie": int(near_tie.sum()), "tie_eps": tie_eps, "mismatch_clear": int((~same & ~near_tie).sum()),
            "median_ref_gap_5th_6th": round(float(gap.median()), 6), "by_ref_gap": bins}
