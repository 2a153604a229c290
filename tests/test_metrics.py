"""Anticipation metrics from top-k indices (afft_b200/metrics.py) against the reference's score-matrix functions
(common/utils.py:19-56 topk_accuracy / topk_recall; challenge.py:94-106 compute_accuracy).

The committed fixture tests/golden/metrics_ref.npz holds scores, labels and the REFERENCE's results (written here by
`PYTHONPATH=. python tests/test_metrics.py` with the reference functions imported from /root/reference); when the reference is
present the functions are also called live."""
import os
import sys
import types

import numpy as np
import pytest

from afft_b200 import metrics

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "metrics_ref.npz")
REF = "/root/reference"


def _reference_fns():
    if not os.path.isdir(REF):
        return None
    from oracle import ref_shim
    ref_shim.install_stubs()
    sys.modules.setdefault("cv2", types.ModuleType("cv2"))
    import importlib
    u = importlib.import_module("common.utils")
    return u.topk_accuracy, u.topk_recall


def _cases(seed=5):
    rng = np.random.default_rng(seed)
    out = []
    for n, C in ((64, 97), (128, 300), (96, 512), (33, 19)):
        scores = rng.standard_normal((n, C)).astype(np.float32)
        labels = rng.integers(0, C, size=n)
        # make the metric non-trivial: push the label's score up for ~half of the clips
        boost = rng.random(n) < 0.5
        scores[np.arange(n)[boost], labels[boost]] += 3.0
        out.append((scores, labels))
    return out


def _top5(scores):
    # ranking as the GPU kernel produces it: best first, ties -> lower index (no ties in these float scores)
    return np.argsort(-scores, axis=1, kind="stable")[:, :5]


def _write_fixture():
    fns = _reference_fns()
    assert fns is not None
    acc, rec = fns
    data = {}
    for i, (scores, labels) in enumerate(_cases()):
        data[f"scores_{i}"], data[f"labels_{i}"] = scores, labels
        data[f"acc_{i}"] = np.array(acc(scores, labels, ks=(1, 5)))
        data[f"mt5r_{i}"] = np.array(rec(scores, labels, k=5))
        sub = np.unique(labels)[::3]
        data[f"classes_{i}"] = sub
        data[f"mt5r_sub_{i}"] = np.array(rec(scores, labels, k=5, classes=sub))
        data[f"acc_sel_{i}"] = np.array(acc(scores, labels, ks=(1, 5), selected_class=int(labels[0])))
    np.savez_compressed(GOLDEN, **data)


def test_metrics_match_reference_fixture():
    z = np.load(GOLDEN)
    for i in range(4):
        scores, labels = z[f"scores_{i}"], z[f"labels_{i}"]
        tk = _top5(scores)
        assert np.allclose(metrics.topk_accuracy(tk, labels, ks=(1, 5)), z[f"acc_{i}"], rtol=0, atol=1e-12)
        assert abs(metrics.topk_recall(tk, labels, k=5) - float(z[f"mt5r_{i}"])) < 1e-12
        assert abs(metrics.topk_recall(tk, labels, k=5, classes=z[f"classes_{i}"]) - float(z[f"mt5r_sub_{i}"])) < 1e-12
        assert np.allclose(metrics.topk_accuracy(tk, labels, ks=(1, 5), selected_class=int(labels[0])), z[f"acc_sel_{i}"], atol=1e-12)


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present (GPU box)")
def test_metrics_match_reference_live():
    acc, rec = _reference_fns()
    for scores, labels in _cases(seed=99):
        tk = _top5(scores)
        assert np.allclose(metrics.topk_accuracy(tk, labels, ks=(1, 5)), acc(scores, labels, ks=(1, 5)), atol=1e-12)
        assert abs(metrics.topk_recall(tk, labels, k=5) - rec(scores, labels, k=5)) < 1e-12


def test_epic_metrics_keys_and_errors():
    scores, labels = _cases()[0]
    tk = _top5(scores)
    m = metrics.epic_metrics(tk, tk, tk, labels, labels, labels, many_shot=(np.unique(labels)[:5],) * 3)
    assert set(m) == {f"{p}{k}" for p in "vna" for k in ("top1", "top5", "mt5r", "mt5r_ms")}
    assert all(0.0 <= v <= 100.0 for v in m.values())
    with pytest.raises(ValueError):
        metrics.topk_accuracy(tk[:, :3], labels, ks=(1, 5))


if __name__ == "__main__":
    _write_fixture()
    print("wrote", GOLDEN)
