"""CPU: MixUp and the loss / accuracy block of the training iteration (afft_b200/runner.py) against golden vectors written
from the reference's own modules (common/mixup.py, common/runner.py) by tests/golden/make_runner_golden.py.
Bit-exact for MixUp (same operations in the same order, same CPU generator for lambda); 1e-6 for the losses (the
reference averages over the kept rows after dropping the ignored ones, here a masked sum is divided by the count)."""
import os

import pytest
import torch

from afft_b200 import runner

GOLD = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "runner_mixup_loss.pt"))
CASES = GOLD["cases"]


def _clone(d):
    return None if d is None else {k: v.clone() for k, v in d.items()}


def _mixup(case, **kw):
    return runner.MixUp(alpha=case["alpha"], label_smoothing=GOLD["label_smoothing"], num_classes=GOLD["num_classes"], **kw)


@pytest.mark.parametrize("case", CASES, ids=[f"seed{c['seed']}" for c in CASES])
def test_mixup_matches_reference_bitwise(case):
    fn = _mixup(case)
    torch.manual_seed(100 + case["seed"])
    x_out, l_out, s_out, s_ign = fn(_clone(case["x"]), _clone(case["labels"]), _clone(case["sub"]))
    for m in case["x"]:
        assert torch.equal(x_out[m], case["x_out"][m]), m
    for k in case["labels"]:
        assert torch.equal(l_out[k], case["labels_out"][k]), k
    if case["sub"] is None:
        assert s_out is None and s_ign is None
    else:
        for k in case["sub"]:
            assert torch.equal(s_out[k], case["sub_out"][k]) and torch.equal(s_ign[k], case["sub_ignore"][k]), k


@pytest.mark.parametrize("case", CASES, ids=[f"seed{c['seed']}" for c in CASES])
def test_mixup_static_shape_branch(case, monkeypatch):
    """The path a captured step takes (no host read of the number of mixable clips): same result, also when fewer than two
    clips are mixable (seed 2: the reference returns the inputs untouched)."""
    monkeypatch.setattr(runner, "_capturing", lambda: True)
    fn = _mixup(case)
    torch.manual_seed(100 + case["seed"])
    x_out, l_out, s_out, _ = fn(_clone(case["x"]), _clone(case["labels"]), _clone(case["sub"]))
    for m in case["x"]:
        assert torch.equal(x_out[m], case["x_out"][m]), m
    for k in case["labels"]:
        assert torch.equal(l_out[k], case["labels_out"][k]), k
    if case["sub"] is not None:
        for k in case["sub"]:
            assert torch.equal(s_out[k], case["sub_out"][k]), k


@pytest.mark.parametrize("case", CASES, ids=[f"seed{c['seed']}" for c in CASES])
@pytest.mark.parametrize("mode", ["mixup", "hard"])
def test_losses_and_accuracy_match_reference(case, mode):
    if mode == "mixup":
        losses, keeps, metrics = runner.loss_and_accuracy(case["outputs"], case["labels_out"], case["sub_out"], mixup_enable=True,
                                                          target_subclips_ignore_index=case["sub_ignore"])
    else:
        losses, keeps, metrics = runner.loss_and_accuracy(case["outputs"], case["labels"], case["sub"], mixup_enable=False)
    total, means = runner.reduce_loss(losses, keeps, case["loss_wts"])
    ref = case["ref"][mode]
    assert set(means) == set(ref["means"])
    for k, v in ref["means"].items():
        assert abs(float(means[k]) - v) < 1e-6 * max(1.0, abs(v)), k
    assert abs(float(total) - ref["total"]) < 1e-6 * max(1.0, abs(ref["total"]))
    assert set(metrics) == set(ref["acc"])
    for k, v in ref["acc"].items():
        assert abs(float(metrics[k]) - v) < 1e-4, k


def test_partner_index_and_device_lambda():
    sel = torch.tensor([True, False, True, True, False, True])
    assert runner._partner_index(sel).tolist() == [5, 1, 3, 2, 4, 0]
    assert runner._partner_index(torch.zeros(4, dtype=torch.bool)).tolist() == [0, 1, 2, 3]
    fn = runner.MixUp(alpha=0.1, label_smoothing={"action": 0.4}, num_classes={"action": 7}, device_lambda=True)
    torch.manual_seed(3)
    lams = torch.stack([fn._lambda(torch.device("cpu")) for _ in range(200)])
    assert ((lams >= 0) & (lams <= 1)).all() and 0.3 < float(lams.mean()) < 0.7  # Beta(0.1, 0.1): symmetric, U-shaped
    assert float(((lams < 0.05) | (lams > 0.95)).float().mean()) > 0.5
    with pytest.raises(AssertionError):
        fn({"rgb": torch.zeros(1, 2, 3)}, {"action": torch.zeros(1, 1, dtype=torch.long)}, None)
    with pytest.raises(ValueError):
        runner.get_loss_wts({"cls_action": 1.0}, "past_reg_all-fused")
