"""Writes tests/golden/runner_mixup_loss.pt: inputs and outputs of the REFERENCE's MixUp (common/mixup.py) and
BasicLossAccuracy / Runner._reduce_loss (common/runner.py) on small seeded cases, for tests/test_runner.py.
Run in the build container (needs /root/reference):  python tests/golden/make_runner_golden.py"""
import os
import sys

import torch

import types

REF = os.environ.get("AFFT_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
for _name in ("submitit", "cv2"):  # imported by common/utils.py, not used by the functions exercised here
    try:
        __import__(_name)
    except ImportError:
        sys.modules.setdefault(_name, types.ModuleType(_name))
from common import mixup as ref_mixup  # noqa: E402
from common import runner as ref_runner  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "runner_mixup_loss.pt")
NUM_CLASSES = {"action": 11, "verb": 5}
SMOOTH = {"action": 0.4, "verb": 0.01}
B, T, D = 6, 4, 8
cases = []
for seed, n_ignored_clips, with_subclips in ((0, 0, True), (1, 2, True), (2, 5, True), (3, 0, False), (4, 3, True)):
    g = torch.Generator().manual_seed(seed)
    x = {"rgb": torch.randn(B, T, D, generator=g), "flow": torch.randn(B, T, 5, generator=g)}
    labels = {k: torch.randint(0, c, (B, 1), generator=g) for k, c in NUM_CLASSES.items()}
    sub = None
    if with_subclips:
        sub = {k: torch.randint(0, c, (B, T), generator=g) for k, c in NUM_CLASSES.items()}
        for b in range(n_ignored_clips):  # the ignore class marks the same positions for every label type
            t = int(torch.randint(0, T, (1,), generator=g))
            for k in sub:
                sub[k][b, t] = -1
    fn = ref_mixup.MixUp(alpha=0.1 if seed % 2 == 0 else 0.8, label_smoothing=SMOOTH, num_classes=NUM_CLASSES)
    torch.manual_seed(100 + seed)
    x_in = {m: t.clone() for m, t in x.items()}
    x_out, l_out, ls_out, ls_ign = fn({m: t.clone() for m, t in x.items()}, {k: v.clone() for k, v in labels.items()},
                                      None if sub is None else {k: v.clone() for k, v in sub.items()})
    # losses on logits drawn from the same generator, through the reference's loss block and reduction
    outputs = {"orig_past": {"all-fused": torch.randn(B, T, D, generator=g)},
               "past_futures": {"all-fused": torch.randn(B, T, D, generator=g)}}
    for k, c in NUM_CLASSES.items():
        outputs[f"logits/{k}"] = {"all-fused": torch.randn(B, 1, c, generator=g)}
        outputs[f"past_logits/{k}"] = {"all-fused": torch.randn(B, T, c, generator=g)}
    crit = ref_runner.BasicLossAccuracy()
    wts = {"cls_action": 1.0, "cls_verb": 0.5, "past_cls_action": 1.0, "past_cls_verb": 0.0, "past_reg": 2.0}
    res = {}
    for mode in ("mixup", "hard"):
        if mode == "mixup":
            losses, metrics = crit(outputs, l_out, ls_out, mixup_enable=True, target_subclips_ignore_index=ls_ign)
        else:
            losses, metrics = crit(outputs, labels, sub, mixup_enable=False, target_subclips_ignore_index=None)
        total, means = ref_runner.Runner._reduce_loss(losses, wts)
        res[mode] = {"total": float(total), "means": {k: float(v) for k, v in means.items() if k != "total_loss"},
                     "acc": {k: float(v) for k, v in metrics.items() if k.startswith("acc")}}
    cases.append({"seed": seed, "alpha": fn.mixup_beta_sampler.concentration0.item(), "x": x_in, "labels": labels, "sub": sub,
                  "x_out": x_out, "labels_out": l_out, "sub_out": ls_out, "sub_ignore": ls_ign, "outputs": outputs,
                  "loss_wts": wts, "ref": res})
torch.save({"num_classes": NUM_CLASSES, "label_smoothing": SMOOTH, "cases": cases}, OUT)
print("wrote", OUT, os.path.getsize(OUT), "bytes")
