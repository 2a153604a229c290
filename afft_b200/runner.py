"""The training iteration's glue around the model: MixUp and the loss / accuracy block.

Reference: common/mixup.py (MixUp with an ignore class, :10-182), common/runner.py (MultiDimCrossEntropy :12-37,
BasicLossAccuracy :40-168, Runner._reduce_loss :196-211, the call order :213-267).  The experiment behind BASELINE
config 5 trains with ``train.use_mixup=true, mixup_backbone=true, mixup_alpha=0.1`` and label smoothing 0.4 / 0.01 / 0.03
(expts/01_SA-Fuser_ek100_train.txt:10-12, conf/config.yaml:16-22): the model receives ``mixup_fn`` and applies it to the
backbone outputs (models/base_model.py:53-56), and the losses are soft-label cross-entropies.

Same arithmetic, restated with STATIC SHAPES so that the whole step (MixUp -> forward -> losses -> backward -> optimizer)
can sit in one CUDA graph: the reference selects the clips without ignored sub-clip labels by boolean indexing
(``inputs[batch_wo_ignore_index]``, ``inp[keep_index]``: data-dependent shapes, host synchronisation); here the same clips
are addressed through a partner-index gather and masks.  Element for element the operations are the reference's
(``x * lam + flip(x) * (1 - lam)`` in that order), so for the same lambda the results are bit-identical
(tests/test_runner.py, golden vectors written from the reference modules).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

PAST_LOGITS_PREFIX = "past_"


def _capturing() -> bool:
    return torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()


def batch_wo_ignore_cls(target_subclips: torch.Tensor, ignore_cls: int = -1) -> torch.Tensor:
    """(B,) bool: clips none of whose sub-clip labels is the ignore class (common/mixup.py:10-15)."""
    t = target_subclips.squeeze(-1)
    assert t.ndim == 2, "Target subclips should have dimension of 2."
    return (t != ignore_cls).all(-1)


def convert_to_one_hot(targets: torch.Tensor, num_class: int, label_smooth: float = 0.0) -> torch.Tensor:
    """Index labels -> smoothed one-hot rows (common/mixup.py:18-47).  The reference's range assertion reads the maximum
    back to the host; it is skipped while a CUDA graph is being captured."""
    if not _capturing():
        assert torch.max(targets).item() < num_class, "Class Index must be less than number of classes"
    assert 0 <= label_smooth < 1.0, "Label smooth value needs to be between 0 and 1."
    targets = targets.squeeze(-1)
    non_target_value = label_smooth / num_class
    target_value = 1.0 - label_smooth + non_target_value
    one_hot = torch.full((*targets.shape, num_class), non_target_value, device=targets.device)
    one_hot.scatter_(-1, targets.unsqueeze(-1), target_value)
    return one_hot


def _partner_index(sel: torch.Tensor) -> torch.Tensor:
    """For ``sel`` (B,) bool: the row each selected row is mixed with when the selected rows are flipped among themselves
    (common/mixup.py:85-89: ``inputs[sel].flip(0)``); unselected rows (and a lone selected row) point at themselves."""
    B = sel.shape[0]
    idx = torch.arange(B, device=sel.device)
    order = torch.argsort((~sel).to(torch.int8), stable=True)  # selected rows first, original order kept
    n_sel = sel.sum()
    rank = torch.cumsum(sel.to(torch.int64), 0) - 1            # rank of a selected row among the selected
    partner = order[(n_sel - 1 - rank).clamp(0, B - 1)]
    return torch.where(sel, partner, idx)


def _mix(x: torch.Tensor, sel: Optional[torch.Tensor], lam: torch.Tensor) -> torch.Tensor:
    """common/mixup.py:76-89 on the rows ``sel`` (None: every row, the reference's ``[...]`` index)."""
    if sel is None:
        return x * lam + x.flip(0) * (1.0 - lam)
    partner = _partner_index(sel)
    mixed = x * lam + x[partner] * (1.0 - lam)
    return torch.where(sel.view(-1, *([1] * (x.ndim - 1))), mixed, x)


class MixUp(torch.nn.Module):
    """common/mixup.py:92-182.  ``forward(x_video, labels, labels_subclips)`` ->
    ``(x_out, labels_out, labels_subclips_out, labels_subclips_ignore_index)`` with one-hot (smoothed, mixed) labels.

    lambda ~ Beta(alpha, alpha).  ``device_lambda=False`` draws it exactly as the reference does (CPU generator, one
    host scalar per call); ``device_lambda=True`` draws it on the inputs' device from the CUDA generator, which is what a
    captured training step needs (a new lambda on every replay)."""

    def __init__(self, alpha: float = 1.0, label_smoothing: Dict = 0.0, num_classes: Dict = None, one_hot: bool = False,
                 ignore_cls: int = -1, device_lambda: bool = False) -> None:
        super().__init__()
        self.alpha = float(alpha)
        self.mixup_beta_sampler = torch.distributions.beta.Beta(alpha, alpha)
        self.label_smoothing = label_smoothing
        self.num_classes = num_classes
        self.one_hot = one_hot
        self.ignore_cls = ignore_cls
        self.device_lambda = device_lambda

    def _lambda(self, device) -> torch.Tensor:
        if not self.device_lambda:
            return self.mixup_beta_sampler.sample().to(device)
        a = torch.full((2,), self.alpha, device=device)
        g = torch._standard_gamma(a)  # Beta(a, a) = G1 / (G1 + G2)
        return g[0] / (g[0] + g[1])

    def forward(self, x_video: Dict, labels: Dict, labels_subclips: Optional[Dict]):
        first = next(iter(x_video.values()))
        assert first.size(0) > 1, "MixUp cannot be applied to a single instance."
        labels_out = {k: convert_to_one_hot(v, self.num_classes[k], self.label_smoothing[k]) for k, v in labels.items()}
        sel = None
        labels_subclips_out = labels_subclips_ignore_index = None
        if labels_subclips is not None:
            sel = batch_wo_ignore_cls(next(iter(labels_subclips.values())), self.ignore_cls)
            labels_subclips_out, labels_subclips_ignore_index = {}, {}
            for k, v in labels_subclips.items():
                ignore = v == self.ignore_cls
                labels_subclips_ignore_index[k] = ignore
                # ignored positions get class 0 so that the one-hot conversion works; the loss drops them (:150-156)
                labels_subclips_out[k] = convert_to_one_hot(torch.where(ignore, torch.zeros_like(v), v), self.num_classes[k],
                                                            self.label_smoothing[k])
            # fewer than two mixable clips: the reference returns before drawing lambda (:158-160).  Eagerly that is
            # mirrored; under capture the partner gather makes every row its own partner, which is the same result
            if not _capturing() and int(sel.sum()) <= 1:
                return x_video, labels_out, labels_subclips_out, labels_subclips_ignore_index
            sel = sel & (sel.sum() > 1)
        lam = self._lambda(first.device)
        x_out = {m: _mix(x, sel, lam) for m, x in x_video.items()}
        labels_out = {k: _mix(v, sel, lam) for k, v in labels_out.items()}
        if labels_subclips is None:
            return x_video, labels_out, None, None  # the reference hands back the UNMIXED inputs here (:176-177)
        labels_subclips_out = {k: _mix(v, sel, lam) for k, v in labels_subclips_out.items()}
        return x_out, labels_out, labels_subclips_out, labels_subclips_ignore_index


def multi_dim_cross_entropy(inp: torch.Tensor, tgt: torch.Tensor, one_hot: bool = False,
                            ignore_index: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """MultiDimCrossEntropy (common/runner.py:12-37; ignore_index=-1, reduction='none') -> (per-row loss, keep mask).
    The reference drops the ignored rows before the loss; here they stay, with loss 0 and keep = False, and ``reduce_loss``
    averages over the kept rows - the same mean without a data-dependent shape."""
    inp = inp.reshape(-1, inp.size(-1))
    if not one_hot:
        assert ignore_index is None, "Target should be one-hotted."
        return F.cross_entropy(inp, tgt.reshape(-1), ignore_index=-1, reduction="none"), None
    tgt = tgt.reshape(-1, tgt.size(-1))
    loss = F.cross_entropy(inp, tgt, reduction="none")
    if ignore_index is None:
        return loss, None
    keep = ~ignore_index.reshape(-1)
    return torch.where(keep, loss, torch.zeros_like(loss)), keep


def accuracy(output: torch.Tensor, target: torch.Tensor, topk=(1,)):
    """common/utils.py:59-86: top-k hits as a percentage of ALL rows (negative targets never match, and when every target is
    negative the reference returns zeros - which the same formula yields, so its host-side branch is not needed)."""
    with torch.no_grad():
        output = output.flatten(0, -2)
        target = target.flatten()
        maxk = max(topk)
        batch_size = target.size(0)
        _, pred = output.topk(maxk, 1, True, True)
        correct = pred.t().eq(target[None])
        return [correct[:k].flatten().sum(dtype=torch.float32) * (100.0 / batch_size) for k in topk]


def loss_and_accuracy(outputs, target: Dict, target_subclips: Optional[Dict], mixup_enable: bool = False,
                      target_subclips_ignore_index: Optional[Dict] = None):
    """BasicLossAccuracy.forward (common/runner.py:112-168): per-row future / past classification losses, the past feature
    regression, acc1 / acc5.  Returns (losses, keeps, metrics); ``keeps[key]`` is the row mask of a loss whose mean runs
    over a subset (past classification under MixUp), else None.  (The reference also parks the logits on the host for the
    epoch-level MT5R; that belongs to the metric tracker, not to the step.)"""
    losses, keeps, metrics = {}, {}, {}
    for tgt_type, tgt_val in target.items():
        for modk, logits in outputs[f"logits/{tgt_type}"].items():
            assert logits.ndim == 3
            key = f"cls_{tgt_type}_{modk}"
            losses[key], keeps[key] = multi_dim_cross_entropy(logits, tgt_val, one_hot=mixup_enable)
            if mixup_enable:  # the two largest soft labels are the two mixed classes: count either as a hit (:61-77)
                top2 = torch.topk(tgt_val, 2, dim=1).indices
                preds = logits.detach().clone()
                p0 = preds[:, 0]  # view of sequence index 0; gather / scatter keep every index on the device (capturable)
                first, second = top2[:, :1], top2[:, 1:2]
                p0.scatter_(1, first, p0.gather(1, first) + p0.gather(1, second))
                p0.scatter_(1, second, 0.0)
                labels = top2[:, 0]
            else:
                preds, labels = logits.detach(), tgt_val
            if labels.ndim == 1:
                labels = labels.unsqueeze(-1)
            acc1, acc5 = accuracy(preds, labels, topk=(1, min(5, preds.size(-1))))
            metrics[f"acc1_{tgt_type}_{modk}"], metrics[f"acc5_{tgt_type}_{modk}"] = acc1, acc5
        past_key = f"{PAST_LOGITS_PREFIX}logits/{tgt_type}"
        if past_key in outputs and target_subclips is not None:
            for modk, past_logits in outputs[past_key].items():
                key = f"past_cls_{tgt_type}_{modk}"
                past_target = target_subclips[tgt_type]
                if mixup_enable:
                    assert past_logits.shape == past_target.shape
                    assert target_subclips_ignore_index is not None
                    losses[key], keeps[key] = multi_dim_cross_entropy(past_logits, past_target, one_hot=True,
                                                                      ignore_index=target_subclips_ignore_index[tgt_type])
                else:
                    past_target = past_target.squeeze(-1) if past_target.ndim == past_logits.ndim else past_target
                    assert past_logits.shape[:-1] == past_target.shape
                    losses[key], keeps[key] = multi_dim_cross_entropy(past_logits, past_target)
        if "orig_past" in outputs and "past_futures" in outputs:
            for modk, upd in outputs["past_futures"].items():
                if modk not in outputs["orig_past"]:
                    continue
                key = f"past_reg_{modk}"
                losses[key], keeps[key] = F.mse_loss(upd[:, 1:], outputs["orig_past"][modk][:, 1:]), None
    return losses, keeps, metrics


def get_loss_wts(loss_wts: Dict, key: str) -> float:
    """common/runner.py:170-174: the first configured prefix of the key."""
    for k, v in loss_wts.items():
        if key.startswith(k):
            return v
    raise ValueError(f"{key} not contained in predefined loss_wts: {loss_wts}")


DEFAULT_LOSS_WTS = {"cls_action": 1.0, "cls_verb": 1.0, "cls_noun": 1.0, "past_cls_action": 1.0, "past_cls_verb": 1.0,
                    "past_cls_noun": 1.0, "past_reg": 1.0}  # conf/config.yaml:24-35


def reduce_loss(losses: Dict, keeps: Dict, loss_wts: Dict = None):
    """Runner._reduce_loss (common/runner.py:196-211): mean of every loss (over its kept rows), weighted sum of those
    with a positive weight.  Returns (total, {key: mean}) as tensors - no host read, so it can be captured; the
    reference's NaN check and ``.item()`` bookkeeping are the caller's business."""
    loss_wts = DEFAULT_LOSS_WTS if loss_wts is None else loss_wts
    means = {}
    for key, val in losses.items():
        keep = keeps.get(key)
        means[key] = torch.mean(val) if keep is None else val.sum() / keep.sum()
    terms = [get_loss_wts(loss_wts, k) * v for k, v in means.items() if get_loss_wts(loss_wts, k) > 0]
    return torch.sum(torch.stack(terms)), means


def training_losses(model, feature_dict: Dict, target: Dict, target_subclips: Optional[Dict], mixup_fn=None,
                    mixup_backbone: bool = True, loss_wts: Dict = None):
    """Runner.__call__ (common/runner.py:213-267) from device tensors: MixUp on the inputs or on the backbone outputs,
    the model, the losses.  Returns (total loss, per-loss means, metrics)."""
    kwargs = dict(mixup_fn=None, target=target, target_subclips=target_subclips, target_subclips_ignore_index=None)
    if mixup_fn is not None:
        if not mixup_backbone:
            feature_dict, tgt, tgt_sub, ignore = mixup_fn(feature_dict, target, target_subclips)
            kwargs.update(target=tgt, target_subclips=tgt_sub, target_subclips_ignore_index=ignore)
        else:
            kwargs["mixup_fn"] = mixup_fn
    outputs, out_tgt = model(feature_dict, **kwargs)
    losses, keeps, metrics = loss_and_accuracy(outputs, out_tgt["target"], out_tgt["target_subclips"],
                                               mixup_enable=mixup_fn is not None,
                                               target_subclips_ignore_index=out_tgt["target_subclips_ignore_index"])
    total, means = reduce_loss(losses, keeps, loss_wts)
    return total, means, metrics
