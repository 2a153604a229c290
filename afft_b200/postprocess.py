"""Logit post-processing on the GPU: softmax -> verb/noun marginalisation -> top-k.

Replaces the CPU round trip of reference challenge.py:196-210 (`marginalize_verb_noun`: scipy softmax + two
matmuls with the 0/1 `class_mappings` matrices of datasets/epic_kitchens.py:87-106) and the argsort ranking of
common/utils.py:19-42.  One kernel, one CTA per clip, no fallback.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch

from . import _capi


class VerbNounMarginalizer:
    def __init__(self, verb_in_action: torch.Tensor, noun_in_action: torch.Tensor, device="cuda"):
        """verb_in_action (A, V), noun_in_action (A, N): the reference's class_mappings[('verb','action')] /
        [('noun','action')] - exactly one 1 per action row."""
        for m in (verb_in_action, noun_in_action):
            if m.dim() != 2 or not bool(((m == 0) | (m == 1)).all()) or not bool((m.sum(1) == 1).all()):
                raise ValueError("class mapping must be a 0/1 matrix with exactly one 1 per action row")
        if verb_in_action.shape[0] != noun_in_action.shape[0]:
            raise ValueError("verb and noun mappings disagree on the number of actions")
        self.A, self.n_verb = verb_in_action.shape
        self.n_noun = noun_in_action.shape[1]
        self.device = torch.device(device)
        self.verb_of = verb_in_action.argmax(1).to(torch.int32).to(self.device).contiguous()
        self.noun_of = noun_in_action.argmax(1).to(torch.int32).to(self.device).contiguous()

    def __call__(self, logits: torch.Tensor, k: int = 5, want_probs: bool = False) -> Dict[str, torch.Tensor]:
        """logits (B, A) fp32 CUDA (may be a strided view of the padded logits buffer).
        Returns verb (B, V) / noun (B, N) scores, action = the raw logits (as the reference does), optional action
        probabilities, and topk (B, 3, k) int32 indices ordered action / verb / noun."""
        if logits.dim() != 2 or logits.shape[1] != self.A or logits.dtype != torch.float32 or logits.stride(1) != 1:
            raise _capi.AfftError("logits must be (B, A) fp32 with unit inner stride")
        B = logits.shape[0]
        dev = logits.device
        verb = torch.empty(B, self.n_verb, device=dev)
        noun = torch.empty(B, self.n_noun, device=dev)
        probs = torch.empty(B, self.A, device=dev) if want_probs else None
        topk = torch.empty(B, 3, k, device=dev, dtype=torch.int32)
        _capi.check(_capi.lib().afft_marginalize_topk(
            _capi.ptr(logits), logits.stride(0), B, self.A, _capi.ptr(self.verb_of), _capi.ptr(self.noun_of), self.n_verb,
            self.n_noun, _capi.ptr(probs), _capi.ptr(verb), _capi.ptr(noun), _capi.ptr(topk), k,
            _capi.current_stream_ptr(dev)))
        out = {"verb": verb, "noun": noun, "action": logits, "topk": topk}
        if want_probs:
            out["action_probs"] = probs
        return out


def marginalize_verb_noun(logits: torch.Tensor, class_mappings: Dict[Tuple[str, str], torch.Tensor], k: int = 5):
    """Functional form with the reference's `dataset.class_mappings` dict."""
    m = VerbNounMarginalizer(class_mappings[("verb", "action")], class_mappings[("noun", "action")], logits.device)
    return m(logits, k=k)
