"""Golden fixture for the logit post-processing row, generated with the reference's own
arithmetic (build container only): class mappings built as datasets/epic_kitchens.py:87-106 does from
annotations/ek100_rulstm/actions.csv, scores as challenge.py:196-210 (scipy softmax + np.matmul), ranking as
common/utils.py:19-42 (argsort()[:, ::-1]).  Writes tests/golden/postprocess_ek100.npz."""
import csv
import os

import numpy as np
import torch
from scipy.special import softmax

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("AFFT_REFERENCE_ROOT", "/root/reference")

verb_noun_to_action = {}
with open(os.path.join(REF, "annotations", "ek100_rulstm", "actions.csv")) as f:  # datasets/epic_kitchens.py:156-165
    for line in csv.DictReader(f, delimiter=","):
        verb_noun_to_action[(int(line["verb"]), int(line["noun"]))] = int(line["id"])
num_actions, num_verbs, num_nouns = 3806, 97, 300
assert len(verb_noun_to_action) == num_actions
verb_in_action = torch.zeros((num_actions, num_verbs), dtype=torch.float)
noun_in_action = torch.zeros((num_actions, num_nouns), dtype=torch.float)
for (verb, noun), action in verb_noun_to_action.items():
    verb_in_action[action, verb] = 1.0
    noun_in_action[action, noun] = 1.0

g = torch.Generator().manual_seed(2024)
logits = (torch.randn(6, num_actions, generator=g) * 1.5).numpy()
probs = softmax(logits, axis=-1)
res_verb = np.matmul(probs, verb_in_action).numpy()
res_noun = np.matmul(probs, noun_in_action).numpy()
top5 = np.stack([x.argsort()[:, ::-1][:, :5] for x in (logits, res_verb, res_noun)], axis=1)
np.savez_compressed(os.path.join(HERE, "postprocess_ek100.npz"),
                    verb_of=verb_in_action.argmax(1).numpy().astype(np.int16),
                    noun_of=noun_in_action.argmax(1).numpy().astype(np.int16),
                    logits=logits, verb=res_verb, noun=res_noun, top5=top5.astype(np.int32))
print("wrote postprocess_ek100.npz", res_verb.shape, res_noun.shape, top5.shape)
