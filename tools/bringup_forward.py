"""GPU bring-up of the end-to-end path: every golden case, fast (bf16) and strict (bf16x3) modes,
compared with the reference module's fp32 outputs stored in tests/golden/*.npz.
Usage on the GPU box:  python tools/bringup_forward.py [case ...]
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from afft_b200 import configs, synthetic  # noqa: E402
from afft_b200.models import BaseModel  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def run_case(case, cfg_name, B, seed, family, strict):
    cfg, T, ncls, _ = configs.named_config(cfg_name)
    dev = torch.device("cuda:0")
    model = BaseModel(cfg, ncls, {}, strict=strict)
    model.load_state_dict(synthetic.synthetic_state_dict(model, seed=0))
    model = model.to(dev).eval()
    feats = synthetic.synthetic_features(cfg["modal_dims"], B, T, seed=seed, family=family, six_d=True)
    feats = {m: t.to(dev) for m, t in feats.items()}
    with torch.no_grad():
        out, _ = model(feats, mixup_fn=None, target=None, target_subclips=None, target_subclips_ignore_index=None)
    torch.cuda.synchronize()
    gold = np.load(os.path.join(GOLDEN, case + ".npz"))
    rec = {"case": case, "strict": strict, "launches": model.future_predictor.last_launch_count()}

    def cmp(name, ours, ref):
        ours = ours.float().cpu().numpy()
        rec[name] = float(np.abs(ours - ref).max())
        rec[name + "_nan"] = int(np.isnan(ours).sum())

    cmp("logits", out["logits/action"]["all-fused"], gold["logits"])
    cmp("past_logits0", out["past_logits/action"]["all-fused"][:1], gold["past_logits_clip0"])
    cmp("orig_past", out["orig_past"]["all-fused"], gold["orig_past"])
    cmp("future", out["future"]["all-fused"], gold["future"])
    cmp("past_futures", out["past_futures"]["all-fused"], gold["past_futures"])
    ma = out["attentions"]["all-fused"]["modality_attns"]
    if gold["modality_attns"].ndim > 1:
        cmp("attn", ma, gold["modality_attns"])
    t5 = out["logits/action"]["all-fused"][:, 0].topk(5, dim=-1).indices.cpu().numpy()
    rec["top5_same"] = bool((t5 == gold["top5"]).all())
    rec["logit_scale"] = float(np.abs(gold["logits"]).max())
    return rec


def main():
    pin = json.load(open(os.path.join(GOLDEN, "oracle_pin.json")))["cases"]
    names = sys.argv[1:] or list(pin.keys())
    bad = 0
    for case in names:
        cfg_name, B, seed, family = pin[case]
        for strict in (False, True):
            t0 = time.time()
            try:
                rec = run_case(case, cfg_name, B, seed, family, strict)
            except Exception as e:  # noqa: BLE001
                rec = {"case": case, "strict": strict, "error": repr(e)}
                bad += 1
            rec["sec"] = round(time.time() - t0, 1)
            print(json.dumps(rec), flush=True)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
