"""Generate the golden fixtures from the UNMODIFIED reference module (build container only).

    python tests/golden/make_golden.py

For every case it
  1. builds the reference BaseModel from /root/reference (oracle/ref_shim.py stand-ins),
  2. loads the deterministic synthetic weights (afft_b200/synthetic.py, seed 0),
  3. runs the reference forward (test.py:72-82 call pattern) on deterministic synthetic features,
  4. checks oracle/afft_oracle.py against the module in float64 (pin) and float32,
  5. writes tests/golden/<case>.npz (reference fp32 outputs) and records the pin in oracle_pin.json,
     plus the reference's parameter names/shapes (the state-dict key contract, train.py:55-103).
The fixtures and this script are committed; /root/reference is never needed at test time.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from afft_b200 import configs, synthetic  # noqa: E402
from oracle import afft_oracle, ref_shim  # noqa: E402

# (case name, config, B, feature seed, feature family)
CASES = [
    ("egtea_sa_b3", "egtea_sa", 3, 123, "randn"),
    ("ek100_sa_tsn_b2", "ek100_sa_tsn", 2, 123, "randn"),
    ("ek100_sa_tsn_relu_b2", "ek100_sa_tsn", 2, 7, "relu"),
    ("ek100_sa_tsn_wo_audio_b2", "ek100_sa_tsn_wo_audio", 2, 123, "randn"),
    ("ek100_sa_swin_b2", "ek100_sa_swin", 2, 123, "randn"),
    ("ek100_tsa_b2", "ek100_tsa", 2, 123, "randn"),
    ("ek100_ca_b2", "ek100_ca", 2, 123, "randn"),
    ("ek100_sa_wo_token_b2", "ek100_sa_wo_token", 2, 123, "randn"),
    ("egtea_sa_rollout3_b3", "egtea_sa_rollout3", 3, 123, "randn"),
]

# head / mapping variants.  Their outputs are keyed per modality, so the fixtures
# store every leaf as "<outer>|<inner>" (flatten_generic) instead of the fixed "all-fused" layout above.
CASES_N3 = [
    ("ek100_individual_b2", "ek100_individual", 2, 123, "randn"),
    ("ek100_matt_b2", "ek100_matt", 2, 123, "randn"),
    ("ek100_sa_gatedlinear_b2", "ek100_sa_gatedlinear", 2, 123, "randn"),
    ("ek100_sa_nonlinear_b2", "ek100_sa_nonlinear", 2, 123, "randn"),
    ("ek100_sa_linear_ln_b2", "ek100_sa_linear_ln", 2, 123, "randn"),
]


# constructor options without a shipped experiment.  Generic "<outer>|<inner>" layout plus the fuser's
# attention probabilities.
CASES_OPT = [
    ("ek100_sa_3head_b2", "ek100_sa_3head", 2, 123, "randn"),
    ("ek100_sa_modenc_flt_b2", "ek100_sa_modenc_flt", 2, 123, "randn"),
    ("ek100_sa_cross_attn_b2", "ek100_sa_cross_attn", 2, 123, "randn"),
    ("ek100_tsa_mean_b2", "ek100_tsa_mean", 2, 123, "randn"),
    ("ek100_sa_identity_enc_b2", "ek100_sa_identity_enc", 2, 123, "randn"),
    ("egtea_sa_identity_rollout3_b3", "egtea_sa_identity_rollout3", 3, 123, "randn"),
]


def flatten_generic(out):
    flat = {}
    for k, inner in out.items():
        if k in ("attentions", "modality_attns"):
            continue
        for kk, t in inner.items():
            flat[f"{k}|{kk}"] = t
    return flat


def flatten_outputs(out):
    flat = {}
    for k in ("orig_past", "future", "all-fused", "past_futures"):
        flat[k] = out[k]["all-fused"]
    for k in out:
        if k.startswith("logits/") or k.startswith("past_logits/"):
            flat[k] = out[k]["all-fused"]
    ma = out["attentions"]["all-fused"]["modality_attns"]
    flat["modality_attns"] = ma
    return flat


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    only = set(sys.argv[1:])  # optional: regenerate just these cases (merged into oracle_pin.json)
    pin = {}
    names_written = set()
    for case, cfg_name, B, seed, family in CASES:
        if only and case not in only:
            continue
        cfg, T, ncls, _ = configs.named_config(cfg_name)
        model = ref_shim.build_reference_model(cfg, ncls)
        sd = synthetic.synthetic_state_dict(model, seed=0)
        missing, unexpected = model.load_state_dict(sd, strict=False)
        assert not unexpected, unexpected
        assert all(("attn.bias" in m or "masked_bias" in m) for m in missing), missing
        feats6 = synthetic.synthetic_features(cfg["modal_dims"], B, T, seed=seed, family=family, six_d=True)
        feats = {m: t.reshape(B, T, -1) for m, t in feats6.items()}

        ref32 = flatten_outputs(ref_shim.reference_forward(model, feats6))
        model64 = model.double()
        ref64 = flatten_outputs(ref_shim.reference_forward(model64, {m: t.double() for m, t in feats6.items()}))
        model.float()
        sd_full = {k: v for k, v in model.state_dict().items()}
        o64 = flatten_outputs(afft_oracle.forward(sd_full, cfg, ncls, feats, dtype=torch.float64))
        o32 = flatten_outputs(afft_oracle.forward(sd_full, cfg, ncls, feats, dtype=torch.float32))
        rec = {}
        for k in ref64:
            d64 = (o64[k].double() - ref64[k].double()).abs().max().item()
            d32 = (o32[k].double() - ref32[k].double()).abs().max().item()
            r3264 = (ref32[k].double() - ref64[k].double()).abs().max().item()
            rec[k] = {"oracle64_vs_ref64": d64, "oracle32_vs_ref32": d32, "ref32_vs_ref64": r3264,
                      "scale": ref64[k].abs().max().item()}
            assert d64 < 1e-11, (case, k, d64)
            assert d32 < 5e-5, (case, k, d32)
        t5_ref = afft_oracle.top5(ref32["logits/action"][:, 0])
        t5_or = afft_oracle.top5(o32["logits/action"][:, 0])
        rec["top5_identical_oracle32_vs_ref32"] = bool((t5_ref == t5_or).all())
        pin[case] = rec
        print(case, {k: (round(v["oracle64_vs_ref64"], 18), round(v["oracle32_vs_ref32"], 9)) for k, v in rec.items()
                     if isinstance(v, dict)}, flush=True)

        save = {
            "logits": ref32["logits/action"].numpy(),                      # (B, 1, C)
            "past_logits_clip0": ref32["past_logits/action"][:1].numpy(),   # (1, T, C)
            "orig_past": ref32["orig_past"].numpy(),                        # (B, T, D)
            "future": ref32["future"].numpy(),                              # (B, 1, D)
            "past_futures": ref32["past_futures"].numpy(),                  # (B, T, D)
            "modality_attns": ref32["modality_attns"].numpy(),
            "logits64": ref64["logits/action"].numpy(),                     # float64 reference logits
            "top5": t5_ref.numpy(),
        }
        np.savez_compressed(os.path.join(HERE, case + ".npz"), **save)

        if cfg_name not in names_written:
            names_written.add(cfg_name)
            with open(os.path.join(HERE, f"param_names_{cfg_name}.json"), "w") as f:
                json.dump({n: list(p.shape) for n, p in model.named_parameters()}, f, indent=0)

    for case, cfg_name, B, seed, family in CASES_N3 + CASES_OPT:
        if only and case not in only:
            continue
        with_attn = (case, cfg_name, B, seed, family) in CASES_OPT
        cfg, T, ncls, _ = configs.named_config(cfg_name)
        model = ref_shim.build_reference_model(cfg, ncls)
        sd = synthetic.synthetic_state_dict(model, seed=0)
        missing, unexpected = model.load_state_dict(sd, strict=False)
        assert not unexpected, unexpected
        assert all(("attn.bias" in m or "masked_bias" in m) for m in missing), missing
        feats6 = synthetic.synthetic_features(cfg["modal_dims"], B, T, seed=seed, family=family, six_d=True)
        feats = {m: t.reshape(B, T, -1) for m, t in feats6.items()}
        raw32 = ref_shim.reference_forward(model, feats6)
        ref32 = flatten_generic(raw32)
        model64 = model.double()
        raw64 = ref_shim.reference_forward(model64, {m: t.double() for m, t in feats6.items()})
        ref64 = flatten_generic(raw64)
        model.float()
        sd_full = {k: v for k, v in model.state_dict().items()}
        rawo64 = afft_oracle.forward(sd_full, cfg, ncls, feats, dtype=torch.float64)
        o64 = flatten_generic(rawo64)
        o32 = flatten_generic(afft_oracle.forward(sd_full, cfg, ncls, feats, dtype=torch.float32))
        rec, save = {}, {}
        if with_attn:
            ma = raw32["attentions"]["all-fused"]["modality_attns"]
            d = (rawo64["attentions"]["all-fused"]["modality_attns"].double() -
                 raw64["attentions"]["all-fused"]["modality_attns"].double()).abs().max().item()
            assert d < 1e-11, (case, "modality_attns", d)
            rec["modality_attns"] = {"oracle64_vs_ref64": d}
            save["modality_attns"] = ma.numpy()
        for k in ref64:
            d64 = (o64[k].double() - ref64[k].double()).abs().max().item()
            d32 = (o32[k].double() - ref32[k].double()).abs().max().item()
            rec[k] = {"oracle64_vs_ref64": d64, "oracle32_vs_ref32": d32,
                      "ref32_vs_ref64": (ref32[k].double() - ref64[k].double()).abs().max().item(),
                      "scale": ref64[k].abs().max().item()}
            assert d64 < 1e-11, (case, k, d64)
            assert d32 < 5e-5, (case, k, d32)
            if k.startswith("orig_past"):
                continue  # the inputs themselves (individual heads) or covered by past_futures
            if k.startswith("past_logits"):
                save[k] = ref32[k][:1].numpy()  # clip 0 only: keeps the fixtures small
            else:
                save[k] = ref32[k].numpy()
            if k.startswith("logits/"):
                save["logits64|" + k] = ref64[k].numpy()
                save["top5|" + k] = afft_oracle.top5(ref32[k][:, 0]).numpy()
        pin[case] = rec
        print(case, {k: (round(v["oracle64_vs_ref64"], 18), round(v.get("oracle32_vs_ref32", 0.0), 9)) for k, v in rec.items()},
              flush=True)
        np.savez_compressed(os.path.join(HERE, case + ".npz"), **save)
        with open(os.path.join(HERE, f"param_names_{cfg_name}.json"), "w") as f:
            json.dump({n: list(p.shape) for n, p in model.named_parameters()}, f, indent=0)

    pin_path = os.path.join(HERE, "oracle_pin.json")
    if only and os.path.exists(pin_path):
        with open(pin_path) as f:
            old = json.load(f)
        old["pin"].update(pin)
        pin = old["pin"]
    with open(pin_path, "w") as f:
        json.dump({"generated_with": {"torch": torch.__version__, "transformers": __import__("transformers").__version__,
                                      "reference": ref_shim.REFERENCE_ROOT},
                   "cases": {c: list(x) for c, *x in CASES + CASES_N3 + CASES_OPT}, "pin": pin}, f, indent=1)
    print("wrote fixtures to", HERE)


if __name__ == "__main__":
    main()
