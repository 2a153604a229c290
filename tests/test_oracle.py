"""CPU: the oracle (oracle/afft_oracle.py) against the golden fixtures written from the reference module
(tests/golden/make_golden.py), plus - when /root/reference is present - against the live reference."""
import os

import numpy as np
import pytest
import torch

from afft_b200 import configs, synthetic
from afft_b200.models import BaseModel
from oracle import afft_oracle, ref_shim

# oracle fp32 vs reference-module fp32: both are fp32 evaluations of the same formulae; they differ only by
# summation order inside matmul/softmax (measured <= 5.5e-6 at generation time).
TOL_F32 = 1e-4
# oracle fp64 vs reference-module fp64 logits
TOL_F64 = 1e-11

_state = {}


def _weights(cfg_name):
    if cfg_name not in _state:
        cfg, T, ncls, _ = configs.named_config(cfg_name)
        model = BaseModel(cfg, ncls, {})
        _state.clear()  # keep one ~1.5 GB state dict alive at a time
        _state[cfg_name] = synthetic.synthetic_state_dict(model, seed=0)
    return _state[cfg_name]


CASES = ["egtea_sa_b3", "ek100_sa_tsn_b2", "ek100_sa_tsn_relu_b2", "ek100_sa_tsn_wo_audio_b2", "ek100_sa_swin_b2",
         "ek100_tsa_b2", "ek100_ca_b2", "ek100_sa_wo_token_b2", "egtea_sa_rollout3_b3"]


@pytest.mark.parametrize("case", CASES)
def test_oracle_matches_golden(case, golden_dir, golden_cases):
    cfg_name, B, seed, family = golden_cases[case]
    cfg, T, ncls, _ = configs.named_config(cfg_name)
    sd = _weights(cfg_name)
    feats = synthetic.synthetic_features(cfg["modal_dims"], B, T, seed=seed, family=family)
    gold = np.load(os.path.join(golden_dir, case + ".npz"))
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    out = afft_oracle.forward(sd, cfg, ncls, feats, dtype=torch.float32)
    assert np.abs(out["logits/action"]["all-fused"].numpy() - gold["logits"]).max() < TOL_F32
    assert np.abs(out["past_logits/action"]["all-fused"][:1].numpy() - gold["past_logits_clip0"]).max() < TOL_F32
    assert np.abs(out["orig_past"]["all-fused"].numpy() - gold["orig_past"]).max() < TOL_F32
    assert np.abs(out["future"]["all-fused"].numpy() - gold["future"]).max() < TOL_F32
    assert np.abs(out["past_futures"]["all-fused"].numpy() - gold["past_futures"]).max() < TOL_F32
    if gold["modality_attns"].ndim > 1:
        ma = out["attentions"]["all-fused"]["modality_attns"].numpy()
        assert ma.shape == gold["modality_attns"].shape
        assert np.abs(ma - gold["modality_attns"]).max() < 1e-5
    assert (afft_oracle.top5(out["logits/action"]["all-fused"][:, 0]).numpy() == gold["top5"]).all()
    if case in ("egtea_sa_b3", "ek100_sa_tsn_b2"):
        o64 = afft_oracle.forward(sd, cfg, ncls, feats, dtype=torch.float64)
        assert np.abs(o64["logits/action"]["all-fused"].numpy() - gold["logits64"]).max() < TOL_F64


CASES_N3 = ["ek100_individual_b2", "ek100_matt_b2", "ek100_sa_gatedlinear_b2", "ek100_sa_nonlinear_b2",
            "ek100_sa_linear_ln_b2"]
# constructor options no shipped experiment switches on: three heads, SA modal_encoding + frame_level_token,
# cross_attn=True, T-SA without frame-level token (same fixture layout + the fuser's attention probabilities)
CASES_OPT = ["ek100_sa_3head_b2", "ek100_sa_modenc_flt_b2", "ek100_sa_cross_attn_b2", "ek100_tsa_mean_b2",
             "ek100_sa_identity_enc_b2", "egtea_sa_identity_rollout3_b3"]  # + Identity dim_encoder / dim_decoder (+ roll-out)


@pytest.mark.parametrize("case", CASES_N3 + CASES_OPT)
def test_oracle_matches_golden_head_and_mapping_variants(case, golden_dir, golden_cases):
    """SURVEY 8f row N3 (IndividualFuturePrediction, CMFPScoreFusion + MATT, GatedLinear / NonLinear / layer-normed
    mappings): fixtures hold every output leaf of the reference module as "<outer>|<inner>"."""
    cfg_name, B, seed, family = golden_cases[case]
    cfg, T, ncls, _ = configs.named_config(cfg_name)
    sd = _weights(cfg_name)
    feats = synthetic.synthetic_features(cfg["modal_dims"], B, T, seed=seed, family=family)
    gold = np.load(os.path.join(golden_dir, case + ".npz"))
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    out = afft_oracle.forward(sd, cfg, ncls, feats, dtype=torch.float32)
    checked = 0
    for key in gold.files:
        if key.startswith(("logits64|", "top5|")):
            continue
        if key == "modality_attns":
            ma = out["attentions"]["all-fused"]["modality_attns"].numpy()
            assert ma.shape == gold[key].shape
            assert np.abs(ma - gold[key]).max() < 1e-5
            continue
        outer, inner = key.split("|")
        mine = out[outer][inner]
        if outer.startswith("past_logits"):
            mine = mine[:1]
        assert np.abs(mine.numpy() - gold[key]).max() < TOL_F32, key
        if outer.startswith("logits/"):
            assert (afft_oracle.top5(out[outer][inner][:, 0]).numpy() == gold["top5|" + key]).all()
        checked += 1
    assert checked >= 5


def test_oracle_output_contract():
    """Keys and shapes of CMFPEarly.forward (reference models/future_prediction.py:282-291)."""
    cfg, T, ncls, _ = configs.named_config("egtea_sa")
    sd = _weights("egtea_sa")
    B = 2
    feats = synthetic.synthetic_features(cfg["modal_dims"], B, T, seed=5)
    out = afft_oracle.forward(sd, cfg, ncls, feats)
    assert set(out) == {"orig_past", "future", "all-fused", "past_futures", "past_logits/action", "logits/action",
                        "attentions"}
    assert out["orig_past"]["all-fused"].shape == (B, T, 1024)
    assert out["future"]["all-fused"].shape == (B, 1, 1024)
    assert out["all-fused"]["all-fused"].shape == (B, 1, 1024)
    assert out["past_futures"]["all-fused"].shape == (B, T, 1024)
    assert out["past_logits/action"]["all-fused"].shape == (B, T, 106)
    assert out["logits/action"]["all-fused"].shape == (B, 1, 106)
    assert out["attentions"]["all-fused"]["modality_attns"].shape == (B, 2, T, 4, 3, 3)
    # past_futures = cat(z[:, :1], z_hat[:, :T-1]); all-fused is the fused last step (not the prediction)
    assert torch.equal(out["past_futures"]["all-fused"][:, 0], out["orig_past"]["all-fused"][:, 0])
    assert torch.equal(out["all-fused"]["all-fused"][:, 0], out["orig_past"]["all-fused"][:, T - 1])
    # causality of the predictor: past logits at step t must not depend on features after t
    feats2 = {m: f.clone() for m, f in feats.items()}
    g = torch.Generator().manual_seed(1)
    for f in feats2.values():
        f[:, T - 1] = torch.randn(f[:, T - 1].shape, generator=g)
    out2 = afft_oracle.forward(sd, cfg, ncls, feats2)
    assert torch.allclose(out2["past_logits/action"]["all-fused"][:, :T - 1], out["past_logits/action"]["all-fused"][:, :T - 1],
                          atol=1e-6)
    assert not torch.allclose(out2["logits/action"]["all-fused"], out["logits/action"]["all-fused"], atol=1e-4)


@pytest.mark.skipif(not ref_shim.available(), reason="/root/reference not present (GPU box)")
def test_oracle_matches_live_reference():
    """Pin: the restatement against the unmodified reference module, float64, fresh random weights."""
    cfg, T, ncls, _ = configs.named_config("egtea_sa")
    torch.manual_seed(11)
    model = ref_shim.build_reference_model(cfg, ncls).double()
    B = 2
    feats6 = {m: t.double() for m, t in synthetic.synthetic_features(cfg["modal_dims"], B, T, seed=3, six_d=True).items()}
    ref = ref_shim.reference_forward(model, feats6)
    out = afft_oracle.forward(model.state_dict(), cfg, ncls, {m: t.reshape(B, T, -1) for m, t in feats6.items()},
                              dtype=torch.float64)
    for k in ("logits/action", "past_logits/action", "orig_past", "future", "past_futures", "all-fused"):
        assert (out[k]["all-fused"] - ref[k]["all-fused"]).abs().max().item() < 1e-12, k
    assert (out["attentions"]["all-fused"]["modality_attns"] - ref["attentions"]["all-fused"]["modality_attns"]).abs().max() < 1e-13


def test_aten_spelling_agrees_with_elementary_spelling():
    """bench.py times the oracle with ATEN_OPS=True (the reference's own library calls); same numbers."""
    cfg, T, ncls, _ = configs.named_config("egtea_sa")
    sd = _weights("egtea_sa")
    feats = synthetic.synthetic_features(cfg["modal_dims"], 2, T, seed=8)
    a = afft_oracle.forward(sd, cfg, ncls, feats)
    afft_oracle.ATEN_OPS = True
    try:
        b = afft_oracle.forward(sd, cfg, ncls, feats)
    finally:
        afft_oracle.ATEN_OPS = False
    for k in ("logits/action", "past_logits/action", "orig_past", "past_futures"):
        assert (a[k]["all-fused"] - b[k]["all-fused"]).abs().max().item() < 2e-5
