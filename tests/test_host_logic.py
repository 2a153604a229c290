"""CPU: host-side mirror of the reference models/ API - module tree, state-dict key contract, config handling,
crop/shape glue and error behaviour.  No native compute."""
import json
import os

import pytest
import torch

from afft_b200 import _capi, configs, synthetic
from afft_b200.models import BaseModel
from afft_b200.models import future_prediction as fp


N3_CONFIGS = ["ek100_individual", "ek100_matt", "ek100_sa_gatedlinear", "ek100_sa_nonlinear", "ek100_sa_linear_ln",
              "ek100_sa_3head", "ek100_sa_modenc_flt", "ek100_sa_cross_attn", "ek100_tsa_mean", "ek100_sa_identity_enc",
              "egtea_sa_identity_rollout3"]


@pytest.mark.parametrize("name", configs.CONFIG_NAMES + N3_CONFIGS)
def test_state_dict_contract_matches_reference(name, golden_dir):
    """Same parameter names and shapes as the reference module (train.py:55-103 init_model contract);
    tests/golden/param_names_*.json was written from the reference's named_parameters()."""
    cfg, T, ncls, _ = configs.named_config(name)
    model = BaseModel(cfg, ncls, {})
    ours = {k: list(v.shape) for k, v in model.named_parameters()}
    ref = json.load(open(os.path.join(golden_dir, f"param_names_{name}.json")))
    assert ours == ref


def test_flops_per_clip_match_survey():
    expect = {"egtea_sa": 3.609, "ek100_sa_tsn": 24.773, "ek100_sa_tsn_wo_audio": 22.055, "ek100_sa_swin": 22.022,
              "ek100_tsa": 13.766, "ek100_ca": 7.223}
    for name, gf in expect.items():
        cfg, T, ncls, _ = configs.named_config(name)
        assert abs(configs.gemm_flops_per_clip(cfg, T, ncls) / 1e9 - gf) < 5e-4


def test_checkpoint_with_gpt2_buffers_loads_non_strict():
    """Checkpoints written with transformers 4.18 carry attn.bias / attn.masked_bias buffers."""
    cfg, T, ncls, _ = configs.named_config("egtea_sa")
    model = BaseModel(cfg, ncls, {})
    sd = synthetic.synthetic_state_dict(model, seed=3)
    sd["future_predictor.future_predictor.gpt_model.h.0.attn.bias"] = torch.ones(1, 1, 1024, 1024)
    sd["future_predictor.future_predictor.gpt_model.h.0.attn.masked_bias"] = torch.tensor(-1e4)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert missing == [] and len(unexpected) == 2
    assert torch.equal(model.state_dict()["future_predictor.dim_encoder.weight"], sd["future_predictor.dim_encoder.weight"])


def test_class_mappings_become_buffers():
    cfg, T, ncls, _ = configs.named_config("egtea_sa")
    m = BaseModel(cfg, ncls, {("action", "verb"): torch.ones(106, 19)})
    assert "cls_map_action_verb" in dict(m.named_buffers())


class _Recorder(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.seen = []

    def forward(self, feats):
        self.seen.append({k: v.clone() for k, v in feats.items()})
        B = next(iter(feats.values())).shape[0]
        x = torch.stack([v.sum(dim=(1, 2)) for v in feats.values()]).sum(0)
        return {"logits/action": {"all-fused": x.reshape(B, 1, 1).expand(B, 1, 4).clone()},
                "attentions": {"all-fused": {"modality_attns": torch.zeros(B), "temporal_attns": {}}}}


def _glue_model():
    cfg, T, ncls, _ = configs.named_config("egtea_sa")
    m = BaseModel(cfg, ncls, {})
    m.future_predictor = _Recorder()
    return m.eval(), T


def test_basemodel_glue_6d_input():
    m, T = _glue_model()
    B = 3
    x = {"rgb": torch.randn(B, T, 1024, 1, 1, 1), "flow": torch.randn(B, T, 1024, 1, 1, 1)}
    out, tgt = m(x, mixup_fn=None, target=None, target_subclips=None, target_subclips_ignore_index=None)
    seen = m.future_predictor.seen[0]
    assert seen["rgb"].shape == (B, T, 1024) and torch.equal(seen["rgb"], x["rgb"].reshape(B, T, 1024))
    assert tgt == {"target": None, "target_subclips": None, "target_subclips_ignore_index": None}
    assert out["logits/action"]["all-fused"].shape == (B, 1, 4)


def test_basemodel_glue_spatial_mean_and_crops():
    m, T = _glue_model()
    B = 2
    # (B, #clips, #crops=3, C, T'=1, H=2, W=2): spatial mean, crops averaged (reference base_model.py:40-46,100-117)
    x = {"rgb": torch.randn(B, T, 3, 1024, 1, 2, 2), "flow": torch.randn(B, T, 1, 1024, 1, 2, 2)}
    out, _ = m(dict(x), mixup_fn=None, target=None, target_subclips=None, target_subclips_ignore_index=None)
    assert len(m.future_predictor.seen) == 3
    exp0 = x["rgb"][:, :, 0].mean(dim=(-1, -2)).permute(0, 1, 3, 2).flatten(1, 2)
    assert torch.allclose(m.future_predictor.seen[0]["rgb"], exp0)
    per_crop = [s["rgb"].sum(dim=(1, 2)) + s["flow"].sum(dim=(1, 2)) for s in m.future_predictor.seen]
    assert torch.allclose(out["logits/action"]["all-fused"][:, 0, 0], torch.stack(per_crop).mean(0), rtol=1e-5, atol=1e-3)
    with pytest.raises(NotImplementedError):
        m({"rgb": torch.zeros(2, 3)})


def test_mixup_hook_is_called():
    m, T = _glue_model()
    calls = []

    def mix(feats, target, target_subclips):
        calls.append(1)
        return feats, "t", "ts", "ig"
    x = {"rgb": torch.randn(1, T, 1024, 1, 1, 1), "flow": torch.randn(1, T, 1024, 1, 1, 1)}
    _, tgt = m(x, mixup_fn=mix, target=None, target_subclips=None, target_subclips_ignore_index=None)
    assert calls and tgt == {"target": "t", "target_subclips": "ts", "target_subclips_ignore_index": "ig"}


def test_error_behaviour():
    cfg, T, ncls, _ = configs.named_config("egtea_sa")
    model = BaseModel(cfg, ncls, {})
    x = {"rgb": torch.zeros(1, T, 1024, 1, 1, 1), "flow": torch.zeros(1, T, 1024, 1, 1, 1)}
    with pytest.raises(_capi.AfftError):  # CPU tensors in training mode: there is no CPU fallback
        model.train()(dict(x))
    with pytest.raises(_capi.AfftError):  # CPU tensors: there is no CPU fallback
        model.eval()(dict(x))
    with pytest.raises(ValueError):
        model.eval()({"rgb": torch.zeros(1, T, 1024, 1, 1, 1)})
    bad = configs.named_config("egtea_sa")[0]
    bad["common"]["fp_output_len"] = 0
    with pytest.raises(ValueError):
        BaseModel(bad, ncls, {})
    bad = configs.named_config("ek100_matt")[0]
    bad["common"]["fp_output_len"] = 2  # one attention row cannot weight several future steps
    with pytest.raises(NotImplementedError):
        BaseModel(bad, {"action": 3806}, {})
    bad = configs.named_config("ek100_individual")[0]
    bad["common"]["fusion_cls"] = True  # reference future_prediction.py:194
    with pytest.raises(AssertionError):
        BaseModel(bad, {"action": 3806}, {})
    gated = BaseModel(configs.named_config("ek100_sa_gatedlinear")[0], {"action": 3806}, {})
    with pytest.raises(_capi.AfftError):  # ablation mappings run as library kernels too: no CPU path
        gated.eval()({m: torch.zeros(1, 10, d, 1, 1, 1) for m, d in (("rgb", 1024), ("objects", 352), ("flow", 1024))})


def test_instantiate_accepts_attribute_configs():
    class NS(dict):
        __getattr__ = dict.__getitem__
    lin = fp.instantiate(NS(_target_="models.feature_mapping.Linear", use_layernorm=False, sparse_mapping=True),
                         in_features=352, out_features=1024)
    assert lin.mapping[0].weight.shape == (1024, 352)
    ident = fp.instantiate({"_target_": "models.feature_mapping.Linear"}, in_features=1024, out_features=1024)
    assert isinstance(ident.mapping[0], torch.nn.Identity)
    assert isinstance(fp.instantiate({"_target_": "torch.nn.Identity"}), torch.nn.Identity)


def test_synthetic_weights_are_deterministic():
    cfg, T, ncls, _ = configs.named_config("egtea_sa")
    m = BaseModel(cfg, ncls, {})
    a = synthetic.synthetic_state_dict(m, seed=0)
    b = synthetic.synthetic_state_dict(m, seed=0)
    assert all(torch.equal(a[k], b[k]) for k in a)
    f1 = synthetic.synthetic_features(cfg["modal_dims"], 2, T, seed=9)
    f2 = synthetic.synthetic_features(cfg["modal_dims"], 2, T, seed=9)
    assert all(torch.equal(f1[k], f2[k]) for k in f1)


def test_splitk_plan_cost_model():
    """afft_plan_ksplit (host arithmetic of the split-K scheduler): no split when the tiles fill the GPU,
    a split when few tiles carry a long K loop, never an empty split, never beyond the cap."""
    from afft_b200 import _capi
    lib = _capi.lib()
    plan = lambda tiles, slots, kb, ctas=1, cap=4: lib.afft_plan_ksplit(tiles, slots, kb, ctas, cap)  # noqa: E731
    # headline batch (B = 256): every GEMM has >= 1 wave of tiles -> unsplit (bitwise identical to the unsplit build)
    for tiles, kb in ((360, 16), (360, 64), (1080, 16), (1440, 16), (144, 32), (144, 128), (285, 16), (72, 32)):
        assert plan(tiles, 74, kb, ctas=2) == 1, (tiles, kb)
    # batch 1: 8 tiles with K = 4096 (64 K blocks) on 148 CTAs -> split to the cap; cap 1 switches it off
    assert plan(8, 148, 64) == 4
    assert plan(8, 148, 64, cap=1) == 1
    assert plan(8, 148, 64, cap=16) >= 4
    # batch 32, GPT-2 mlp c_proj: 80 tiles, K = 8192: 3 splits (2 waves of 43 K blocks instead of 1 of 128)
    assert plan(80, 148, 128) == 3
    # full single wave: nothing to gain
    assert plan(144, 148, 64) == 1
    # a split never leaves a K range empty and never exceeds num_kb / 2
    for kb in range(1, 40):
        for tiles in (1, 3, 8, 24):
            s = plan(tiles, 148, kb, cap=16)
            assert 1 <= s <= max(1, kb // 2)
            per = -(-kb // s)
            assert (s - 1) * per < kb


def test_workspace_size_query_is_host_arithmetic():
    """afft_workspace_bytes_for: the caller-owned-workspace contract of SURVEY 8b starts with a size query that needs no
    device; it grows with max_batch and with the strict mode's hi/lo pairs, and afft_create_in refuses a short buffer."""
    import ctypes as C
    from afft_b200 import _capi
    lib = _capi.lib()

    def need(max_batch, precision=0, stages=0):
        cfg = _capi.Config()
        cfg.fuser_kind, cfg.T, cfg.n_mod = 0, 18, 4
        for i, (n, d) in enumerate([("rgb", 1024), ("objects", 352), ("audio", 1024), ("flow", 1024)]):
            cfg.mod_name[i].value = n.encode()
            cfg.mod_dim[i] = d
        cfg.dim, cfg.fuser_depth, cfg.fuser_heads, cfg.norm_elementwise = 1024, 6, 4, 1
        cfg.gpt_dim, cfg.gpt_layers, cfg.gpt_heads = 2048, 6, 4
        cfg.n_cls = 1
        cfg.cls_name[0].value = b"action"
        cfg.cls_dim[0] = 3806
        cfg.precision, cfg.max_batch, cfg.device, cfg.fp_output_len, cfg.stages = precision, max_batch, 0, 1, stages
        n = C.c_size_t()
        assert lib.afft_workspace_bytes_for(C.byref(cfg), C.byref(n)) == 0
        return n.value, cfg

    b32, _ = need(32)
    b256, cfg = need(256)
    assert 0 < b32 < b256 < 2 * 1024 ** 3
    assert need(256, precision=1)[0] > b256            # strict: hi/lo pairs and fp32 q|k|v
    assert need(256, precision=2)[0] == b256           # fp16 operands take the room of bf16 operands
    assert need(256, stages=1)[0] < b256               # fuser-only handle: no GPT-2 buffers
    h = C.c_void_p()
    assert lib.afft_create_in(C.byref(cfg), None, 0, None, C.byref(h)) != 0  # no buffer: refused before any device work...
