"""The model configurations of the reference's experiment files, as plain nested dicts with the same
key structure Hydra would hand to ``BaseModel`` (``cfg.model``; reference conf/config.yaml:39-43,
conf/model/**, expts/*.txt).  ``_target_`` strings are the reference's own, so the same dict builds
either the reference module (through oracle/ref_shim.py) or this package's drop-in classes.
"""
from __future__ import annotations

import copy
from typing import Dict

MODAL_FEATURE_ORDER = ["rgb", "objects", "audio", "poses", "flow"]  # conf/config.yaml:41

_FUSER_TARGETS = {
    "SA-Fuser": "models.fusion.ModalTokenCMFuser",          # conf/model/fuser/SA-Fuser.yaml
    "SA-Fuser_wo_token": "models.fusion.CMFuser",           # conf/model/fuser/SA-Fuser_wo_token.yaml
    "T-SA-Fuser": "models.fusion.TemporalCMFuser",          # conf/model/fuser/T-SA-Fuser.yaml
    "CA-Fuser": "models.fusion.TemporalCrossAttentFuser",   # conf/model/fuser/CA-Fuser.yaml
}


def model_cfg(modal_dims: Dict[str, int], *, fuser: str = "SA-Fuser", depth: int = 6, num_heads: int = 4,
              fp_layers: int = 6, fp_heads: int = 4, fp_inter_dim: int = 2048, common_dim: int = 1024,
              dropout: float = 0.2, fuser_kwargs: Dict = None) -> Dict:
    """cfg.model for an early-fusion (CMFPEarly) experiment."""
    drop = dict(embd_drop_rate=0.1, drop_rate=0.1, attn_drop_rate=0.1, drop_path_rate=0.1)
    if fuser == "SA-Fuser":
        fcfg = dict(dim=common_dim, depth=depth, num_heads=num_heads, **drop, cross_attn=False, norm_elementwise=True,
                    modalities=dict(modal_dims), modal_encoding=False, frame_level_token=False,
                    temporal_sequence_length=None)
    elif fuser == "SA-Fuser_wo_token":
        fcfg = dict(dim=common_dim, depth=depth, num_heads=num_heads, **drop, cross_attn=False)
    elif fuser == "T-SA-Fuser":
        fcfg = dict(dim=common_dim, depth=depth, num_heads=num_heads, **drop, modalities=dict(modal_dims),
                    modal_encoding=True, frame_level_token=False, temporal_sequence_length=None)
    elif fuser == "CA-Fuser":
        fcfg = dict(dim=common_dim, modalities=dict(modal_dims), num_heads=num_heads, **drop)
    else:
        raise ValueError(f"unknown fuser {fuser}")
    fcfg.update(fuser_kwargs or {})
    fcfg["_target_"] = _FUSER_TARGETS[fuser]
    return {
        "modal_dims": dict(modal_dims),
        "modal_feature_order": list(MODAL_FEATURE_ORDER),
        "common_dim": common_dim,
        "dropout": dropout,
        "common": {  # conf/model/common.yaml
            "in_features": common_dim,
            "share_classifiers": True, "share_predictors": True, "modality_cls": False, "fusion_cls": True,
            "backbones": {m: {"_target_": "torch.nn.Identity"} for m in modal_dims},
            "fp_output_len": 1, "fp_inter_dim": fp_inter_dim, "fp_layers": fp_layers, "fp_heads": fp_heads,
            "fp_output_attentions": False, "embd_pdrop": 0.1, "resid_pdrop": 0.1, "attn_pdrop": 0.1,
        },
        "backbone": {"_target_": "torch.nn.Identity"},
        "future_predictor": {  # conf/model/future_predictor/base_future_predictor.yaml
            "_target_": "models.future_prediction.BaseFuturePredictor",
            "in_features": common_dim, "inter_dim": fp_inter_dim, "n_layer": fp_layers, "n_head": fp_heads,
            "output_attentions": False, "embd_pdrop": 0.1, "resid_pdrop": 0.1, "attn_pdrop": 0.1,
        },
        "fuser": fcfg,
        "CMFP": {"_target_": "models.future_prediction.CMFPEarly", "model_cfg": None},
        "mapping": {"_target_": "models.feature_mapping.Linear", "use_layernorm": False, "sparse_mapping": True},
    }


_MAPPINGS = {  # conf/model/mapping/*.yaml
    "linear": {"_target_": "models.feature_mapping.Linear", "use_layernorm": False, "sparse_mapping": True},
    "gatedlinear": {"_target_": "models.feature_mapping.GatedLinear", "use_layernorm": True},
    "nonlinear": {"_target_": "models.feature_mapping.NonLinear", "use_layernorm": True, "activation": "relu"},
}


def _with_mapping(cfg: Dict, name: str, **overrides) -> Dict:
    cfg["mapping"] = dict(_MAPPINGS[name], **overrides)
    return cfg


def _unimodal_heads(cfg: Dict, head: str) -> Dict:
    """expts/00 (model/CMFP=individual) and expts/05 (model/CMFP=scorefusion, model/fuser=MATT): per-modality
    predictors and classifiers, no fused classifier."""
    cfg["common"].update(share_classifiers=False, share_predictors=False, modality_cls=True, fusion_cls=False)
    if head == "individual":
        cfg["CMFP"] = {"_target_": "models.future_prediction.IndividualFuturePrediction", "model_cfg": None}
    else:
        cfg["CMFP"] = {"_target_": "models.future_prediction.CMFPScoreFusion", "model_cfg": None}
        cfg["fuser"] = {"_target_": "models.fusion.MATT", "modal_dims": dict(cfg["modal_dims"]),
                        "dim": cfg["common"]["in_features"], "drop_rate": 0.8}  # conf/model/fuser/MATT.yaml
    return cfg


def _with_output_len(cfg: Dict, n: int) -> Dict:
    cfg["common"]["fp_output_len"] = n
    return cfg


# name -> (cfg.model, T, num_classes, eval batch size of the experiment file)
def named_config(name: str):
    ek4 = {"rgb": 1024, "objects": 352, "audio": 1024, "flow": 1024}
    ek3 = {"rgb": 1024, "objects": 352, "flow": 1024}
    table = {
        # expts/06_SA-Fuser_egtea_val.txt
        "egtea_sa": (lambda: model_cfg({"rgb": 1024, "flow": 1024}, depth=2, fp_layers=2), 10, {"action": 106}, 32),
        # expts/01_SA-Fuser_ek100_val_TSN.txt  (north-star headline: R-TSN+O+AU+F, 4h_18s)
        "ek100_sa_tsn": (lambda: model_cfg(ek4), 18, {"action": 3806}, 32),
        # expts/01_SA-Fuser_ek100_val_TSN_wo_audio.txt
        "ek100_sa_tsn_wo_audio": (lambda: model_cfg(ek3), 18, {"action": 3806}, 32),
        # expts/01_SA-Fuser_ek100_val_Swin.txt (4h_16s)
        "ek100_sa_swin": (lambda: model_cfg(ek4), 16, {"action": 3806}, 32),
        # expts/03_T-SA-Fuser_ek100_train.txt
        "ek100_tsa": (lambda: model_cfg(ek4, fuser="T-SA-Fuser", fuser_kwargs=dict(
            modal_encoding=True, frame_level_token=True, temporal_sequence_length=10)), 10, {"action": 3806}, 16),
        # expts/04_CA-Fuser_ek100_train.txt
        "ek100_ca": (lambda: model_cfg(ek4, fuser="CA-Fuser"), 10, {"action": 3806}, 16),
        # expts/02_SA-Fuser_wo_token_ek100_train.txt
        "ek100_sa_wo_token": (lambda: model_cfg(ek4, fuser="SA-Fuser_wo_token"), 10, {"action": 3806}, 16),
        # model.common.fp_output_len=3 on the EGTEA model: autoregressive roll-out (future_prediction.py:395-412)
        "egtea_sa_rollout3": (lambda: _with_output_len(model_cfg({"rgb": 1024, "flow": 1024}, depth=2, fp_layers=2), 3),
                              10, {"action": 106}, 32),
        # expts/00_RGB_TSN_ek100_train.txt (model/CMFP=individual), here with a second, 352-wide modality
        "ek100_individual": (lambda: _unimodal_heads(model_cfg({"rgb": 1024, "objects": 352}, fp_layers=2), "individual"),
                             10, {"action": 3806}, 16),
        # expts/05_MATT_ek100_train.txt (model/CMFP=scorefusion, model/fuser=MATT, fp_layers=2)
        "ek100_matt": (lambda: _unimodal_heads(model_cfg(ek4, fp_layers=2), "scorefusion"), 10, {"action": 3806}, 16),
        # conf/model/mapping/gatedlinear.yaml and nonlinear.yaml on a shallow SA-Fuser model
        "ek100_sa_gatedlinear": (lambda: _with_mapping(model_cfg(ek3, depth=2, fp_layers=2), "gatedlinear"),
                                 10, {"action": 3806}, 16),
        "ek100_sa_nonlinear": (lambda: _with_mapping(model_cfg(ek3, depth=2, fp_layers=2), "nonlinear"),
                               10, {"action": 3806}, 16),
        # Linear mapping with use_layernorm=true, sparse_mapping=false (feature_mapping.py:58-67)
        "ek100_sa_linear_ln": (lambda: _with_mapping(model_cfg(ek3, depth=2, fp_layers=2), "linear", use_layernorm=True,
                                                     sparse_mapping=False), 10, {"action": 3806}, 16),
        # options of the fusers no shipped experiment switches on but the constructors accept:
        # three classifier heads (future_prediction.py:97-122,144-153 with num_classes = {action, verb, noun})
        "ek100_sa_3head": (lambda: model_cfg(ek3, depth=2, fp_layers=2), 10, {"action": 3806, "verb": 97, "noun": 300}, 16),
        # SA-Fuser with modality embedding and a frame-level modality token (fusion.py:300-317,338-353)
        "ek100_sa_modenc_flt": (lambda: model_cfg(ek3, depth=2, fp_layers=2, fuser_kwargs=dict(
            modal_encoding=True, frame_level_token=True, temporal_sequence_length=10)), 10, {"action": 3806}, 16),
        # SA-Fuser with cross_attn=True: a token does not attend to itself (fusion.py:330-335)
        "ek100_sa_cross_attn": (lambda: model_cfg(ek3, depth=2, fp_layers=2, fuser_kwargs=dict(cross_attn=True)),
                                10, {"action": 3806}, 16),
        # common_dim == fp_inter_dim: dim_encoder / dim_decoder are nn.Identity (future_prediction.py:245-255); GPT-2 width 1024
        "ek100_sa_identity_enc": (lambda: model_cfg(ek3, depth=2, fp_layers=2, fp_inter_dim=1024), 10, {"action": 3806}, 16),
        "egtea_sa_identity_rollout3": (lambda: _with_output_len(model_cfg({"rgb": 1024, "flow": 1024}, depth=2, fp_layers=2,
                                                                         fp_inter_dim=1024), 3), 10, {"action": 106}, 32),
        # T-SA-Fuser without frame-level token: per-timestep mean over the modalities (fusion.py:211-214)
        "ek100_tsa_mean": (lambda: model_cfg(ek4, fuser="T-SA-Fuser", depth=2, fp_layers=2, fuser_kwargs=dict(
            modal_encoding=True, frame_level_token=False, temporal_sequence_length=None)), 10, {"action": 3806}, 16),
    }
    if name not in table:
        raise KeyError(f"unknown config {name}; have {sorted(table)}")
    fn, T, ncls, bs = table[name]
    return copy.deepcopy(fn()), T, dict(ncls), bs


CONFIG_NAMES = ["egtea_sa", "ek100_sa_tsn", "ek100_sa_tsn_wo_audio", "ek100_sa_swin", "ek100_tsa", "ek100_ca",
                "ek100_sa_wo_token"]


def gemm_flops_per_clip(cfg: Dict, T: int, num_classes: Dict[str, int]) -> float:
    """2*M*K*N of every linear the reference executes for one clip: the
    algorithmic work behind roofline.achieved.  Attention, LayerNorm and GELU are not counted."""
    D = cfg["common_dim"]
    G = cfg["common"]["fp_inter_dim"]
    dims = cfg["modal_dims"]
    order = [m for m in cfg["modal_feature_order"] if m in dims]
    M = len(order)
    tgt = cfg["fuser"]["_target_"].rsplit(".", 1)[-1]
    fl = 0.0
    for m in order:
        if dims[m] != D:
            fl += 2.0 * T * dims[m] * D
    per_row_block = 2.0 * (3 * D * D + D * D + 4 * D * D + 4 * D * D)
    if tgt == "ModalTokenCMFuser":
        fl += cfg["fuser"]["depth"] * T * (M + 1) * per_row_block
    elif tgt == "CMFuser":
        fl += cfg["fuser"]["depth"] * T * M * per_row_block
    elif tgt == "TemporalCMFuser":
        n = M + (1 if cfg["fuser"].get("frame_level_token") else 0)
        fl += cfg["fuser"]["depth"] * T * n * per_row_block
    else:  # CA: self-attn block + cross-attn (q,k,v,proj) + mlp per memory modality
        fl += (M - 1) * T * (per_row_block + 2.0 * 4 * D * D)
    fl += 2.0 * T * (D * G + G * D)                                   # dim_encoder / dim_decoder
    fl += cfg["common"]["fp_layers"] * T * 2.0 * (3 * G * G + G * G + 4 * G * G + 4 * G * G)
    for c in num_classes.values():
        fl += 2.0 * (T + 1) * D * c                                     # past (T rows) + future (1 row) heads
    return fl
