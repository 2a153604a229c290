"""Mirror of reference models/fusion.py: the fusers as parameter containers with the reference's
constructor signatures, parameter names, shapes and initialisation.

Name map (reference models/fusion.py:3-9):  ModalTokenCMFuser = SA-Fuser, CMFuser = SA-Fuser without token,
TemporalCMFuser = T-SA-Fuser, TemporalCrossAttentFuser = CA-Fuser.  Their arithmetic runs inside
``afft_forward`` (afft_b200/csrc/afft_api.cu: run_fuser); ``afft_kind`` selects the native code path.
"""
from functools import partial

import torch
import torch.nn as nn

from .. import _capi, hostops
from .transformerblock import Block, DecoderBlock


def _init_weights(m):
    # reference models/fusion.py:21-27 (timm ViT init): trunc-normal(0.02) Linear weights, zero biases
    if isinstance(m, nn.Linear):
        nn.init.trunc_normal_(m.weight, std=.02)
        if m.bias is not None:
            nn.init.constant_(m.bias, 0)


class _Fuser(nn.Module):
    afft_kind = None
    precision = "bf16"  # 'bf16' | 'fp16' | 'strict' (set by the owning head; see CMFPEarly)
    max_batch = 64

    def forward(self, modal_feats, ordered_feature_list):
        """The reference's inner seam ``fuser(modal_feats, ordered_feature_list) -> (fused (B, T, C), attn)``
        (models/fusion.py:86,159,243,319), evaluated by a native fuser-only handle (AFFT_STAGE_FUSER): the same kernels
        CMFPEarly's fused call runs, stopping after the fuser's final LayerNorm.  ``modal_feats`` are the mapped
        features {modality: (B, T, dim)}.  Inference only (the training step goes through afft_b200.train)."""
        from ..engine import Engine
        if self.training:
            raise NotImplementedError("the standalone fuser seam is inference-only; training runs through CMFPEarly.forward")
        shape = next(iter(modal_feats.values())).shape
        assert all(v.shape == shape for v in modal_feats.values()), \
            'The shape of all inputs of the fusion module should be the same!'
        order = list(modal_feats.keys())
        feats = ordered_feature_list(modal_feats)
        assert len(feats) == len(order)
        # the caller's ordering function decides the token order; recover the modality names in that order
        names = []
        for f in feats:
            names.append(next(m for m in order if modal_feats[m] is f and m not in names))
        B, T, D = shape
        first = feats[0]
        if first.device.type != "cuda":
            raise _capi.AfftError("afft_b200 runs on CUDA devices only (sm_100a); there is no CPU path")
        if getattr(self, "frame_level_token", False) and getattr(self, "temporal_sequence_length", None) is not None:
            assert self.temporal_sequence_length == T, f"Temporal sequence length not valid {self.temporal_sequence_length} vs {T}"
        engines = self.__dict__.setdefault("_seam_engines", {})
        key = (tuple(names), T, str(first.device), self.precision)
        eng = engines.get(key)
        if eng is not None and B > eng.max_batch:
            eng.close()
            eng = None
        if eng is None:
            eng = Engine(fuser_kind=self.afft_kind, T=T, mod_names=names, mod_dims=[D] * len(names), dim=D,
                         fuser_depth=self.depth, fuser_heads=self.num_heads, modal_encoding=bool(self.modal_encoding),
                         frame_level_token=bool(self.frame_level_token), cross_attn=bool(self.cross_attn),
                         norm_elementwise=bool(self.norm_elementwise), gpt_dim=2 * D, gpt_layers=1, gpt_heads=max(1, 2 * D // 512),
                         cls_names=[], cls_dims=[], precision=self.precision, max_batch=max(B, self.max_batch),
                         device=first.device, stages=_capi.STAGE_FUSER)
            engines[key] = eng
        eng.sync_weights({"fuser." + n: p for n, p in self.named_parameters()})
        fused, attn = eng.forward_fuser([f.to(torch.float32).contiguous() for f in feats])
        if attn is None:
            attn = torch.zeros(B)  # CA-Fuser's dummy attention (reference fusion.py:269)
        return fused, attn

    def _check_rates(self, act_layer, mlp_ratio, qkv_bias, qk_scale):
        if act_layer is not nn.GELU or mlp_ratio != 4. or qkv_bias or qk_scale is not None:
            raise NotImplementedError("fused fuser supports act_layer=nn.GELU, mlp_ratio=4, qkv_bias=False, qk_scale=None")


class ModalTokenCMFuser(_Fuser):
    """SA-Fuser with modality token - reference models/fusion.py:273-365"""
    afft_kind = _capi.FUSER_SA

    def __init__(self, dim, depth=1, num_heads=4, mlp_ratio=4., qkv_bias=False, qk_scale=None, embd_drop_rate=0.,
                 drop_rate=0., attn_drop_rate=0., drop_path_rate=0., act_layer=nn.GELU,
                 norm_elementwise=True, cross_attn=False, modalities=None, modal_encoding=False,
                 frame_level_token=False, temporal_sequence_length=None):
        super().__init__()
        self._check_rates(act_layer, mlp_ratio, qkv_bias, qk_scale)
        norm_layer = partial(nn.LayerNorm, eps=1e-6, elementwise_affine=norm_elementwise)
        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, depth)]
        self.blocks = nn.ModuleList([
            Block(dim=dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale,
                  drop=drop_rate, attn_drop=attn_drop_rate, drop_path=dpr[i], act_layer=act_layer,
                  norm_layer=norm_layer) for i in range(depth)])
        self.norm = norm_layer(dim)
        self.num_mods = len(modalities) + 1
        self.modality_embedding = nn.Parameter(torch.zeros(1, self.num_mods, dim)) if modal_encoding else None
        self.embd_drop = nn.Dropout(embd_drop_rate)
        self.cross_attn = cross_attn
        self.frame_level_token = frame_level_token
        self.temporal_sequence_length = temporal_sequence_length
        if not frame_level_token:
            self.modal_token = nn.Parameter(torch.zeros(1, 1, dim))
        else:
            assert temporal_sequence_length is not None, "Temporal sequence length must be provided!"
            self.modal_token = nn.Parameter(torch.zeros(1, temporal_sequence_length, dim))
        nn.init.trunc_normal_(self.modal_token, std=.02)
        if self.modality_embedding is not None:
            nn.init.trunc_normal_(self.modality_embedding, std=.02)
        self.apply(_init_weights)
        self.dim, self.depth, self.num_heads = dim, depth, num_heads
        self.norm_elementwise, self.modal_encoding = norm_elementwise, modal_encoding


class CMFuser(_Fuser):
    """SA-Fuser without modality token - reference models/fusion.py:61-118"""
    afft_kind = _capi.FUSER_SA_NOTOKEN

    def __init__(self, dim, depth=1, num_heads=4, mlp_ratio=4., qkv_bias=False, qk_scale=None, embd_drop_rate=0.,
                 drop_rate=0., attn_drop_rate=0., drop_path_rate=0., act_layer=nn.GELU,
                 norm_layer=partial(nn.LayerNorm, eps=1e-6), cross_attn=False):
        super().__init__()
        self._check_rates(act_layer, mlp_ratio, qkv_bias, qk_scale)
        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, depth)]
        self.blocks = nn.ModuleList([
            Block(dim=dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale,
                  drop=drop_rate, attn_drop=attn_drop_rate, drop_path=dpr[i], act_layer=act_layer,
                  norm_layer=norm_layer) for i in range(depth)])
        self.norm = norm_layer(dim)
        self.embd_drop = nn.Dropout(embd_drop_rate)
        self.cross_attn = cross_attn
        self.apply(_init_weights)
        self.dim, self.depth, self.num_heads = dim, depth, num_heads
        self.norm_elementwise, self.modal_encoding, self.frame_level_token = True, False, False


class TemporalCMFuser(_Fuser):
    """T-SA-Fuser - reference models/fusion.py:121-215"""
    afft_kind = _capi.FUSER_TSA

    def __init__(self, dim, depth=1, num_heads=4, mlp_ratio=4., qkv_bias=False, qk_scale=None, embd_drop_rate=0.,
                 drop_rate=0., attn_drop_rate=0., drop_path_rate=0., act_layer=nn.GELU,
                 norm_layer=partial(nn.LayerNorm, eps=1e-6), modalities=None, modal_encoding=True,
                 frame_level_token=False, temporal_sequence_length=None, max_position_embeddings=64):
        super().__init__()
        self._check_rates(act_layer, mlp_ratio, qkv_bias, qk_scale)
        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, depth)]
        self.blocks = nn.ModuleList([
            Block(dim=dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale,
                  drop=drop_rate, attn_drop=attn_drop_rate, drop_path=dpr[i], act_layer=act_layer,
                  norm_layer=norm_layer) for i in range(depth)])
        self.norm = norm_layer(dim)
        self.num_mods = len(modalities) + 1 if frame_level_token else len(modalities)
        self.modality_embedding = nn.Parameter(torch.zeros(self.num_mods, dim)) if modal_encoding else None
        self.position_embeddings = nn.Embedding(max_position_embeddings, dim)
        self.embd_drop = nn.Dropout(embd_drop_rate)
        self.frame_level_token = frame_level_token
        self.temporal_sequence_length = temporal_sequence_length
        self.modal_token = None
        if frame_level_token:
            assert temporal_sequence_length is not None, "Temporal sequence length must be provided!"
            self.modal_token = nn.Parameter(torch.zeros(1, temporal_sequence_length, dim))
        if self.modal_token is not None:
            nn.init.trunc_normal_(self.modal_token, std=.02)
        if self.modality_embedding is not None:
            nn.init.trunc_normal_(self.modality_embedding, std=.02)
        self.apply(_init_weights)
        self.dim, self.depth, self.num_heads = dim, depth, num_heads
        self.norm_elementwise, self.modal_encoding, self.cross_attn = True, modal_encoding, False


class TemporalCrossAttentFuser(_Fuser):
    """CA-Fuser - reference models/fusion.py:218-270 (depth = number of modalities - 1)"""
    afft_kind = _capi.FUSER_CA

    def __init__(self, dim, modalities=None, num_heads=4, mlp_ratio=4., qkv_bias=False, qk_scale=None,
                 embd_drop_rate=0., drop_rate=0., attn_drop_rate=0., drop_path_rate=0., act_layer=nn.GELU,
                 norm_layer=partial(nn.LayerNorm, eps=1e-6), max_position_embeddings=128):
        super().__init__()
        self._check_rates(act_layer, mlp_ratio, qkv_bias, qk_scale)
        depth = len(modalities) - 1
        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, depth)]
        self.blocks = nn.ModuleList([
            DecoderBlock(dim=dim, mem_dim=None, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias,
                         qk_scale=qk_scale, drop=drop_rate, attn_drop=attn_drop_rate, drop_path=dpr[i],
                         act_layer=act_layer, norm_layer=norm_layer) for i in range(depth)])
        self.norm = norm_layer(dim)
        self.embd_drop = nn.Dropout(embd_drop_rate)
        self.position_embeddings = nn.Embedding(max_position_embeddings, dim)
        self.apply(_init_weights)
        self.dim, self.depth, self.num_heads = dim, depth, num_heads
        self.norm_elementwise, self.modal_encoding, self.frame_level_token, self.cross_attn = True, False, False, False


class MATT(nn.Module):
    """Modality attention of RULSTM - reference models/fusion.py:35-58: a 3-layer MLP over the concatenated mapped
    features whose softmax weights the per-modality classification scores (CMFPScoreFusion).  The three linears
    (ReLU in the GEMM epilogue) and the softmax are library kernels (afft_gemm, afft_score_fusion)."""

    def __init__(self, modal_dims, dim=None, drop_rate=0.8):
        super().__init__()
        num_modality = len(modal_dims)
        in_size = dim * num_modality if dim else sum(modal_dims.values())
        self.matt = nn.Sequential(nn.Linear(in_size, int(in_size / 4)), nn.ReLU(), nn.Dropout(drop_rate),
                                  nn.Linear(int(in_size / 4), int(in_size / 8)), nn.ReLU(), nn.Dropout(drop_rate),
                                  nn.Linear(int(in_size / 8), num_modality))
        self.num_modality = num_modality
        self.precision = "bf16"
        self.__dict__["_wcache"] = hostops.WeightCache()

    def attn_logits(self, modal_feats, ordered_feature_list):
        """The pre-softmax modality scores, (B*S, num_modality) fp32 (row pitch padded to 4)."""
        if self.training:
            raise NotImplementedError("MATT: the training-step path is not built for score fusion")
        x = torch.cat(ordered_feature_list(modal_feats), dim=2)
        x = x.reshape(-1, x.shape[-1]).to(torch.float32)
        cache = self.__dict__["_wcache"]
        for idx, act in ((0, _capi.ACT_RELU), (3, _capi.ACT_RELU), (6, _capi.ACT_NONE)):
            lin = self.matt[idx]
            x = hostops.dense(x, lin.weight, lin.bias, cache, precision=self.precision, act=act)
        return x

    def forward(self, modal_feats, ordered_feature_list):
        first = next(iter(modal_feats.values()))
        scores = self.attn_logits(modal_feats, ordered_feature_list)
        attn = torch.empty(scores.shape[0], self.num_modality, device=scores.device, dtype=torch.float32)
        _capi.score_fusion(scores, attn=attn)
        return attn.reshape(first.shape[0], first.shape[1], self.num_modality)
