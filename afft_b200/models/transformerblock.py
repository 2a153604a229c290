"""Parameter containers mirroring reference models/transformerblock.py (same attribute names, shapes and
defaults).  They hold weights only: the arithmetic (LayerNorm -> QKV GEMM -> tiny attention -> proj GEMM ->
MLP GEMMs with fused bias/GELU/residual epilogues) is executed by libafft_b200 inside
``CMFPEarly.forward``; calling a block on its own is not a supported entry point.
"""
import torch.nn as nn


class _FusedOnly(nn.Module):
    def forward(self, *args, **kwargs):
        raise NotImplementedError(
            f"{type(self).__name__} is executed inside the fused afft_forward() call of CMFPEarly; "
            "it has no standalone (PyTorch) forward")


class Attention(_FusedOnly):
    """reference models/transformerblock.py:7-36"""

    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop=0., proj_drop=0.):
        super().__init__()
        if qkv_bias:
            raise NotImplementedError("qkv_bias=True is not used by any reference config and is not supported")
        if qk_scale is not None:
            raise NotImplementedError("qk_scale override is not supported (head_dim ** -0.5 is used)")
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=False)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)


class CrossAttention(_FusedOnly):
    """reference models/transformerblock.py:39-76"""

    def __init__(self, dim, mem_dim=None, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop=0., proj_drop=0.):
        super().__init__()
        if qkv_bias or qk_scale is not None or (mem_dim not in (None, dim)):
            raise NotImplementedError("CrossAttention: only qkv_bias=False, default scale and mem_dim == dim are supported")
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.w_q = nn.Linear(dim, dim, bias=False)
        self.w_k = nn.Linear(dim, dim, bias=False)
        self.w_v = nn.Linear(dim, dim, bias=False)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)


class MLP(_FusedOnly):
    """reference models/transformerblock.py:79-93 (Linear, erf-GELU, Linear, Dropout)"""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.):
        super().__init__()
        if act_layer is not nn.GELU:
            raise NotImplementedError("only nn.GELU (erf) is fused into the FC1 epilogue")
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.mlp = nn.Sequential(nn.Linear(in_features, hidden_features), act_layer(),
                                 nn.Linear(hidden_features, out_features), nn.Dropout(drop))


class DropPath(nn.Module):
    """reference models/transformerblock.py:108-115; identity in eval mode (the only mode the fused path runs)."""

    def __init__(self, drop_prob=None):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        if self.training and self.drop_prob:
            raise NotImplementedError("stochastic depth (training mode) is not implemented in the fused path")
        return x


class Block(_FusedOnly):
    """reference models/transformerblock.py:118-135"""

    def __init__(self, dim, num_heads, mlp_ratio=4., qkv_bias=False, qk_scale=None, drop=0., attn_drop=0.,
                 drop_path=0., act_layer=nn.GELU, norm_layer=nn.LayerNorm):
        super().__init__()
        if mlp_ratio != 4.:
            raise NotImplementedError("mlp_ratio must be 4 (the only value the reference uses)")
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale, attn_drop=attn_drop,
                              proj_drop=drop)
        self.drop_path = DropPath(drop_path) if drop_path > 0. else nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = MLP(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)


class DecoderBlock(_FusedOnly):
    """reference models/transformerblock.py:138-162"""

    def __init__(self, dim, mem_dim=None, num_heads=4, mlp_ratio=4., qkv_bias=False, qk_scale=None, drop=0.,
                 attn_drop=0., drop_path=0., act_layer=nn.GELU, norm_layer=nn.LayerNorm):
        super().__init__()
        if mlp_ratio != 4.:
            raise NotImplementedError("mlp_ratio must be 4 (the only value the reference uses)")
        self.norm_self = norm_layer(dim)
        self.attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale, attn_drop=attn_drop,
                              proj_drop=drop)
        self.cross_attn = CrossAttention(dim, mem_dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale,
                                         attn_drop=attn_drop, proj_drop=drop)
        self.drop_path = DropPath(drop_path) if drop_path > 0. else nn.Identity()
        self.norm_q = norm_layer(dim)
        self.norm_kv = norm_layer(mem_dim or dim)
        self.norm_mlp = norm_layer(dim)
        self.mlp = MLP(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)
