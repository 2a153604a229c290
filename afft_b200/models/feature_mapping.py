"""Mirror of reference models/feature_mapping.py: the per-modality projection into the common width."""
from functools import partial

from torch import nn as nn


class Linear(nn.Module):
    """reference models/feature_mapping.py:54-78: bias-free nn.Linear, or Identity when in == out with
    sparse_mapping.  The projection GEMM itself runs inside afft_forward (it writes straight into the
    modality's token slot of the fuser's residual stream)."""

    def __init__(self, in_features, out_features, use_layernorm: bool = False, sparse_mapping=True):
        super().__init__()
        if use_layernorm:
            raise NotImplementedError("mapping.use_layernorm=true is not used by any fusion config and is not supported")
        if not sparse_mapping and in_features == out_features:
            raise NotImplementedError("sparse_mapping=false with equal widths is not supported")
        layers = [nn.Linear(in_features, out_features, bias=False) if in_features != out_features else nn.Identity()]
        self.mapping = nn.Sequential(*layers)
        self.use_layernorm = use_layernorm
        self.sparse_mapping = sparse_mapping
        self.in_features, self.out_features = in_features, out_features

    def forward(self, x):
        raise NotImplementedError("feature mapping is executed inside the fused afft_forward() call")

    def __str__(self):
        return f'Linear mapping layer with use_layernorm: {self.use_layernorm}, ' \
               f'and sparse_mapping: {self.sparse_mapping}'


class GatedLinear(nn.Module):
    def __init__(self, *a, **k):
        raise NotImplementedError("GatedLinear mapping (ablation configs, SURVEY.md section 8f row N3) is not built yet")


class NonLinear(nn.Module):
    def __init__(self, *a, **k):
        raise NotImplementedError("NonLinear mapping (ablation configs, SURVEY.md section 8f row N3) is not built yet")


norm_layer_1e6 = partial(nn.LayerNorm, eps=1e-6)
