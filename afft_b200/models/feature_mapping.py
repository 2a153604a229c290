"""Mirror of reference models/feature_mapping.py: the per-modality projection into the common width.

``Linear`` in its shipped form (bias-free, no LayerNorm, Identity when the widths match) is executed inside
``afft_forward`` - the projection GEMM writes straight into the modality's token slot of the fuser's residual
stream - and ``fused_in_forward`` says so.  The ablation variants (``use_layernorm``, ``sparse_mapping=false``,
``NonLinear``, ``GatedLinear``: conf/model/mapping/*.yaml) run here as a short chain
of library kernels (afft_b200/hostops.py) and hand a common-width fp32 tensor to the fuser.
"""
from functools import partial

import torch
from torch import nn as nn

from .. import _capi, hostops

norm_layer_1e6 = partial(nn.LayerNorm, eps=1e-6)


class _Mapping(nn.Module):
    fused_in_forward = False  # True: afft_forward applies self.mapping[0] itself

    def _prepare(self, x):
        if self.training:
            raise NotImplementedError(f"{type(self).__name__}: the training-step path supports the plain Linear mapping only")
        lead = x.shape[:-1]
        return x.reshape(-1, x.shape[-1]).to(torch.float32).contiguous(), lead

    def _finish(self, y, lead):
        if self.use_layernorm:
            y = hostops.layernorm(y.contiguous(), self.mapping[-1])
        return y.reshape(*lead, y.shape[-1])

    def _cache(self) -> hostops.WeightCache:
        c = self.__dict__.get("_wcache")
        if c is None:
            c = self.__dict__["_wcache"] = hostops.WeightCache()
        return c

    precision = "bf16"  # set by the owning head ('bf16' | 'fp16' | 'strict')


class Linear(_Mapping):
    """reference models/feature_mapping.py:54-78"""

    def __init__(self, in_features, out_features, use_layernorm: bool = False, sparse_mapping=True):
        super().__init__()
        if sparse_mapping:
            layers = [nn.Linear(in_features, out_features, bias=False)
                      if in_features != out_features else nn.Identity()]
        else:
            layers = [nn.Linear(in_features, out_features, bias=False)]
        if use_layernorm:
            layers.append(norm_layer_1e6(out_features))
        self.mapping = nn.Sequential(*layers)
        self.use_layernorm = use_layernorm
        self.sparse_mapping = sparse_mapping
        self.in_features, self.out_features = in_features, out_features
        self.fused_in_forward = not use_layernorm and (sparse_mapping or in_features != out_features)

    def forward(self, x):
        x2, lead = self._prepare(x)
        lin = self.mapping[0]
        y = x2 if isinstance(lin, nn.Identity) else hostops.dense(x2, lin.weight, None, self._cache(), precision=self.precision)
        return self._finish(y, lead)

    def __str__(self):
        return f'Linear mapping layer with use_layernorm: {self.use_layernorm}, ' \
               f'and sparse_mapping: {self.sparse_mapping}'


class ContextGating(nn.Module):
    """reference models/feature_mapping.py:21-33: x * sigmoid(fc(x)) (cat + glu); a parameter container here - the
    gate is the epilogue of the fc GEMM (AFFT_ACT_GATE)."""

    def __init__(self, dimension):
        super().__init__()
        self.fc = nn.Linear(dimension, dimension)


class GatedLinear(_Mapping):
    """reference models/feature_mapping.py:36-51"""

    def __init__(self, in_features, out_features, use_layernorm: bool = True):
        super().__init__()
        layers = [nn.Linear(in_features, out_features), ContextGating(out_features)]
        if use_layernorm:
            layers.append(norm_layer_1e6(out_features))
        self.mapping = nn.Sequential(*layers)
        self.use_layernorm = use_layernorm

    def forward(self, x):
        x2, lead = self._prepare(x)
        lin, cg = self.mapping[0], self.mapping[1]
        u = hostops.dense(x2, lin.weight, lin.bias, self._cache(), precision=self.precision)
        y = hostops.dense(u, cg.fc.weight, cg.fc.bias, self._cache(), precision=self.precision, act=_capi.ACT_GATE, res=u)
        return self._finish(y, lead)

    def __str__(self):
        return f'Gated linear mapping layer with use_layernorm: {self.use_layernorm}'


_ACTS = {'relu': (nn.ReLU, _capi.ACT_RELU), 'gelu': (nn.GELU, _capi.ACT_GELU_ERF), 'none': (nn.Identity, _capi.ACT_NONE)}


class NonLinear(_Mapping):
    """reference models/feature_mapping.py:91-107"""

    def __init__(self, in_features, out_features, use_layernorm: bool = False, activation='relu'):
        super().__init__()
        assert activation in _ACTS, f'{activation} is not supported in {list(_ACTS)}.'
        layers = [nn.Linear(in_features, out_features), _ACTS[activation][0]()]
        if use_layernorm:
            layers.append(norm_layer_1e6(out_features))
        self.mapping = nn.Sequential(*layers)
        self.use_layernorm = use_layernorm
        self.activation = activation

    def forward(self, x):
        x2, lead = self._prepare(x)
        lin = self.mapping[0]
        y = hostops.dense(x2, lin.weight, lin.bias, self._cache(), precision=self.precision, act=_ACTS[self.activation][1])
        return self._finish(y, lead)

    def __str__(self):
        return f'Nonlinear mapping layer with use_layernorm: {self.use_layernorm}, activation: {self.activation}'
