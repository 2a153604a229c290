"""CPU oracle for the AFFT fusion-and-anticipation forward path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import it; nothing under afft_b200/ does.  The product path
has no CPU fallback.

It is an op-by-op restatement, in plain torch tensor arithmetic (matmul / mean / exp / erf / tanh - no
nn.Module, no F.layer_norm, no transformers), of what the reference computes under
BaseModel.future_predictor in eval mode:

    reference                                                    here
    models/future_prediction.py:257-291  CMFPEarly.forward        forward()
    models/future_prediction.py:133-142  feature_mapping          _mapping()
    models/feature_mapping.py:54-78      Linear (Identity if =)   _mapping()
    models/fusion.py:319-365             ModalTokenCMFuser        _sa_fuser(token=True)
    models/fusion.py:86-118              CMFuser                  _sa_fuser(token=False)
    models/fusion.py:159-215             TemporalCMFuser          _tsa_fuser()
    models/fusion.py:243-270             TemporalCrossAttentFuser _ca_fuser()
    models/transformerblock.py:19-36     Attention                _attention()
    models/transformerblock.py:60-76     CrossAttention           _cross_attention()
    models/transformerblock.py:91-93     MLP (erf GELU)           _mlp()
    models/transformerblock.py:131-135   Block                    _block()
    models/transformerblock.py:157-162   DecoderBlock             _decoder_block()
    models/future_prediction.py:387-415  BaseFuturePredictor      _gpt2() (output_len == 1)
    models/future_prediction.py:155-182  prepare_output           forward()
    models/future_prediction.py:144-153  apply_classifier         forward()
    models/feature_mapping.py:21-48,91-107 GatedLinear / NonLinear _mapping_layer()
    models/fusion.py:35-58               MATT                     _matt()
    models/future_prediction.py:189-225  IndividualFuturePrediction _forward_individual()
    models/future_prediction.py:294-351  CMFPScoreFusion          _forward_scorefusion()

GPT-2 itself is third-party: transformers (pinned 4.18.0 in the reference's environment.yml:166; 5.5.0 in
this image), class GPT2Model called with inputs_embeds / position_ids=arange(T) / past_key_values=None
(models/future_prediction.py:372-383,400-403).  _gpt2() restates its published algorithm: h = x + wpe[:T];
per layer h += c_proj(causal_softmax(q k^T / sqrt(head_dim)) v), h += c_proj(gelu_new(c_fc(ln_2 h))), with
pre-LayerNorm (eps 1e-5), Conv1D weights stored [in, out]; final ln_f.

PARITY PIN: the reference ships no tests, golden vectors or checkpoints, so the pin
is the reference module itself, imported in the build container by tests/golden/make_golden.py with
deterministic weights (afft_b200/synthetic.py).  That script checks this file against the module in
float64 (max |diff| ~1e-15) and stores the module's float32 outputs under tests/golden/*.npz;
tests/test_oracle.py re-checks this file against those fixtures on every run.
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import torch

PREFIX = "future_predictor."
MODAL_FEATURE_ORDER = ["rgb", "objects", "audio", "poses", "flow"]  # conf/config.yaml:41


# ------------------------------------------------------------------------------------------------
# primitives
# ------------------------------------------------------------------------------------------------
# ATEN_OPS = False (default): every primitive is spelled out in elementary tensor arithmetic (the checker).
# ATEN_OPS = True: the same primitives are issued as the single ATen library calls the reference module makes
# (F.layer_norm, F.linear/addmm, F.gelu, Tensor.softmax), so that timing this port on the host CPU costs what
# the reference module costs there (bench.py cpu_baseline / --impl reference).  tests/test_oracle.py checks
# that both spellings agree.
ATEN_OPS = False


def _ln(x, w, b, eps):
    if ATEN_OPS:
        return torch.nn.functional.layer_norm(x, (x.shape[-1],), w, b, eps)
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)  # biased, as nn.LayerNorm
    y = (x - mu) / torch.sqrt(var + eps)
    if w is not None:
        y = y * w + b
    return y


def _linear(x, w, b=None):  # nn.Linear: weight [out, in]
    if ATEN_OPS:
        return torch.nn.functional.linear(x, w, b)
    y = x @ w.t()
    return y if b is None else y + b


def _conv1d(x, w, b):  # transformers Conv1D: weight [in, out]
    if ATEN_OPS:
        return torch.addmm(b, x.reshape(-1, x.shape[-1]), w).reshape(*x.shape[:-1], w.shape[1])
    return x @ w + b


def _gelu_erf(x):
    if ATEN_OPS:
        return torch.nn.functional.gelu(x)
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def _gelu_new(x):  # transformers activations.py NewGELUActivation
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * x ** 3)))


def _softmax(s):
    if ATEN_OPS:
        return s.softmax(dim=-1)
    s = s - s.max(-1, keepdim=True).values
    e = torch.exp(s)
    return e / e.sum(-1, keepdim=True)


def _heads(x, H):  # (B, N, C) -> (B, H, N, C/H)
    B, N, C = x.shape
    return x.reshape(B, N, H, C // H).permute(0, 2, 1, 3)


def _merge(x):  # (B, H, N, c) -> (B, N, H*c)
    B, H, N, c = x.shape
    return x.permute(0, 2, 1, 3).reshape(B, N, H * c)


class _W:
    """State-dict view with a name prefix and dtype cast."""

    def __init__(self, sd, prefix, dtype):
        self.sd, self.prefix, self.dtype = sd, prefix, dtype

    def __call__(self, name, optional=False):
        k = self.prefix + name
        if k not in self.sd:
            if optional:
                return None
            raise KeyError(k)
        t = self.sd[k]
        return t.to(self.dtype) if t.requires_grad else t.detach().to(self.dtype)  # keeps autograd for gradient checks

    def sub(self, p):
        return _W(self.sd, self.prefix + p, self.dtype)


# ------------------------------------------------------------------------------------------------
# transformer pieces (models/transformerblock.py)
# ------------------------------------------------------------------------------------------------
def _attention(x, w: _W, H, mask):
    # transformerblock.py:19-36 - fused qkv without bias, scale head_dim^-0.5, additive mask, returns probs
    B, N, C = x.shape
    qkv = _linear(x, w("qkv.weight"), w("qkv.bias", optional=True)).reshape(B, N, 3, H, C // H).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    s = (q @ k.transpose(-2, -1)) * (C // H) ** -0.5
    if mask is not None:
        s = s + mask
    p = _softmax(s)
    o = _merge(p @ v)
    return _linear(o, w("proj.weight"), w("proj.bias")), p


def _cross_attention(x, mem, w: _W, H, mask):
    # transformerblock.py:60-76
    C = x.shape[-1]
    q = _heads(_linear(x, w("w_q.weight"), w("w_q.bias", optional=True)), H)
    k = _heads(_linear(mem, w("w_k.weight"), w("w_k.bias", optional=True)), H)
    v = _heads(_linear(mem, w("w_v.weight"), w("w_v.bias", optional=True)), H)
    s = (q @ k.transpose(-2, -1)) * (C // H) ** -0.5
    if mask is not None:
        s = s + mask
    o = _merge(_softmax(s) @ v)
    return _linear(o, w("proj.weight"), w("proj.bias"))


def _mlp(x, w: _W):
    # transformerblock.py:82-93: Linear, nn.GELU (erf), Linear
    return _linear(_gelu_erf(_linear(x, w("mlp.0.weight"), w("mlp.0.bias"))), w("mlp.2.weight"), w("mlp.2.bias"))


def _block(x, w: _W, H, mask, eps=1e-6):
    # transformerblock.py:131-135 (DropPath/Dropout are identities in eval mode)
    a, p = _attention(_ln(x, w("norm1.weight", True), w("norm1.bias", True), eps), w.sub("attn."), H, mask)
    x = x + a
    x = x + _mlp(_ln(x, w("norm2.weight", True), w("norm2.bias", True), eps), w.sub("mlp."))
    return x, p


def _decoder_block(x, mem, w: _W, H, mask, eps=1e-6):
    # transformerblock.py:157-162
    a, _ = _attention(_ln(x, w("norm_self.weight"), w("norm_self.bias"), eps), w.sub("attn."), H, mask)
    x = x + a
    x = x + _cross_attention(_ln(x, w("norm_q.weight"), w("norm_q.bias"), eps),
                             _ln(mem, w("norm_kv.weight"), w("norm_kv.bias"), eps), w.sub("cross_attn."), H, mask)
    x = x + _mlp(_ln(x, w("norm_mlp.weight"), w("norm_mlp.bias"), eps), w.sub("mlp."))
    return x


def _causal_mask(T, dtype):
    return torch.triu(torch.full((T, T), float("-inf"), dtype=dtype), diagonal=1)  # fusion.py:30-32


# ------------------------------------------------------------------------------------------------
# fusers (models/fusion.py)
# ------------------------------------------------------------------------------------------------
def _sa_fuser(feats: List[torch.Tensor], w: _W, fcfg, token: bool):
    B, T, C = feats[0].shape
    H, depth = fcfg["num_heads"], fcfg["depth"]
    x = torch.cat([f.reshape(B * T, 1, C) for f in feats], dim=1)  # fusion.py:338 / :105
    if token:
        tok = w("modal_token")
        if not fcfg.get("frame_level_token", False):
            toks = tok.expand(B * T, -1, -1)  # fusion.py:342
        else:
            toks = tok.expand(B, -1, -1).reshape(B * T, 1, -1)  # fusion.py:346
        x = torch.cat((toks, x), dim=1)  # fusion.py:349
        me = w("modality_embedding", optional=True)
        if me is not None:
            x = x + me  # fusion.py:352-353
    n = x.shape[1]
    mask = None
    if fcfg.get("cross_attn", False):  # fusion.py:313-317,331-332
        mask = torch.eye(n, dtype=x.dtype)
        mask = mask.masked_fill(mask == 1, float("-inf"))
    probs = []
    for i in range(depth):
        x, p = _block(x, w.sub(f"blocks.{i}."), H, mask)
        probs.append(p.reshape(B, T, *p.shape[1:]))  # fusion.py:360
    x = _ln(x, w("norm.weight", True), w("norm.bias", True), 1e-6)
    z = x[:, 0, :] if token else x.mean(dim=1)  # fusion.py:363 / :115
    return z.reshape(-1, T, C), torch.stack(probs).transpose(0, 1)


def _tsa_fuser(feats: List[torch.Tensor], w: _W, fcfg):
    B, T, C = feats[0].shape
    H, depth = fcfg["num_heads"], fcfg["depth"]
    flt = fcfg.get("frame_level_token", False)
    n = len(feats) + (1 if flt else 0)
    mask = _causal_mask(T, feats[0].dtype).repeat(n, n)  # fusion.py:170-171
    x = torch.cat(feats, dim=1)  # (B, n*T, C), token index m*T + t   fusion.py:177
    if flt:
        x = torch.cat((w("modal_token").expand(B, -1, -1), x), dim=1)  # fusion.py:183-184
    pos = w("position_embeddings.weight")[:T]
    x = x + pos.repeat(n, 1)  # fusion.py:187-190
    me = w("modality_embedding", optional=True)
    if me is not None:
        x = x + torch.cat([e.repeat(T, 1) for e in me], dim=0)  # fusion.py:193-196
    probs = []
    for i in range(depth):
        x, p = _block(x, w.sub(f"blocks.{i}."), H, mask)
        probs.append(p)
    x = _ln(x, w("norm.weight"), w("norm.bias"), 1e-6)
    if flt:
        z = x[:, :T, :]  # fusion.py:207-209
    else:
        z = torch.cat([x[:, list(range(i, x.size(1), T)), :].mean(dim=1, keepdim=True) for i in range(T)], dim=1)
    return z, torch.stack(probs).transpose(0, 1)


def _ca_fuser(feats: List[torch.Tensor], w: _W, fcfg):
    B, T, C = feats[0].shape
    H = fcfg["num_heads"]
    mask = _causal_mask(T, feats[0].dtype)
    pos = w("position_embeddings.weight")[:T]
    feats = [f + pos for f in feats]  # fusion.py:262
    x, mems = feats[0], feats[1:]
    for i in range(len(mems)):  # depth = len(modalities) - 1, fusion.py:229,266-267
        x = _decoder_block(x, mems[i], w.sub(f"blocks.{i}."), H, mask)
    x = _ln(x, w("norm.weight"), w("norm.bias"), 1e-6)
    return x, torch.zeros(B)  # fusion.py:269 dummy attention


# ------------------------------------------------------------------------------------------------
# GPT-2 (transformers GPT2Model, restated)
# ------------------------------------------------------------------------------------------------
def _gpt2_once(emb, w: _W, n_layer, n_head, probs_out=None):
    """GPT2Model on input embeddings that already include the position embeddings.  probs_out (a list) receives the
    attention probabilities of every layer, (B, H, T, T) each - what `output_attentions=True` returns from
    transformers' eager GPT-2 attention (modeling_gpt2.py: softmax of the masked, scaled scores, before attn_dropout,
    which is the identity in eval mode)."""
    B, T, G = emb.shape
    h = emb
    mask = _causal_mask(T, emb.dtype)
    for i in range(n_layer):
        wl = w.sub(f"h.{i}.")
        y = _ln(h, wl("ln_1.weight"), wl("ln_1.bias"), 1e-5)
        qkv = _conv1d(y, wl("attn.c_attn.weight"), wl("attn.c_attn.bias"))
        q, k, v = qkv.split(G, dim=2)
        q, k, v = _heads(q, n_head), _heads(k, n_head), _heads(v, n_head)
        s = (q @ k.transpose(-1, -2)) / math.sqrt(G // n_head) + mask
        p = _softmax(s)
        if probs_out is not None:
            probs_out.append(p)
        a = _merge(p @ v)
        h = h + _conv1d(a, wl("attn.c_proj.weight"), wl("attn.c_proj.bias"))
        y = _ln(h, wl("ln_2.weight"), wl("ln_2.bias"), 1e-5)
        h = h + _conv1d(_gelu_new(_conv1d(y, wl("mlp.c_fc.weight"), wl("mlp.c_fc.bias"))), wl("mlp.c_proj.weight"),
                        wl("mlp.c_proj.bias"))
    return _ln(h, w("ln_f.weight"), w("ln_f.bias"), 1e-5)


def _gpt2(x, w: _W, n_layer, n_head, output_len=1, probs_out=None):
    """BaseFuturePredictor.forward (future_prediction.py:387-415).  output_len > 1: the last hidden state is fed
    back as the next input embedding at position T + k - 1 (:398-412).  The reference runs that single position
    against its KV cache; because attention is causal that equals re-running the extended sequence, which is
    what this checker does."""
    B, T, G = x.shape
    wpe = w("wpe.weight")
    emb = x + wpe[:T]
    hidden = _gpt2_once(emb, w, n_layer, n_head, probs_out)
    outs = [hidden]
    for k in range(1, output_len):
        emb = torch.cat([emb, hidden[:, -1:] + wpe[T + k - 1]], dim=1)
        hidden = _gpt2_once(emb, w, n_layer, n_head)
        outs.append(hidden[:, -1:])
    return torch.cat(outs, dim=1)  # (B, T + output_len - 1, G)


# ------------------------------------------------------------------------------------------------
# feature mapping variants (models/feature_mapping.py) and MATT (models/fusion.py:35-58)
# ------------------------------------------------------------------------------------------------
def _mapping_layer(x, w: _W, mcfg):
    """One modality's mapping nn.Sequential; ``w`` is rooted at ``mapping.<mod>.mapping.``.
    Linear (feature_mapping.py:54-78): [Linear(no bias) | Identity if in == out and sparse_mapping] [+ LN 1e-6]
    NonLinear (:91-107): Linear(bias), relu | gelu | none [+ LN]
    GatedLinear (:36-51): Linear(bias), ContextGating x * sigmoid(fc(x)) (:21-33; cat + glu over dim 1) [+ LN]"""
    kind = mcfg["_target_"].rsplit(".", 1)[-1] if mcfg is not None else "Linear"
    use_ln = bool(mcfg.get("use_layernorm", kind != "Linear")) if mcfg is not None else False
    if kind == "Linear":
        mw = w("0.weight", optional=True)
        y = x if mw is None else _linear(x, mw)
        ln_idx = 1
    elif kind == "NonLinear":
        y = _linear(x, w("0.weight"), w("0.bias"))
        act = mcfg.get("activation", "relu")
        if act == "relu":
            y = torch.clamp(y, min=0)
        elif act == "gelu":
            y = _gelu_erf(y)
        elif act != "none":
            raise ValueError(act)
        ln_idx = 2
    elif kind == "GatedLinear":
        y = _linear(x, w("0.weight"), w("0.bias"))
        gate = _linear(y, w("1.fc.weight"), w("1.fc.bias"))
        y = y * (1.0 / (1.0 + torch.exp(-gate)))
        ln_idx = 2
    else:
        raise ValueError(kind)
    if use_ln:
        y = _ln(y, w(f"{ln_idx}.weight"), w(f"{ln_idx}.bias"), 1e-6)
    return y


def _matt(flist: List[torch.Tensor], w: _W):
    """MATT.forward (fusion.py:49-58): concat over channels, Linear-ReLU-Linear-ReLU-Linear, softmax over modalities
    (Dropout layers at indices 2 and 5 are identities in eval mode)."""
    x = torch.cat(flist, dim=2)
    x = torch.clamp(_linear(x, w("matt.0.weight"), w("matt.0.bias")), min=0)
    x = torch.clamp(_linear(x, w("matt.3.weight"), w("matt.3.bias")), min=0)
    return _softmax(_linear(x, w("matt.6.weight"), w("matt.6.bias")))


def _predict_unimodal(z, w: _W, cfg, m):
    """dim_encoder[m] -> (shared or per-modality) GPT-2 -> dim_decoder[m]  (future_prediction.py:203-214)."""
    enc = w(f"dim_encoder.{m}.weight", optional=True)
    dec = w(f"dim_decoder.{m}.weight", optional=True)
    gw = w.sub("future_predictor.gpt_model.") if cfg["common"]["share_predictors"] else \
        w.sub(f"future_predictor.{m}.gpt_model.")
    g = _gpt2(z if enc is None else _linear(z, enc), gw, cfg["common"]["fp_layers"], cfg["common"]["fp_heads"],
              cfg["common"].get("fp_output_len", 1))
    return g if dec is None else _linear(g, dec)


def _unimodal_outputs(w: _W, cfg, num_classes, feats):
    """The part IndividualFuturePrediction and CMFPScoreFusion share: per-modality prediction, prepare_output
    (:155-182) and the per-modality classifiers (:144-153)."""
    z_hat = {m: _predict_unimodal(z, w, cfg, m) for m, z in feats.items()}
    T = next(iter(feats.values())).shape[1]
    out = {"orig_past": dict(feats), "future": {m: zh[:, T - 1:] for m, zh in z_hat.items()}, "all-fused": {},
           "past_futures": {m: torch.cat([feats[m][:, :1], z_hat[m][:, :T - 1]], dim=1) for m in z_hat}}
    for prefix, src in (("past_", out["past_futures"]), ("", out["future"])):
        for c in num_classes:
            out[f"{prefix}logits/{c}"] = {m: _linear(x, w(f"classifiers.{c}.{m}.1.weight"), w(f"classifiers.{c}.{m}.1.bias"))
                                          for m, x in src.items()}
    return out, z_hat


def _forward_individual(w: _W, cfg, num_classes, feats):
    out, _ = _unimodal_outputs(w, cfg, num_classes, feats)  # future_prediction.py:200-225
    return out


def _forward_scorefusion(w: _W, cfg, num_classes, feats):
    """CMFPScoreFusion.forward (future_prediction.py:307-351): per-modality logits weighted by MATT's softmax over
    the mapped [first frame | predictions] sequence."""
    order = [m for m in cfg["modal_feature_order"] if m in feats]
    out, z_hat = _unimodal_outputs(w, cfg, num_classes, feats)
    mapped = [_mapping_layer(torch.cat([feats[m][:, :1], z_hat[m]], dim=1), w.sub(f"mapping.{m}.mapping."), cfg.get("mapping"))
              for m in order]  # :327-333
    attn = _matt(mapped, w.sub("fuser."))  # (B, T + 1, M)
    for c in num_classes:  # :341-350
        past, fut = out[f"past_logits/{c}"], out[f"logits/{c}"]
        out[f"past_logits/{c}"] = {"all-fused": sum(attn[:, :-1, i].unsqueeze(-1) * past[m] for i, m in enumerate(order))}
        out[f"logits/{c}"] = {"all-fused": sum(attn[:, -1:, i].unsqueeze(-1) * fut[m] for i, m in enumerate(order))}
    out["modality_attns"] = attn  # not a reference output key; exposed for the parity tests
    return out


# ------------------------------------------------------------------------------------------------
# the path
# ------------------------------------------------------------------------------------------------
def forward(state_dict: Dict[str, torch.Tensor], cfg: Dict, num_classes: Dict[str, int],
            feats: Dict[str, torch.Tensor], dtype=torch.float32) -> Dict:
    """CMFPEarly.forward (or IndividualFuturePrediction / CMFPScoreFusion, by cfg["CMFP"]["_target_"]) on
    {mod: (B, T, C_mod)} features; returns the reference's dict-of-dict outputs."""
    w = _W(state_dict, PREFIX, dtype)
    feats = {m: f.reshape(f.shape[0], f.shape[1], -1).to(dtype) for m, f in feats.items()}
    head = cfg.get("CMFP", {}).get("_target_", "CMFPEarly").rsplit(".", 1)[-1]
    if head == "IndividualFuturePrediction":
        return _forward_individual(w, cfg, num_classes, feats)
    if head == "CMFPScoreFusion":
        return _forward_scorefusion(w, cfg, num_classes, feats)
    order = [m for m in cfg["modal_feature_order"] if m in feats]  # future_prediction.py:258
    D = cfg["common"]["in_features"]

    # feature mapping (future_prediction.py:133-142): by default a bias-free Linear, Identity when the width
    # already matches (feature_mapping.py:59-63)
    mapped = {}
    for m, x in feats.items():
        mapped[m] = _mapping_layer(x, w.sub(f"mapping.{m}.mapping."), cfg.get("mapping"))
        assert mapped[m].shape[-1] == D
    flist = [mapped[m] for m in order]

    fcfg = cfg["fuser"]
    kind = fcfg["_target_"].rsplit(".", 1)[-1]
    wf = w.sub("fuser.")
    if kind == "ModalTokenCMFuser":
        z, mattn = _sa_fuser(flist, wf, fcfg, token=True)
    elif kind == "CMFuser":
        z, mattn = _sa_fuser(flist, wf, fcfg, token=False)
    elif kind == "TemporalCMFuser":
        z, mattn = _tsa_fuser(flist, wf, fcfg)
    elif kind == "TemporalCrossAttentFuser":
        z, mattn = _ca_fuser(flist, wf, fcfg)
    else:
        raise ValueError(kind)

    B, T, _ = z.shape
    enc = w("dim_encoder.weight", optional=True)
    dec = w("dim_decoder.weight", optional=True)
    z_enc = z if enc is None else _linear(z, enc)  # future_prediction.py:267
    # fp_output_attentions (future_prediction.py:403-409): {'gpt2_att_0': (B, n_layer, n_head, T, T)}.  PARITY UNPINNED for
    # this one output: transformers 5.5 (installed) runs GPT-2 attention through SDPA and returns no probabilities - the
    # reference module itself fails at :409 here - so this restates the eager attention of the pinned 4.18.
    want_t = bool(cfg.get("future_predictor", {}).get("output_attentions", False))
    t_probs = [] if want_t else None
    g = _gpt2(z_enc, w.sub("future_predictor.gpt_model."), cfg["common"]["fp_layers"], cfg["common"]["fp_heads"],
              cfg["common"].get("fp_output_len", 1), t_probs)
    z_hat = g if dec is None else _linear(g, dec)  # future_prediction.py:269

    out = {  # prepare_output, future_prediction.py:155-182
        "orig_past": {"all-fused": z},
        "future": {"all-fused": z_hat[:, T - 1:]},
        "all-fused": {"all-fused": z[:, T - 1:]},
        "past_futures": {"all-fused": torch.cat([z[:, :1], z_hat[:, :T - 1]], dim=1)},
    }
    for prefix, src in (("past_", out["past_futures"]["all-fused"]), ("", out["future"]["all-fused"])):
        for c in num_classes:  # apply_classifier, future_prediction.py:144-153 (Dropout is identity in eval)
            out[f"{prefix}logits/{c}"] = {"all-fused": _linear(src, w(f"classifiers.{c}.all-fused.1.weight"),
                                                               w(f"classifiers.{c}.all-fused.1.bias"))}
    temporal = {"gpt2_att_0": torch.stack(t_probs).transpose(0, 1)} if want_t else {}
    out["attentions"] = {"all-fused": {"modality_attns": mattn, "temporal_attns": temporal}}
    return out


def top5(logits: torch.Tensor) -> torch.Tensor:
    """Ordered top-5 class indices per row (the quantity test.py / challenge.py consume)."""
    return logits.topk(5, dim=-1).indices


def marginalize_verb_noun(logits: torch.Tensor, verb_in_action: torch.Tensor, noun_in_action: torch.Tensor, k: int = 5):
    """Restatement of reference challenge.py:196-210 (softmax over action logits, marginalisation with the 0/1
    class_mappings matrices of datasets/epic_kitchens.py:87-106) and the ranking of common/utils.py:19-42
    (`scores.argsort()[:, ::-1][:, :k]`).  Returns verb scores, noun scores, and top-k indices (action, verb, noun)."""
    probs = _softmax(logits.double()).float()
    verb = probs @ verb_in_action.float()
    noun = probs @ noun_in_action.float()

    def rank(x):
        return torch.argsort(x, dim=-1, descending=True, stable=True)[:, :k]
    return verb, noun, torch.stack([rank(logits), rank(verb), rank(noun)], dim=1), probs
