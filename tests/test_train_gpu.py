"""GPU: the training-step path (forward + backward through the native kernels) against torch autograd through the
CPU oracle on the same seeded inputs, all stochastic rates set to 0 (the only setting in which forward/backward
parity is checkable - SURVEY.md section 7).  Tolerances: bf16 GEMM operands, fp32 everything else:
forward as the bf16 inference mode; gradients: relative L2 error < 5e-2 and cosine > 0.995 per parameter."""
import os

import pytest
import torch

from afft_b200 import _capi, configs, synthetic
from afft_b200 import train as atrain
from afft_b200.models import BaseModel

pytestmark = pytest.mark.gpu
KW = dict(mixup_fn=None, target=None, target_subclips=None, target_subclips_ignore_index=None)


def _no_dropout_cfg(name):
    cfg, T, ncls, _ = configs.named_config(name)
    cfg["dropout"] = 0.0
    for k in ("embd_drop_rate", "drop_rate", "attn_drop_rate", "drop_path_rate"):
        cfg["fuser"][k] = 0.0
    for k in ("embd_pdrop", "resid_pdrop", "attn_pdrop"):
        cfg["common"][k] = 0.0
        cfg["future_predictor"][k] = 0.0
    return cfg, T, ncls


def test_backward_operators_against_torch():
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    # LinearFn (nn.Linear and Conv1D layouts, bias, odd N)
    for conv1d, (M, K, N) in ((False, (264, 352, 1024)), (True, (264, 1024, 3072)), (False, (272, 1024, 3806))):
        x = torch.randn(M, K, generator=g).to(dev).requires_grad_()
        w = (torch.randn(*((K, N) if conv1d else (N, K)), generator=g) * 0.05).to(dev).requires_grad_()
        b = torch.randn(N, generator=g).to(dev).requires_grad_()
        y = atrain.LinearFn.apply(x, w, b, conv1d)
        ref = (x @ w + b) if conv1d else torch.nn.functional.linear(x, w, b)
        assert (y - ref).abs().max().item() < 5e-2
        dy = torch.randn(M, N, generator=g).to(dev)
        gx, gw, gb = torch.autograd.grad(y, (x, w, b), dy)
        rx, rw, rb = torch.autograd.grad(ref, (x, w, b), dy)
        for a_, r_ in ((gx, rx), (gw, rw), (gb, rb)):
            assert ((a_ - r_).norm() / r_.norm()).item() < 1e-2
    # LayerNorm
    x = (torch.randn(300, 1024, generator=g) * 2 + 0.3).to(dev).requires_grad_()
    gm = torch.randn(1024, generator=g).to(dev).requires_grad_()
    bt = torch.randn(1024, generator=g).to(dev).requires_grad_()
    y = atrain.LayerNormFn.apply(x, gm, bt, 1e-6)
    ref = torch.nn.functional.layer_norm(x, (1024,), gm, bt, 1e-6)
    dy = torch.randn(300, 1024, generator=g).to(dev)
    for a_, r_ in zip(torch.autograd.grad(y, (x, gm, bt), dy), torch.autograd.grad(ref, (x, gm, bt), dy)):
        assert ((a_ - r_).norm() / r_.norm()).item() < 1e-4
    # GELU (erf / tanh)
    for kind, approx in ((_capi.ACT_GELU_ERF, "none"), (_capi.ACT_GELU_TANH, "tanh")):
        x = (torch.randn(257, 130, generator=g) * 2).to(dev).requires_grad_()
        y = atrain.GeluFn.apply(x, kind)
        ref = torch.nn.functional.gelu(x, approximate=approx)
        dy = torch.randn(257, 130, generator=g).to(dev)
        assert (y - ref).abs().max().item() < 1e-5
        assert (torch.autograd.grad(y, x, dy)[0] - torch.autograd.grad(ref, x, dy)[0]).abs().max().item() < 1e-4
    # attention (SA 5 tokens, GPT causal 18)
    for n_seq, L, H, hd, mask in ((24, 5, 4, 256, 0), (6, 18, 4, 512, 1)):
        D = H * hd
        qkv = torch.randn(n_seq * L, 3 * D, generator=g).to(dev).requires_grad_()
        out, probs = atrain.AttentionFn.apply(qkv, n_seq, L, H, hd, mask, L)
        t = qkv.view(n_seq, L, 3, H, hd).permute(2, 0, 3, 1, 4)
        s = (t[0] @ t[1].transpose(-1, -2)) * hd ** -0.5
        if mask == 1:
            s = s + torch.triu(torch.full((L, L), float("-inf"), device=dev), 1)
        ref = (s.softmax(-1) @ t[2]).transpose(1, 2).reshape(n_seq * L, D)
        assert (out - ref).abs().max().item() < 2e-4
        dy = torch.randn(n_seq * L, D, generator=g).to(dev)
        ga, gr = torch.autograd.grad(out, qkv, dy)[0], torch.autograd.grad(ref, qkv, dy)[0]
        assert ((ga - gr).norm() / gr.norm()).item() < 1e-4


def test_training_step_gradients_match_oracle_autograd():
    from oracle import afft_oracle
    cfg, T, ncls = _no_dropout_cfg("egtea_sa")
    B = 8
    model = BaseModel(cfg, ncls, {})
    sd = synthetic.synthetic_state_dict(model, seed=0)
    model.load_state_dict(sd)
    model = model.to("cuda:0").train()
    feats = synthetic.synthetic_features(cfg["modal_dims"], B, T, seed=77)
    gl = torch.Generator().manual_seed(5)
    target = torch.randint(0, 106, (B, 1), generator=gl)
    target_sub = torch.randint(0, 106, (B, T), generator=gl)
    target_sub[0, :3] = -1  # ignored past frames

    out, _ = model({m: t.reshape(B, T, -1, 1, 1, 1).cuda() for m, t in feats.items()}, **KW)
    losses = atrain.reference_losses(out, target.cuda(), target_sub.cuda())
    losses["total"].backward()
    torch.cuda.synchronize()

    sd_ref = {k: v.clone().requires_grad_() for k, v in sd.items()}
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    ref_out = afft_oracle.forward(sd_ref, cfg, ncls, feats)
    ref_losses = atrain.reference_losses(ref_out, target, target_sub)
    ref_losses["total"].backward()

    for k in ("cls_action", "past_cls_action", "past_reg", "total"):
        assert abs(losses[k].item() - ref_losses[k].item()) < 2e-2 * max(1.0, abs(ref_losses[k].item())), k
    assert (out["logits/action"]["all-fused"].detach().cpu() - ref_out["logits/action"]["all-fused"].detach()).abs().max() < 6e-2
    worst = (0.0, None)
    n_checked = 0
    for name, p in model.named_parameters():
        gref = sd_ref[name].grad
        assert p.grad is not None, name
        if gref is None or gref.norm().item() == 0.0:
            continue
        g = p.grad.detach().cpu()
        rel = ((g - gref).norm() / gref.norm()).item()
        cos = torch.nn.functional.cosine_similarity(g.flatten(), gref.flatten(), dim=0).item()
        n_checked += 1
        if rel > worst[0]:
            worst = (rel, name)
        assert rel < 5e-2 and cos > 0.995, (name, rel, cos)
    assert n_checked >= 50, n_checked
    print("worst relative gradient error", worst)


def test_unsupported_training_configs_raise():
    cfg, T, ncls, _ = configs.named_config("ek100_ca")
    m = BaseModel(cfg, ncls, {}).to("cuda:0").train()
    x = {k: torch.zeros(8, T, d, 1, 1, 1, device="cuda:0") for k, d in cfg["modal_dims"].items()}
    with pytest.raises(NotImplementedError):
        m(x, **KW)
