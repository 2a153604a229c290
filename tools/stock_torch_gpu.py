"""Stock PyTorch on the same B200 ("the stock PyTorch on the same box bar").

A plain torch.nn.functional implementation of the headline path (SA-Fuser EK100 R-TSN+O+AU+F, 4h_18s) - the
library calls the reference module makes (F.linear / F.layer_norm / F.gelu / softmax attention), random weights
of the right shapes - timed with CUDA events in fp32 (TF32 off), fp32 with TF32, and bf16 autocast.  It shares no
code with afft_b200 or oracle/: it only answers "what does eager PyTorch deliver on this GPU for this model".

    python tools/stock_torch_gpu.py [--batch 256] [--steps 10]
"""
import argparse
import json
import math

import torch
import torch.nn.functional as F


def make_weights(dev, D=1024, G=2048, depth=6, layers=6, C=3806):
    g = torch.Generator(device="cpu").manual_seed(0)

    def w(*shape):
        return (torch.randn(*shape, generator=g) * 0.02).to(dev)
    W = {"map_obj": w(D, 352), "token": w(1, 1, D), "enc": w(G, D), "dec": w(D, G), "wpe": w(1024, G), "cls_w": w(C, D),
         "cls_b": w(C), "fuser": [], "gpt": [], "norm": (torch.ones(D, device=dev), torch.zeros(D, device=dev)),
         "ln_f": (torch.ones(G, device=dev), torch.zeros(G, device=dev))}
    for _ in range(depth):
        W["fuser"].append(dict(n1=(torch.ones(D, device=dev), torch.zeros(D, device=dev)), qkv=w(3 * D, D), proj=w(D, D),
                               proj_b=w(D), n2=(torch.ones(D, device=dev), torch.zeros(D, device=dev)), fc1=w(4 * D, D),
                               fc1_b=w(4 * D), fc2=w(D, 4 * D), fc2_b=w(D)))
    for _ in range(layers):
        W["gpt"].append(dict(n1=(torch.ones(G, device=dev), torch.zeros(G, device=dev)), qkv=w(3 * G, G), qkv_b=w(3 * G),
                             proj=w(G, G), proj_b=w(G), n2=(torch.ones(G, device=dev), torch.zeros(G, device=dev)),
                             fc1=w(4 * G, G), fc1_b=w(4 * G), fc2=w(G, 4 * G), fc2_b=w(G)))
    return W


def attn(x, qkv_w, qkv_b, H, causal):
    B, N, C = x.shape
    qkv = F.linear(x, qkv_w, qkv_b).reshape(B, N, 3, H, C // H).permute(2, 0, 3, 1, 4)
    s = (qkv[0] @ qkv[1].transpose(-2, -1)) * (C // H) ** -0.5
    if causal:
        s = s + torch.triu(torch.full((N, N), float("-inf"), device=x.device, dtype=s.dtype), diagonal=1)
    return (s.softmax(-1) @ qkv[2]).transpose(1, 2).reshape(B, N, C)


def forward(W, feats, T=18, H=4):
    B = feats[0].shape[0]
    D = 1024
    toks = [feats[0], F.linear(feats[1], W["map_obj"]), feats[2], feats[3]]
    x = torch.cat([W["token"].expand(B * T, -1, -1)] + [t.reshape(B * T, 1, D) for t in toks], dim=1)
    for b in W["fuser"]:
        x = x + F.linear(attn(F.layer_norm(x, (D,), *b["n1"], 1e-6), b["qkv"], None, H, False), b["proj"], b["proj_b"])
        x = x + F.linear(F.gelu(F.linear(F.layer_norm(x, (D,), *b["n2"], 1e-6), b["fc1"], b["fc1_b"])), b["fc2"], b["fc2_b"])
    z = F.layer_norm(x, (D,), *W["norm"], 1e-6)[:, 0].reshape(B, T, D)
    g = F.linear(z, W["enc"]) + W["wpe"][:T]
    G = g.shape[-1]
    for b in W["gpt"]:
        g = g + F.linear(attn(F.layer_norm(g, (G,), *b["n1"], 1e-5), b["qkv"], b["qkv_b"], H, True), b["proj"], b["proj_b"])
        g = g + F.linear(F.gelu(F.linear(F.layer_norm(g, (G,), *b["n2"], 1e-5), b["fc1"], b["fc1_b"]), approximate="tanh"),
                         b["fc2"], b["fc2_b"])
    zh = F.linear(F.layer_norm(g, (G,), *W["ln_f"], 1e-5), W["dec"])
    pf = torch.cat([z[:, :1], zh[:, :T - 1]], dim=1)
    return F.linear(zh[:, T - 1:], W["cls_w"], W["cls_b"]), F.linear(pf, W["cls_w"], W["cls_b"])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--steps", type=int, default=10)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    W = make_weights(dev)
    B, T = args.batch, 18
    feats = [torch.randn(B, T, d, device=dev) for d in (1024, 352, 1024, 1024)]
    res = {"batch": B, "config": "ek100_sa_tsn", "gpu": torch.cuda.get_device_name(0), "torch": torch.__version__}
    for mode in ("fp32", "tf32", "bf16_autocast"):
        torch.backends.cuda.matmul.allow_tf32 = mode == "tf32"
        torch.backends.cudnn.allow_tf32 = mode == "tf32"
        ctx = torch.autocast("cuda", dtype=torch.bfloat16) if mode == "bf16_autocast" else torch.autocast("cuda", enabled=False)
        with torch.no_grad(), ctx:
            for _ in range(3):
                forward(W, feats)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                forward(W, feats)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        res[mode] = {"ms_per_step": round(ms, 3), "clips_per_s": round(B / ms * 1e3, 1)}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
