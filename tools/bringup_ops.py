"""GPU bring-up for the stateless operators: runs each case group in its own subprocess (so a trap
or a hang in one kernel cannot take the others down), compares with torch references, and prints
one line per case.  Usage on the GPU box:  python tools/bringup_ops.py [group ...]
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

GROUPS = ["gemm_2cta", "gemm_basic", "gemm_shapes", "gemm_epilogue", "gemm_strict", "layernorm", "attention", "gemm_perf"]


def _ref_mm(a, w):
    import torch
    return (a.double() @ w.double().t()).float()


def run_group(name):
    import torch
    torch.backends.cuda.matmul.allow_tf32 = False
    from afft_b200 import _capi as capi
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(1234)

    def randn(*shape, scale=1.0):
        return (torch.randn(*shape, generator=g) * scale).to(dev)

    def report(case, err, tol, extra=None):
        rec = {"group": name, "case": case, "max_err": float(err), "tol": tol, "ok": bool(err <= tol)}
        if extra:
            rec.update(extra)
        print(json.dumps(rec), flush=True)

    if name in ("gemm_basic", "gemm_shapes", "gemm_2cta"):
        shapes = [(256, 256, 64, 512), (256, 256, 128, 512), (256, 512, 1024, 512), (1000, 1024, 1024, 512), (300, 3806, 1024, 512),
                  (131, 106, 1024, 512), (576, 1024, 352, 512), (4608, 2048, 8192, 512), (23040, 3072, 1024, 512),
                  (18, 2048, 1024, 512)] if name == "gemm_2cta" else [(128, 128, 64, 128), (128, 256, 64, 256), (128, 256, 128, 256), (256, 512, 1024, 256)] if name == "gemm_basic" else [
            (1000, 1024, 1024, 0), (90, 3072, 1024, 0), (18, 2048, 1024, 0), (576, 1024, 352, 0), (300, 3806, 1024, 0),
            (5760, 1024, 4096, 0), (4608, 8192, 2048, 256), (4608, 2048, 8192, 128), (23040, 3072, 1024, 0), (131, 106, 1024, 0)]
        for (M, N, K, bn) in shapes:
            a = randn(M, K).bfloat16()
            w = randn(N, K, scale=0.05).bfloat16()
            ldo = (N + 3) // 4 * 4
            out = torch.full((M, ldo), float("nan"), device=dev)
            capi.gemm(a, w, out_f32=out, force_block_n=bn)
            torch.cuda.synchronize()
            ref = _ref_mm(a, w)
            err = (out[:, :N] - ref).abs().max().item()
            nan = torch.isnan(out[:, :N]).sum().item()
            report(f"M{M}_N{N}_K{K}_bn{bn}", err if nan == 0 else float("inf"), 2e-3 * (K / 1024) ** 0.5 + 1e-4, {"nan": nan})
    elif name == "gemm_epilogue":
        M, N, K = 777, 1024, 1024
        a = randn(M, K).bfloat16()
        w = randn(N, K, scale=0.05).bfloat16()
        bias = randn(N)
        res = randn(M, N)
        ref0 = _ref_mm(a, w)
        # bias + gelu_erf -> bf16
        out_b = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
        capi.gemm(a, w, bias=bias, act=capi.ACT_GELU_ERF, out_hi=out_b)
        ref = torch.nn.functional.gelu(ref0 + bias)
        report("bias_gelu_erf_bf16", (out_b.float() - ref).abs().max().item(), 3e-2)
        out_f = torch.zeros(M, N, device=dev)
        capi.gemm(a, w, bias=bias, act=capi.ACT_GELU_ERF, out_f32=out_f)
        report("bias_gelu_erf_f32", (out_f - ref).abs().max().item(), 3e-3)
        capi.gemm(a, w, bias=bias, act=capi.ACT_GELU_TANH, out_f32=out_f)
        ref = torch.nn.functional.gelu(ref0 + bias, approximate="tanh")
        report("bias_gelu_tanh_f32", (out_f - ref).abs().max().item(), 3e-3)
        # bias + residual in place
        h = res.clone()
        capi.gemm(a, w, bias=bias, res=h, out_f32=h)
        report("bias_residual_inplace", (h - (ref0 + bias + res)).abs().max().item(), 3e-3)
        # res_mod (position embedding rows) + row mapping
        T = 7
        pos = randn(T, N)
        outm = torch.full((M // T * (T + 1) + T + 1, N), 0.0, device=dev)
        capi.gemm(a, w, res=pos, res_mod=T, out_f32=outm, row_map=(T, T + 1, 1))
        r = torch.arange(M, device=dev)
        orow = (r // T) * (T + 1) + r % T + 1
        refm = ref0 + pos[r % T]
        report("res_mod_rowmap", (outm[orow] - refm).abs().max().item(), 3e-3)
        # strided output (token slot write) + f32 and bf16 at once
        slots = 5
        hbuf = torch.zeros(M * slots, N, device=dev)
        view = hbuf.view(M, slots * N)[:, 2 * N:3 * N]
        ob = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
        capi.gemm(a, w, out_f32=view, out_hi=ob)
        report("strided_out", (hbuf.view(M, slots, N)[:, 2] - ref0).abs().max().item(), 3e-3)
        report("strided_out_others_zero", hbuf.view(M, slots, N)[:, [0, 1, 3, 4]].abs().max().item(), 0.0)
        report("dual_out_bf16", (ob.float() - ref0).abs().max().item(), 3e-2)
        # N tail with bias (classifier shape), padded pitch
        Nc = 3806
        wc = randn(Nc, K, scale=0.05).bfloat16()
        bc = randn(Nc + 2)[:Nc].contiguous()
        bc16 = torch.zeros(Nc + 16, device=dev)[:Nc]
        bc16.copy_(bc)
        oc = torch.full((M, 3808), 7.0, device=dev)
        capi.gemm(a, wc, bias=bc16, out_f32=oc)
        refc = _ref_mm(a, wc) + bc
        report("ntail_bias", (oc[:, :Nc] - refc).abs().max().item(), 3e-3)
        report("ntail_pad_untouched", (oc[:, Nc:] - 7.0).abs().max().item(), 0.0)
    elif name == "gemm_strict":
        for (M, N, K, bn) in [(256, 256, 1024, 128), (777, 1024, 1024, 256), (300, 3806, 1024, 0), (576, 1024, 352, 0)]:
            a32 = randn(M, K)
            w32 = randn(N, K, scale=0.05)
            a_hi, a_lo = capi.split_bf16(a32)
            w_hi, w_lo = capi.split_bf16(w32)
            ldo = (N + 3) // 4 * 4
            out = torch.zeros(M, ldo, device=dev)
            capi.gemm(a_hi, w_hi, a_lo=a_lo, w_lo=w_lo, out_f32=out, force_block_n=bn)
            ref = _ref_mm(a32, w32)
            report(f"strict_M{M}_N{N}_K{K}_bn{bn}", (out[:, :N] - ref).abs().max().item(), 2e-4)
            oh = torch.zeros(M, ldo, device=dev, dtype=torch.bfloat16)
            ol = torch.zeros(M, ldo, device=dev, dtype=torch.bfloat16)
            capi.gemm(a_hi, w_hi, a_lo=a_lo, w_lo=w_lo, out_hi=oh, out_lo=ol, force_block_n=bn)
            report(f"strict_hilo_out_M{M}_N{N}", ((oh.float() + ol.float())[:, :N] - ref).abs().max().item(), 2e-4)
    elif name == "layernorm":
        for dim, eps in [(1024, 1e-6), (2048, 1e-5), (512, 1e-6)]:
            rows = 1237
            x = randn(rows, dim) * 3 + 0.5
            gm, bt = randn(dim), randn(dim)
            yf = torch.zeros(rows, dim, device=dev)
            yh = torch.zeros(rows, dim, device=dev, dtype=torch.bfloat16)
            yl = torch.zeros(rows, dim, device=dev, dtype=torch.bfloat16)
            capi.layernorm(x, gm, bt, eps, y_f32=yf, y_hi=yh, y_lo=yl)
            ref = torch.nn.functional.layer_norm(x.double(), (dim,), gm.double(), bt.double(), eps).float()
            report(f"ln_{dim}_f32", (yf - ref).abs().max().item(), 2e-5)
            report(f"ln_{dim}_hi", (yh.float() - ref).abs().max().item(), 5e-2)
            report(f"ln_{dim}_hilo", (yh.float() + yl.float() - ref).abs().max().item(), 2e-4)
        # token-0 selection with aux scatter (SA-Fuser final norm)
        B, T, n, dim = 5, 6, 4, 1024
        x = randn(B * T * n, dim)
        gm, bt = randn(dim), randn(dim)
        yf = torch.zeros(B * T, dim, device=dev)
        aux = torch.zeros(B * (T + 1), dim, device=dev)
        capi.layernorm(x, gm, bt, 1e-6, rows=B * T, ldx=n * dim, y_f32=yf, aux=(T, T + 1), aux_f32=aux)
        ref = torch.nn.functional.layer_norm(x.view(B * T, n, dim)[:, 0], (dim,), gm, bt, 1e-6)
        report("ln_token0", (yf - ref).abs().max().item(), 2e-5)
        report("ln_aux", (aux.view(B, T + 1, dim)[:, 0] - ref.view(B, T, dim)[:, 0]).abs().max().item(), 2e-5)
        # mean over slots (CMFuser)
        capi.layernorm(x, gm, bt, 1e-6, rows=B * T, ldx=dim, y_f32=yf, in_map=(1, n), avg=(n, 1))
        ref = torch.nn.functional.layer_norm(x.view(B * T, n, dim), (dim,), gm, bt, 1e-6).mean(1)
        report("ln_avg", (yf - ref).abs().max().item(), 2e-5)
    elif name == "attention":
        def ref_attn(qkv, n_seq, L, H, hd, mask, T):
            D = H * hd
            x = qkv.float().view(n_seq, L, 3, H, hd).permute(2, 0, 3, 1, 4)
            q, k, v = x[0], x[1], x[2]
            s = (q @ k.transpose(-1, -2)) * hd ** -0.5
            i = torch.arange(L, device=dev)[:, None]
            j = torch.arange(L, device=dev)[None, :]
            if mask == 1:
                s = s.masked_fill(j > i, float("-inf"))
            elif mask == 2:
                s = s.masked_fill((j % T) > (i % T), float("-inf"))
            elif mask == 3:
                s = s.masked_fill(j == i, float("-inf"))
            p = s.softmax(-1)
            o = (p @ v).transpose(1, 2).reshape(n_seq * L, D)
            return o, p
        for (n_seq, L, H, hd, mask, T, dt) in [(36, 5, 4, 256, 0, 1, torch.bfloat16), (7, 18, 4, 512, 1, 18, torch.bfloat16),
                                               (3, 50, 4, 256, 2, 10, torch.bfloat16), (36, 5, 4, 256, 3, 1, torch.float32),
                                               (7, 18, 4, 512, 1, 18, torch.float32), (3, 50, 4, 256, 2, 10, torch.float32)]:
            D = H * hd
            qkv = randn(n_seq * L, 3 * D).to(dt)
            oh = torch.zeros(n_seq * L, D, device=dev, dtype=torch.bfloat16)
            ol = torch.zeros(n_seq * L, D, device=dev, dtype=torch.bfloat16)
            probs = torch.zeros(n_seq, H, L, L, device=dev)
            capi.attention(qkv, n_seq, L, H, hd, mask=mask, T=T, out_hi=oh, out_lo=ol, probs=probs, p_outer=H * L * L)
            ro, rp = ref_attn(qkv, n_seq, L, H, hd, mask, T)
            tag = f"attn_L{L}_hd{hd}_m{mask}_{'f32' if dt == torch.float32 else 'bf16'}"
            report(tag + "_out", (oh.float() + ol.float() - ro).abs().max().item(), 2e-4)
            report(tag + "_probs", (probs - rp).abs().max().item(), 2e-5)
    elif name == "gemm_perf":
        def bench(M, N, K, bn, strict=False, epi=None):
            a = randn(M, K).bfloat16()
            w = randn(N, K, scale=0.05).bfloat16()
            kw = {}
            if strict:
                kw = dict(a_lo=torch.zeros_like(a), w_lo=torch.zeros_like(w))
            out_b = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
            out_f = torch.zeros(M, N, device=dev)
            bias = randn(N)
            if epi == "gelu":
                kw.update(bias=bias, act=capi.ACT_GELU_ERF, out_hi=out_b)
            elif epi == "res":
                kw.update(bias=bias, res=out_f, out_f32=out_f)
            else:
                kw.update(out_hi=out_b)
            for _ in range(3):
                capi.gemm(a, w, force_block_n=bn, **kw)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            iters = 20
            e0.record()
            for _ in range(iters):
                capi.gemm(a, w, force_block_n=bn, **kw)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            tf = 2.0 * M * N * K / ms / 1e9
            # cuBLAS reference
            for _ in range(3):
                torch.matmul(a, w.t())
            e0.record()
            for _ in range(iters):
                torch.matmul(a, w.t())
            e1.record()
            torch.cuda.synchronize()
            ms_ref = e0.elapsed_time(e1) / iters
            print(json.dumps({"group": name, "case": f"M{M}_N{N}_K{K}_bn{bn}_{'strict' if strict else 'bf16'}_{epi}",
                              "ms": round(ms, 4), "tflops": round(tf, 1), "cublas_ms": round(ms_ref, 4),
                              "cublas_tflops": round(2.0 * M * N * K / ms_ref / 1e9, 1), "ok": True}), flush=True)
        for bn in (512, 256):
            bench(23040, 3072, 1024, bn)
            bench(23040, 4096, 1024, bn, epi="gelu")
            bench(23040, 1024, 4096, bn, epi="res")
            bench(23040, 1024, 1024, bn, epi="res")
            bench(4608, 8192, 2048, bn, epi="gelu")
            bench(4608, 2048, 8192, bn, epi="res")
            bench(4608, 6144, 2048, bn)
        bench(8192, 8192, 8192, 256)
        bench(8192, 8192, 8192, 512)
        bench(23040, 4096, 1024, 256, strict=True)
        bench(23040, 4096, 1024, 512, strict=True)
    else:
        raise SystemExit(f"unknown group {name}")


def main():
    if len(sys.argv) >= 3 and sys.argv[1] == "--child":
        run_group(sys.argv[2])
        return
    groups = sys.argv[1:] or GROUPS
    failed = 0
    for gname in groups:
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", gname], capture_output=True,
                               text=True, timeout=300)
            out, errtxt, rc = p.stdout, p.stderr, p.returncode
        except subprocess.TimeoutExpired as e:
            out, errtxt, rc = (e.stdout or b"").decode() if isinstance(e.stdout, bytes) else (e.stdout or ""), "TIMEOUT", -9
        sys.stdout.write(out)
        bad = [l for l in out.splitlines() if '"ok": false' in l]
        status = "OK" if rc == 0 and not bad else "FAIL"
        if status == "FAIL":
            failed += 1
        print(f"== {gname}: {status} rc={rc} {time.time() - t0:.1f}s bad={len(bad)}", flush=True)
        if rc != 0:
            print(errtxt[-3000:], flush=True)
    sys.exit(1 if failed else 0)


if __name__ == "__main__":
    main()
