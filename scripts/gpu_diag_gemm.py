"""GEMM bring-up diagnostics for the B200 box: every variant in its own subprocess (a trapped launch
poisons the CUDA context), bounded by a timeout, with an error-structure dump on mismatch.

    python scripts/gpu_diag_gemm.py            # run all variants, write gpurun_out/diag_gemm.log
    python scripts/gpu_diag_gemm.py <variant>  # one variant in-process
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

VARIANTS = [
    "nk_small", "nk_ragged", "nk_bn64", "nk_bn128", "nk_bn192", "nk_bn256", "nk_big", "nk_pair_ragged", "nk_pair_odd",
    "nk_f16", "gelu", "gelu_d", "dmul", "gelu_d_big", "dmul_big", "resid", "resid_stats", "qkv_ln", "dgelu", "acc32", "store32",
    "gelu_big", "resid_big", "resid_stats_big", "dgelu_big", "kn_dgrad_res",        # config-2 shapes: the B-resident schedule
    "kn_dgrad", "kn_dgrad_256", "wgrad", "wgrad_split", "wgrad_192",
    "s2d_w128", "s2d_w32", "s2d_c48", "d2s", "d2s_c48", "bs2d", "bs2d_split", "bs2d_c192", "perf",
]


def report(name, got, ref, tol):
    import torch
    got = got.float()
    ref = ref.float()
    err = (got - ref).abs()
    denom = ref.abs().max().clamp_min(1e-6)
    rel = float(err.max() / denom)
    ok = rel < tol and bool(torch.isfinite(got).all())
    print(f"[{name}] max|err|/max|ref| = {rel:.3e}  (tol {tol})  {'OK' if ok else 'MISMATCH'}", flush=True)
    if not ok:
        bad = err > tol * denom
        rows = bad.any(dim=1).nonzero().flatten()
        cols = bad.any(dim=0).nonzero().flatten()
        print(f"   bad rows: {rows.numel()}/{got.shape[0]} first {rows[:16].tolist()}")
        print(f"   bad cols: {cols.numel()}/{got.shape[1]} first {cols[:16].tolist()}")
        print("   got[:4,:8] ", got[:4, :8].tolist())
        print("   ref[:4,:8] ", ref[:4, :8].tolist())
        nz = float((got != 0).float().mean())
        print(f"   nonzero fraction of output {nz:.3f}; ratio got/ref median "
              f"{float((got / ref.clamp_min(1e-3)).median()):.3f}")
    return ok


def run_variant(v):
    import torch
    from bubbleformer_b200 import _lib as L, ops
    torch.manual_seed(0)
    dev = "cuda"
    dt = torch.float16 if v == "nk_f16" else torch.bfloat16

    def rnd(*s, scale=1.0):
        return (torch.randn(*s, device=dev) * scale).to(dt)

    ok = True
    big = v in ("gelu_big", "resid_big", "resid_stats_big", "dgelu_big", "gelu_d_big", "dmul_big")
    if big:
        v = v[:-4]
    if v.startswith("nk_") or v in ("gelu", "gelu_d", "dmul", "resid", "dgelu", "acc32", "store32"):
        shapes = {"nk_small": (128, 128, 64), "nk_ragged": (300, 200, 104), "nk_bn64": (256, 64, 128),
                  "nk_bn128": (4096, 384, 384), "nk_bn192": (4096, 1152, 384), "nk_bn256": (4096, 1536, 384),
                  "nk_big": (40960, 1152, 384), "nk_f16": (1024, 96, 384),
                  # CTA-pair kernel (M >= 4096): second CTA partly / wholly outside the matrix
                  "nk_pair_ragged": (4200, 384, 200), "nk_pair_odd": (4224, 1152, 384)}
        M, N, K = shapes.get(v, (1000, 384, 256))
        if big:
            M, N, K = (40960, 384, 384) if v == "resid" else (40960, 1536, 384)
        bn = {"nk_bn64": 64, "nk_bn128": 128, "nk_bn192": 192, "nk_bn256": 256}.get(v, 0)
        A, B = rnd(M, K), rnd(N, K, scale=K ** -0.5)
        bias = torch.randn(N, device=dev)
        ref = A.float() @ B.float().t() + bias
        if v.startswith("nk_"):
            out = torch.zeros(M, N, device=dev, dtype=dt)
            rpg = 128 if M % 128 == 0 else 0
            st = torch.zeros(M // rpg, N, 2, device=dev) if rpg else None
            ops.gemm(A, B, M, N, K, epilogue=L.EPI_STORE16, bias=bias, out16=out, bn=bn, rows_per_group=max(rpg, 1), stats_out=st)
            ok &= report(v, out, ref, 1e-2)
            if rpg:       # fused InstanceNorm statistics of the stored values
                oi = out.float().reshape(M // rpg, rpg, N)
                ok &= report(v + ".stats", st, torch.stack([oi.sum(1), (oi * oi).sum(1)], dim=-1), 1e-4)
        elif v == "store32":
            out = torch.zeros(M, N, device=dev)
            ops.gemm(A, B, M, N, K, epilogue=L.EPI_STORE32, bias=bias, out32=out)
            ok &= report(v, out, ref, 1e-5)
        elif v == "gelu":
            out = torch.zeros(M, N, device=dev, dtype=dt)
            pre = torch.zeros(M, N, device=dev, dtype=dt)
            ops.gemm(A, B, M, N, K, epilogue=L.EPI_GELU, bias=bias, out16=out, out16b=pre)
            ok &= report(v + ".pre", pre, ref, 1e-2)
            ok &= report(v + ".act", out, torch.nn.functional.gelu(ref), 1e-2)
        elif v == "gelu_d":
            out = torch.zeros(M, N, device=dev, dtype=dt)
            der = torch.zeros(M, N, device=dev, dtype=dt)
            ops.gemm(A, B, M, N, K, epilogue=L.EPI_GELU_D, bias=bias, out16=out, out16b=der)
            r32 = ref.clone().requires_grad_(True)
            g = torch.nn.functional.gelu(r32)
            g.sum().backward()
            ok &= report(v + ".act", out, g.detach(), 1e-2)
            ok &= report(v + ".der", der, r32.grad, 1e-2)
        elif v == "dmul":
            aux = rnd(M, N)
            out = torch.zeros(M, N, device=dev, dtype=dt)
            cs = torch.zeros(N, device=dev)
            if big:
                ops.gemm(A, B.t().contiguous(), M, N, K, epilogue=L.EPI_DMUL, b_mode=L.B_KN, aux16=aux, out16=out, colsum_out=cs)
            else:
                ops.gemm(A, B, M, N, K, epilogue=L.EPI_DMUL, aux16=aux, out16=out, colsum_out=cs)
            ok &= report(v + ".colsum", cs[None], out.float().sum(0)[None], 2e-3)
            ok &= report(v, out, (A.float() @ B.float().t()) * aux.float(), 1e-2)
        elif v == "resid":
            xin = torch.randn(M, N, device=dev)
            cs, ch, cg = torch.randn(N, device=dev), torch.randn(N, device=dev), torch.randn(N, device=dev)
            rpg = 1024 if big else 100
            rs = torch.rand((M + rpg - 1) // rpg, device=dev)
            out32 = torch.zeros(M, N, device=dev)
            out16 = torch.zeros(M, N, device=dev, dtype=dt)
            z = torch.zeros(M, N, device=dev, dtype=dt)
            ops.gemm(A, B, M, N, K, epilogue=L.EPI_RESID, bias=bias, col_scale=cs, col_shift=ch, col_gamma=cg,
                     row_scale=rs, rows_per_group=rpg, in32=xin, out32=out32, out16=out16, out16b=z)
            rsx = rs.repeat_interleave(rpg)[:M, None]
            want = xin + rsx * cg * (ref * cs + ch)
            ok &= report(v + ".z", z, ref, 1e-2)
            ok &= report(v + ".x32", out32, want, 1e-5)
            ok &= report(v + ".x16", out16, want, 1e-2)
        elif v == "dgelu":
            pre = rnd(M, N)
            out = torch.zeros(M, N, device=dev, dtype=dt)
            cs = torch.zeros(N, device=dev)
            if big:      # as the engine calls it: B stored (K, N)
                ops.gemm(A, B.t().contiguous(), M, N, K, epilogue=L.EPI_DGELU, b_mode=L.B_KN, aux16=pre, out16=out,
                         colsum_out=cs)
            else:
                ops.gemm(A, B, M, N, K, epilogue=L.EPI_DGELU, aux16=pre, out16=out, colsum_out=cs)
            ok &= report(v + ".colsum", cs[None], out.float().sum(0)[None], 2e-3)
            p32 = pre.float().requires_grad_(True)
            torch.nn.functional.gelu(p32).sum().backward()
            ok &= report(v, out, (A.float() @ B.float().t()) * p32.grad, 1e-2)
        elif v == "acc32":
            g = torch.randn(M, N, device=dev)
            out = torch.zeros(M, N, device=dev)
            ops.gemm(A, B, M, N, K, epilogue=L.EPI_ACC32, in32=g, out32=out)
            ok &= report(v, out, g + A.float() @ B.float().t(), 1e-5)
    elif v == "qkv_ln":
        for (M, heads, bn) in [(1000, 2, 0), (4096, 6, 0), (900, 3, 0), (40960, 6, 0), (900, 3, 128), (4096, 6, 128)]:
            N, K, d = heads * 192, 384, 64
            A, B = rnd(M, K), rnd(N, K, scale=K ** -0.5)
            bias = torch.randn(N, device=dev)
            ref = (A.float() @ B.float().t() + bias).reshape(M, heads, 3, d)
            qk = ref[:, :, :2]
            mu, var = qk.mean(-1, keepdim=True), qk.var(-1, unbiased=False, keepdim=True)
            want = ref.clone()
            want[:, :, :2] = (qk - mu) * torch.rsqrt(var + 1e-5)
            out = torch.zeros(M, N, device=dev, dtype=dt)
            rstd = torch.zeros(M, heads, 2, device=dev)
            ops.gemm(A, B, M, N, K, epilogue=L.EPI_QKV_LN, bias=bias, out16=out, ln_head_dim=d, ln_rstd=rstd, bn=bn)
            ok &= report(f"{v} xhat|v M={M} bn={bn}", out, want.reshape(M, N), 1e-2)
            ok &= report(f"{v} rstd M={M}", rstd.reshape(M, -1), torch.rsqrt(var + 1e-5).reshape(M, -1), 1e-3)
    elif v == "resid_stats":
        M, N, K, rpg = (40960, 384, 384, 1024) if big else (1024, 384, 384, 256)
        A, B = rnd(M, K), rnd(N, K, scale=K ** -0.5)
        bias = torch.randn(N, device=dev)
        xin = torch.randn(M, N, device=dev)
        cg = torch.randn(N, device=dev)
        rs = torch.rand(M // rpg, device=dev)
        out32 = torch.zeros(M, N, device=dev)
        st = torch.zeros(M // rpg, N, 2, device=dev)
        ops.gemm(A, B, M, N, K, epilogue=L.EPI_RESID, bias=bias, col_gamma=cg, row_scale=rs, rows_per_group=rpg,
                 in32=xin, out32=out32, stats_out=st)
        want = xin + rs.repeat_interleave(rpg)[:, None] * cg * (A.float() @ B.float().t() + bias)
        ok &= report(v + ".x32", out32, want, 1e-5)
        wi = want.reshape(M // rpg, rpg, N)
        ok &= report(v + ".stats", st, torch.stack([wi.sum(1), (wi * wi).sum(1)], dim=-1), 1e-4)
    elif v.startswith("kn_dgrad"):
        M, N, K = {"kn_dgrad": (4096, 384, 1536), "kn_dgrad_res": (40960, 384, 384)}.get(v, (2048, 1536, 384))
        A, Bkn = rnd(M, K), rnd(K, N, scale=K ** -0.5)          # B stored (K, N)
        out = torch.zeros(M, N, device=dev, dtype=dt)
        ops.gemm(A, Bkn, M, N, K, epilogue=L.EPI_STORE16, b_mode=L.B_KN, out16=out)
        ok &= report(v, out, A.float() @ Bkn.float(), 1e-2)
    elif v.startswith("wgrad"):
        T = 4096                                                  # tokens = contraction
        Mw, Nw = (384, 1152) if v == "wgrad_192" else (384, 384)
        dY, X = rnd(T, Mw), rnd(T, Nw, scale=T ** -0.5)
        out = torch.zeros(Mw, Nw, device=dev)
        ops.gemm(dY, X, Mw, Nw, T, epilogue=L.EPI_ATOMIC32, a_mode=L.A_KM, b_mode=L.B_KN,
                 split_k=8 if v == "wgrad_split" else 1, out32=out,
                 bn=192 if v == "wgrad_192" else 0)
        ok &= report(v, out, dY.float().t() @ X.float(), 1e-4)
    elif v.startswith("bs2d"):
        # weight gradient of a 2x2/s2 conv stage: dW[co, (ky,kx,ci)] = sum_pixels dY[pix, co] * patch[pix, (ky,kx,ci)] with
        # B = the implicit patch gather of the image tensor (no gathered copy)
        I, Hin, Win, Cin, Co, tA, sk = {"bs2d": (2, 8, 256, 96, 96, torch.bfloat16, 1),
                                        "bs2d_split": (3, 16, 128, 96, 96, torch.bfloat16, 5),
                                        "bs2d_c192": (1, 8, 128, 192, 192, torch.float16, 3)}[v]
        img = (torch.randn(I, Hin, Win, Cin, device=dev)).to(tA)
        M = I * (Hin // 2) * (Win // 2)
        dY = (torch.randn(M, Co, device=dev) * M ** -0.5).to(tA)
        out = torch.zeros(Co, 4 * Cin, device=dev)
        ops.gemm(dY, img.view(-1, Cin), Co, 4 * Cin, M, epilogue=L.EPI_ATOMIC32, a_mode=L.A_KM, b_mode=L.B_KN_S2D,
                 s2d=(I, Hin, Win, Cin), split_k=sk, out32=out)
        g = img.float().reshape(I, Hin // 2, 2, Win // 2, 2, Cin).permute(0, 1, 3, 2, 4, 5).reshape(M, 4 * Cin)
        ok &= report(v, out, dY.float().t() @ g, 1e-4)
    elif v.startswith("s2d"):
        I, Hin, Win, Cin, N = {"s2d_w128": (2, 8, 256, 96, 96), "s2d_w32": (3, 64, 64, 96, 384),
                               "s2d_c48": (2, 16, 32, 48, 48)}[v]
        img = rnd(I, Hin, Win, Cin)
        Wt = rnd(N, 4 * Cin, scale=(4 * Cin) ** -0.5)            # K order (ky, kx, ci)
        M = I * (Hin // 2) * (Win // 2)
        out = torch.zeros(M, N, device=dev, dtype=dt)
        ops.gemm(img, Wt, M, N, 4 * Cin, epilogue=L.EPI_STORE16, a_mode=L.A_S2D, ldb=4 * Cin,
                 s2d=(I, Hin, Win, Cin), out16=out)
        g = img.float().reshape(I, Hin // 2, 2, Win // 2, 2, Cin).permute(0, 1, 3, 2, 4, 5).reshape(M, 4 * Cin)
        ok &= report(v, out, g @ Wt.float().t(), 1e-2)
    elif v.startswith("d2s"):
        I, h, w, Cin, co = (2, 8, 16, 384, 96) if v == "d2s" else (1, 4, 8, 192, 48)
        X = rnd(I * h * w, Cin)
        Wt = rnd(4 * co, Cin, scale=Cin ** -0.5)                  # rows ordered (ky, kx, co)
        out = torch.zeros(I, 2 * h, 2 * w, co, device=dev, dtype=dt)
        ops.gemm(X, Wt, I * h * w, 4 * co, Cin, epilogue=L.EPI_D2S, d2s=(h, w, co), out16=out, ldo=4 * co)
        y = (X.float() @ Wt.float().t()).reshape(I, h, w, 2, 2, co).permute(0, 1, 3, 2, 4, 5).reshape(I, 2 * h, 2 * w, co)
        ok &= report(v, out.reshape(-1, co), y.reshape(-1, co), 1e-2)
    elif v == "perf":
        for (M, N, K, bn) in [(40960, 1152, 384, 192), (40960, 1152, 384, 128), (40960, 1536, 384, 256),
                              (40960, 1536, 384, 128), (40960, 384, 1536, 192), (40960, 384, 1536, 128),
                              (40960, 384, 384, 192), (40960, 384, 384, 128), (163840, 2304, 768, 256)]:
            A, B = rnd(M, K), rnd(N, K, scale=K ** -0.5)
            out = torch.zeros(M, N, device=dev, dtype=dt)
            for _ in range(3):
                ops.gemm(A, B, M, N, K, epilogue=L.EPI_STORE16, out16=out, bn=bn)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                ops.gemm(A, B, M, N, K, epilogue=L.EPI_STORE16, out16=out, bn=bn)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            print(f"[perf] M={M} N={N} K={K} bn={bn}: {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)
            ref = torch.matmul(A, B.t())
            for _ in range(3):
                torch.matmul(A, B.t(), out=ref)
            e0.record()
            for _ in range(20):
                torch.matmul(A, B.t(), out=ref)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            print(f"[perf]   cuBLAS same shape: {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)
    else:
        raise SystemExit(f"unknown variant {v}")
    torch.cuda.synchronize()
    return ok


def main():
    if len(sys.argv) > 1:
        ok = run_variant(sys.argv[1])
        sys.exit(0 if ok else 3)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    log = open(os.path.join(ROOT, "gpurun_out", "diag_gemm.log"), "w")
    summary = []
    for v in VARIANTS:
        try:
            r = subprocess.run([sys.executable, __file__, v], capture_output=True, text=True, timeout=180)
            out, code = r.stdout + r.stderr, r.returncode
        except subprocess.TimeoutExpired as e:
            out, code = (e.stdout or b"").decode() + (e.stderr or b"").decode() + "\nTIMEOUT", -9
        tail = "\n".join(out.strip().splitlines()[-14:])
        log.write(f"==== {v} (exit {code})\n{out}\n")
        log.flush()
        print(f"==== {v} (exit {code})\n{tail}", flush=True)
        summary.append((v, code))
    print("SUMMARY " + " ".join(f"{v}:{c}" for v, c in summary))
    log.write("SUMMARY " + " ".join(f"{v}:{c}" for v, c in summary) + "\n")


if __name__ == "__main__":
    main()
