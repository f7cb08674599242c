"""Micro-benchmark of individual kernels at config-2 shapes (for CUDA-event timing and ncu captures).

    python scripts/micro.py [gemm_qkv gemm_fc1 gemm_resid attn_fwd attn_bwd attn_t inorm_bwd inorm_apply ...] [--iters 5]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from bubbleformer_b200 import _lib as L, engine, ops
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("names", nargs="*")
    ap.add_argument("--iters", type=int, default=5)
    args = ap.parse_args()
    names, iters = args.names, args.iters
    dev = "cuda"
    N, E, I, P, he = 40960, 384, 40, 1024, 6
    g = engine.Geom(8, 5, 32, 32)
    bf = torch.bfloat16
    X32 = torch.randn(N, E, device=dev)
    Xb = torch.randn(N, E, device=dev).to(bf)
    QKV = torch.randn(N, 3 * E, device=dev).to(bf)
    H = torch.randn(N, 4 * E, device=dev).to(bf)
    Win = (torch.randn(3 * E, E, device=dev) * E ** -0.5).to(bf)
    W1 = (torch.randn(4 * E, E, device=dev) * E ** -0.5).to(bf)
    W2 = (torch.randn(E, 4 * E, device=dev) * (4 * E) ** -0.5).to(bf)
    Wo = (torch.randn(E, E, device=dev) * E ** -0.5).to(bf)
    vE = torch.randn(E, device=dev)
    v3E = torch.randn(3 * E, device=dev)
    v4E = torch.randn(4 * E, device=dev)
    rs = torch.ones(I, device=dev)
    st = torch.zeros(I, E, 2, device=dev)
    ops.inorm_stats(X32, I, P, st)
    ln = [torch.ones(64, device=dev), torch.zeros(64, device=dev), torch.ones(64, device=dev), torch.zeros(64, device=dev)]
    emb = torch.randn(32, he, device=dev)
    sf = torch.ones(he, device=dev)
    O = torch.empty(N, E, device=dev, dtype=bf)
    O2 = torch.empty(N, E, device=dev, dtype=bf)
    out3 = torch.empty(N, 3 * E, device=dev, dtype=bf)
    out4 = torch.empty(N, 4 * E, device=dev, dtype=bf)
    out4b = torch.empty(N, 4 * E, device=dev, dtype=bf)
    o32 = torch.empty(N, E, device=dev)
    red = torch.zeros(I, E, 2, device=dev)
    gr = dict(d_qn_w=torch.zeros(64, device=dev), d_qn_b=torch.zeros(64, device=dev), d_kn_w=torch.zeros(64, device=dev),
              d_kn_b=torch.zeros(64, device=dev), d_bias_emb=torch.zeros(32, he, device=dev),
              d_scale_factor=torch.zeros(he, device=dev))

    rstd = torch.rand(N, he, 2, device=dev) + 0.5
    A96 = torch.randn(655360, 96, device=dev).half()
    W96 = (torch.randn(96, 384, device=dev) / 10).half()
    Z96 = torch.empty(4 * 655360, 96, device=dev, dtype=torch.float16)

    def attn(axis, bwd):
        geo = engine._axis(g, axis)
        kw = dict(heads=he, qn_w=ln[0], qn_b=ln[1], kn_w=ln[2], kn_b=ln[3], bias_emb=emb,
                  bucket=engine.relpos_bucket_vector(geo["L_"], dev), scale_factor=sf, out_scale=0.5, **geo)
        if bwd:
            ops.attention(QKV, out3, dout=Xb, grads=gr, prenorm=True, rstd=rstd, **kw)
        else:
            ops.attention(QKV, O, prenorm=True, **kw)

    table = {
        "gemm_qkv": (lambda: ops.gemm(Xb, Win, N, 3 * E, E, epilogue=L.EPI_STORE16, bias=v3E, out16=out3), 2.0 * N * 3 * E * E, (N * E + N * 3 * E) * 2),
        "gemm_qkv_ln": (lambda: ops.gemm(Xb, Win, N, 3 * E, E, epilogue=L.EPI_QKV_LN, bias=v3E, out16=out3, ln_head_dim=64, ln_rstd=rstd),
                        2.0 * N * 3 * E * E, (N * E + N * 3 * E) * 2),
        "gemm_acc32": (lambda: ops.gemm(H, W1, N, E, 4 * E, epilogue=L.EPI_ACC32, b_mode=L.B_KN, in32=X32, out32=o32), 2.0 * N * 4 * E * E, N * 4 * E * 2 + N * E * 8),
        "gemm_dgrad_qkv": (lambda: ops.gemm(QKV, Win, N, E, 3 * E, epilogue=L.EPI_STORE16, b_mode=L.B_KN, out16=O), 2.0 * N * 3 * E * E, (N * E + N * 3 * E) * 2),
        "gemm_fc1": (lambda: ops.gemm(Xb, W1, N, 4 * E, E, epilogue=L.EPI_GELU, bias=v4E, out16=out4, out16b=out4b), 2.0 * N * 4 * E * E, (N * E + 2 * N * 4 * E) * 2),
        "gemm_fc1d": (lambda: ops.gemm(Xb, W1, N, 4 * E, E, epilogue=L.EPI_GELU_D, bias=v4E, out16=out4, out16b=out4b), 2.0 * N * 4 * E * E, (N * E + 2 * N * 4 * E) * 2),
        "gemm_dmul": (lambda: ops.gemm(Xb, W2, N, 4 * E, E, epilogue=L.EPI_DMUL, b_mode=L.B_KN, aux16=H, out16=out4, colsum_out=v4E), 2.0 * N * 4 * E * E, (N * E + 2 * N * 4 * E) * 2),
        "gemm_fc2": (lambda: ops.gemm(H, W2, N, E, 4 * E, epilogue=L.EPI_STORE16, bias=vE, out16=O), 2.0 * N * 4 * E * E, (N * 4 * E + N * E) * 2),
        "gemm_resid": (lambda: ops.gemm(Xb, Wo, N, E, E, epilogue=L.EPI_RESID, bias=vE, col_gamma=vE, row_scale=rs, rows_per_group=P,
                                        in32=X32, out32=o32, out16=O, out16b=O2), 2.0 * N * E * E, N * E * (2 + 4 + 4 + 2 + 2)),
        "gemm_resid_stats": (lambda: ops.gemm(Xb, Wo, N, E, E, epilogue=L.EPI_RESID, bias=vE, col_gamma=vE, row_scale=rs, rows_per_group=P,
                                        in32=X32, out32=o32, out16b=O2, stats_out=st), 2.0 * N * E * E, N * E * (2 + 4 + 4 + 2)),
        "gemm_resid_nostats": (lambda: ops.gemm(Xb, Wo, N, E, E, epilogue=L.EPI_RESID, bias=vE, col_gamma=vE, row_scale=rs, rows_per_group=P,
                                        in32=X32, out32=o32, out16b=O2), 2.0 * N * E * E, N * E * (2 + 4 + 4 + 2)),
        "gemm_resid64": (lambda: ops.gemm(Xb, Wo, N, E, E, epilogue=L.EPI_RESID, bias=vE, col_gamma=vE, row_scale=rs, rows_per_group=P,
                                          in32=X32, out32=o32, out16=O, out16b=O2, bn=64), 2.0 * N * E * E, N * E * (2 + 4 + 4 + 2 + 2)),
        "gemm_resid_stats64": (lambda: ops.gemm(Xb, Wo, N, E, E, epilogue=L.EPI_RESID, bias=vE, col_gamma=vE, row_scale=rs, rows_per_group=P,
                                        in32=X32, out32=o32, out16b=O2, stats_out=st, bn=64), 2.0 * N * E * E, N * E * (2 + 4 + 4 + 2)),
        "gemm_qkv_ln128": (lambda: ops.gemm(Xb, Win, N, 3 * E, E, epilogue=L.EPI_QKV_LN, bias=v3E, out16=out3, ln_head_dim=64, ln_rstd=rstd, bn=128),
                           2.0 * N * 3 * E * E, (N * E + N * 3 * E) * 2),
        "gemm_resid_noz": (lambda: ops.gemm(Xb, Wo, N, E, E, epilogue=L.EPI_RESID, bias=vE, col_gamma=vE, row_scale=rs, rows_per_group=P,
                                          in32=X32, out32=o32), 2.0 * N * E * E, N * E * (2 + 4 + 4)),
        "gemm_fc2_bn128": (lambda: ops.gemm(H, W2, N, E, 4 * E, epilogue=L.EPI_STORE16, bias=vE, out16=O, bn=128), 2.0 * N * 4 * E * E, (N * 4 * E + N * E) * 2),
        "gemm_dgrad_qkv_bn128": (lambda: ops.gemm(QKV, Win, N, E, 3 * E, epilogue=L.EPI_STORE16, b_mode=L.B_KN, out16=O, bn=128), 2.0 * N * 3 * E * E, (N * E + N * 3 * E) * 2),
        "gemm_acc32_bn128": (lambda: ops.gemm(H, W1, N, E, 4 * E, epilogue=L.EPI_ACC32, b_mode=L.B_KN, in32=X32, out32=o32, bn=128), 2.0 * N * 4 * E * E, N * 4 * E * 2 + N * E * 8),
        "gemm_dgrad_out_bn128": (lambda: ops.gemm(Xb, Wo, N, E, E, epilogue=L.EPI_STORE16, b_mode=L.B_KN, out16=O, bn=128), 2.0 * N * E * E, N * E * 4),
        "gemm_dgrad_out": (lambda: ops.gemm(Xb, Wo, N, E, E, epilogue=L.EPI_STORE16, b_mode=L.B_KN, out16=O), 2.0 * N * E * E, N * E * 4),
        "gemm_d2s": (lambda: ops.gemm(A96, W96, 655360, 384, 96, epilogue=L.EPI_D2S, b_mode=L.B_KN, d2s=(128, 128, 96), out16=Z96, ldo=384),
                     2.0 * 655360 * 384 * 96, 655360 * (96 + 384) * 2),
        "gemm_wgrad_qkv": (lambda: ops.gemm(QKV, Xb, 3 * E, E, N, epilogue=L.EPI_ATOMIC32, a_mode=L.A_KM, b_mode=L.B_KN,
                                            split_k=engine.pick_split(N, 3 * E, E), out32=torch.zeros(3 * E, E, device=dev)), 2.0 * N * 3 * E * E, (N * 4 * E) * 2),
        "gemm_wgrad_out": (lambda: ops.gemm(Xb, Xb, E, E, N, epilogue=L.EPI_ATOMIC32, a_mode=L.A_KM, b_mode=L.B_KN,
                                            split_k=engine.pick_split(N, E, E), out32=torch.zeros(E, E, device=dev)), 2.0 * N * E * E, (N * 2 * E) * 2),
        "gemm_dgelu": (lambda: ops.gemm(Xb, W2, N, 4 * E, E, epilogue=L.EPI_DGELU, b_mode=L.B_KN, aux16=H, out16=out4), 2.0 * N * 4 * E * E, (N * E + 2 * N * 4 * E) * 2),
        "gemm_wgrad": (lambda: ops.gemm(H, Xb, 4 * E, E, N, epilogue=L.EPI_ATOMIC32, a_mode=L.A_KM, b_mode=L.B_KN, split_k=8,
                                        out32=torch.zeros(4 * E, E, device=dev)), 2.0 * N * 4 * E * E, (N * 4 * E + N * E) * 2),
        "attn_fwd": (lambda: attn("x", False), 0, N * 4 * E * 2),
        "attn_fwd_y": (lambda: attn("y", False), 0, N * 4 * E * 2),
        "attn_bwd": (lambda: attn("x", True), 0, N * 7 * E * 2),
        "attn_t": (lambda: attn("t", False), 0, N * 4 * E * 2),
        "attn_t_bwd": (lambda: attn("t", True), 0, N * 7 * E * 2),
        "inorm_stats": (lambda: ops.inorm_stats(X32, I, P, st), 0, N * E * 4),
        "inorm_apply": (lambda: ops.inorm_apply(X32, O, I, P, st, vE, vE), 0, N * E * 6),
        "inorm_apply_resid": (lambda: ops.inorm_apply(Xb, o32, I, P, st, vE, vE, resid_in=X32, row_scale=rs, col_gamma=vE), 0, N * E * 10),
        "inorm_bwd1": (lambda: ops.inorm_bwd(1, Xb, O, I, P, st, vE, vE, red), 0, N * E * 4),
        "inorm_bwd2": (lambda: ops.inorm_bwd(2, Xb, O, I, P, st, vE, vE, red, out=O2), 0, N * E * 6),
        "inorm_bwd2_add": (lambda: ops.inorm_bwd(2, Xb, X32, I, P, st, vE, vE, red, out=o32, add32=X32), 0, N * E * 14),
        "inorm_fwd_fused": (lambda: ops.inorm_apply(Xb, O, I, P, st, vE, vE, compute_stats=True), 0, N * E * 6),
        "inorm_fwd_fused32": (lambda: ops.inorm_apply(X32, O, I, P, st, vE, vE, compute_stats=True), 0, N * E * 10),
        "inorm_bwd3": (lambda: ops.inorm_bwd(3, Xb, O, I, P, st, vE, vE, red, out=O2), 0, N * E * 10),
        "inorm_bwd3_add": (lambda: ops.inorm_bwd(3, Xb, X32, I, P, st, vE, vE, red, out=o32, add32=X32), 0, N * E * 20),
        "inorm_bwd3_mlp": (lambda: ops.inorm_bwd(3, X32, Xb, I, P, st, vE, vE, red, out=O2), 0, N * E * 14),
        "resid_bwd": (lambda: ops.resid_bwd(X32, Xb, O, I, P, rs, vE, torch.zeros(I, E, device=dev), torch.zeros(I, E, device=dev)), 0, N * E * 8),
        "colsum": (lambda: ops.colsum16(QKV, torch.zeros(3 * E, device=dev)), 0, N * 3 * E * 2),
    }
    if any(n.startswith("patch") for n in names) or not names:
        xf = torch.randn(I, 4, 512, 512, device=dev)
        a16 = torch.randn(I, 256, 256, 96, device=dev).half()
        ab16 = a16.bfloat16()
        Wp = torch.randn(96, 16, device=dev) / 4
        dWp = torch.zeros(96, 4, 2, 2, device=dev)
        st96 = torch.zeros(I, 96, 2, device=dev)
        pix = I * 256 * 256
        table.update({
            "patch_in": (lambda: ops.patch_in(xf, Wp.t().contiguous(), a16, st96), 0, pix * (64 + 192)),
            "patch_out": (lambda: ops.patch_out(a16, Wp, xf), 0, pix * (64 + 192)),
            "patch_wgrad": (lambda: ops.patch_wgrad(ab16, xf, dWp), 0, pix * (64 + 192)),
            "patch_wgrad_f16": (lambda: ops.patch_wgrad(a16, xf, dWp), 0, pix * (64 + 192)),
        })
    if not names:
        names = list(table)
    for n in names:
        fn, flops, nbytes = table[n]
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / iters * 1e3
        msg = f"[micro] {n:18s} {us:8.1f} us   {nbytes / us / 1e3:7.1f} GB/s (algorithmic bytes)"
        if flops:
            msg += f"   {flops / us / 1e6:7.1f} TFLOP/s"
        print(msg, flush=True)


if __name__ == "__main__":
    main()
