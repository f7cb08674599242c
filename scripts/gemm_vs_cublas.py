"""The block GEMMs of config 2 (with their fused epilogues) next to cuBLAS on the same shape (torch.matmul, bf16, plain
store).  Back-to-back launches with warm operands: compare the two columns with each other, not with the in-step numbers
(profiles/*_step_breakdown.txt), where the operands arrive cold.

    python scripts/gemm_vs_cublas.py > gpurun_out/<tag>_gemm_vs_cublas.txt
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from bubbleformer_b200 import _lib as L, engine, ops
    dev, bf = "cuda", torch.bfloat16
    N, E, I, P = 40960, 384, 40, 1024
    r = lambda *s: torch.randn(*s, device=dev)
    Xb, QKV, H = r(N, E).to(bf), r(N, 3 * E).to(bf), r(N, 4 * E).to(bf)
    X32 = r(N, E)
    Win, W1, W2, Wo = (r(3 * E, E) / 20).to(bf), (r(4 * E, E) / 20).to(bf), (r(E, 4 * E) / 40).to(bf), (r(E, E) / 20).to(bf)
    vE, v3E, v4E, rs = r(E), r(3 * E), r(4 * E), torch.ones(I, device=dev)
    st = torch.zeros(I, E, 2, device=dev)
    rstd = torch.rand(N, 6, 2, device=dev) + 0.5
    O, O2 = torch.empty(N, E, device=dev, dtype=bf), torch.empty(N, E, device=dev, dtype=bf)
    o32 = torch.empty(N, E, device=dev)
    out3, out4, out4b = torch.empty(N, 3 * E, device=dev, dtype=bf), torch.empty(N, 4 * E, device=dev, dtype=bf), torch.empty(N, 4 * E, device=dev, dtype=bf)
    g3, g4, g1 = torch.zeros(3 * E, E, device=dev), torch.zeros(4 * E, E, device=dev), torch.zeros(E, E, device=dev)
    g2 = torch.zeros(E, 4 * E, device=dev)
    sk = engine.pick_split
    cases = [
        ("QKV + per-head LayerNorm", N, 3 * E, E,
         lambda: ops.gemm(Xb, Win, N, 3 * E, E, epilogue=L.EPI_QKV_LN, bias=v3E, out16=out3, ln_head_dim=64, ln_rstd=rstd),
         lambda: torch.matmul(Xb, Win.t(), out=out3)),
        ("out-projection + residual + statistics", N, E, E,
         lambda: ops.gemm(Xb, Wo, N, E, E, epilogue=L.EPI_RESID, bias=vE, col_gamma=vE, row_scale=rs, rows_per_group=P, in32=X32,
                          out32=o32, out16b=O2, stats_out=st),
         lambda: torch.matmul(Xb, Wo.t(), out=O)),
        ("fc1 + GELU + GELU'", N, 4 * E, E,
         lambda: ops.gemm(Xb, W1, N, 4 * E, E, epilogue=L.EPI_GELU_D, bias=v4E, out16=out4, out16b=out4b),
         lambda: torch.matmul(Xb, W1.t(), out=out4)),
        ("fc2 + statistics", N, E, 4 * E,
         lambda: ops.gemm(H, W2, N, E, 4 * E, epilogue=L.EPI_STORE16, bias=vE, out16=O, rows_per_group=P, stats_out=st),
         lambda: torch.matmul(H, W2.t(), out=O)),
        ("fc2 dgrad x saved GELU' (+ column sums)", N, 4 * E, E,
         lambda: ops.gemm(Xb, W2, N, 4 * E, E, epilogue=L.EPI_DMUL, b_mode=L.B_KN, aux16=H, out16=out4, colsum_out=v4E),
         lambda: torch.matmul(Xb, W2, out=out4)),
        ("fc1 dgrad + fp32 accumulate", N, E, 4 * E,
         lambda: ops.gemm(H, W1, N, E, 4 * E, epilogue=L.EPI_ACC32, b_mode=L.B_KN, in32=X32, out32=o32),
         lambda: torch.matmul(H, W1, out=O)),
        ("QKV dgrad", N, E, 3 * E,
         lambda: ops.gemm(QKV, Win, N, E, 3 * E, epilogue=L.EPI_STORE16, b_mode=L.B_KN, out16=O),
         lambda: torch.matmul(QKV, Win, out=O)),
        ("out-projection dgrad", N, E, E,
         lambda: ops.gemm(Xb, Wo, N, E, E, epilogue=L.EPI_STORE16, b_mode=L.B_KN, out16=O),
         lambda: torch.matmul(Xb, Wo, out=O)),
        ("QKV wgrad (split-K, fp32 reduce-add)", 3 * E, E, N,
         lambda: ops.gemm(QKV, Xb, 3 * E, E, N, epilogue=L.EPI_ATOMIC32, a_mode=L.A_KM, b_mode=L.B_KN, split_k=sk(N, 3 * E, E), out32=g3),
         lambda: torch.matmul(QKV.t(), Xb)),
        ("fc1 wgrad", 4 * E, E, N,
         lambda: ops.gemm(H, Xb, 4 * E, E, N, epilogue=L.EPI_ATOMIC32, a_mode=L.A_KM, b_mode=L.B_KN, split_k=sk(N, 4 * E, E), out32=g4),
         lambda: torch.matmul(H.t(), Xb)),
        ("fc2 wgrad", E, 4 * E, N,
         lambda: ops.gemm(Xb, H, E, 4 * E, N, epilogue=L.EPI_ATOMIC32, a_mode=L.A_KM, b_mode=L.B_KN, split_k=sk(N, E, 4 * E), out32=g2),
         lambda: torch.matmul(Xb.t(), H)),
        ("out-projection wgrad", E, E, N,
         lambda: ops.gemm(Xb, Xb, E, E, N, epilogue=L.EPI_ATOMIC32, a_mode=L.A_KM, b_mode=L.B_KN, split_k=sk(N, E, E), out32=g1),
         lambda: torch.matmul(Xb.t(), Xb)),
    ]

    def timeit(fn, iters=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters * 1e3

    print(f"{'GEMM (config 2: 40960 tokens, E = 384)':44s} {'M':>6s} {'N':>6s} {'K':>6s} {'ours us':>8s} {'TFLOP/s':>8s} {'cuBLAS us':>9s} {'TFLOP/s':>8s}")
    for name, M, Nn, K, ours, ref in cases:
        fl = 2.0 * M * Nn * K
        a, b = timeit(ours), timeit(ref)
        print(f"{name:44s} {M:6d} {Nn:6d} {K:6d} {a:8.1f} {fl / a / 1e6:8.0f} {b:9.1f} {fl / b / 1e6:8.0f}", flush=True)
    print("ours: fused epilogue as named; cuBLAS: torch.matmul, plain bf16 store (wgrads: bf16 output, no split-K accumulate)")


if __name__ == "__main__":
    main()
