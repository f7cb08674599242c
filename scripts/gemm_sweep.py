"""Mainloop diagnostics of the tcgen05 GEMM: time per k iteration against ring depth, grid size and CTA pairing.

    python scripts/gemm_sweep.py            # runs every configuration in its own process (the switches are read once)
    python scripts/gemm_sweep.py one M N K  # one configuration, environment as given
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one(M, N, K, iters=40):
    import torch
    from bubbleformer_b200 import _lib as L, ops
    dev = "cuda"
    A = torch.randn(M, K, device=dev).bfloat16()
    B = (torch.randn(N, K, device=dev) * K ** -0.5).bfloat16()
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    for _ in range(5):
        ops.gemm(A, B, M, N, K, epilogue=L.EPI_STORE16, out16=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ops.gemm(A, B, M, N, K, epilogue=L.EPI_STORE16, out16=out)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    print(f"{us:.1f} us  {2.0 * M * N * K / us / 1e6:.0f} TFLOP/s")


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        return one(*(int(v) for v in sys.argv[2:5]))
    shapes = [(40960, 1152, 384), (40960, 384, 1536), (40960, 1536, 384)]
    for (M, N, K) in shapes:
        for cg in (1, 2):
            for stages in (2, 3, 4, 8):
                for grid in (148, 74, 36):
                    env = dict(os.environ, BF_GEMM_CG=str(cg), BF_GEMM_STAGES=str(stages), BF_GEMM_GRID=str(grid))
                    r = subprocess.run([sys.executable, __file__, "one", str(M), str(N), str(K)], env=env, capture_output=True, text=True)
                    print(f"M{M} N{N} K{K} cg{cg} stages<={stages} grid{grid}: {r.stdout.strip() or r.stderr.strip()[-200:]}", flush=True)


if __name__ == "__main__":
    main()
