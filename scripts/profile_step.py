"""Per-kernel CUDA-event breakdown of one config-2 training step (fwd + loss + bwd) on the B200 box.

    python scripts/profile_step.py [--batch 8] > gpurun_out/profile_step.txt
Events serialise nothing (same stream) but add ~2 us per launch; compare shares, and use bench.py for totals.
"""
import argparse
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--res", type=int, default=512)
    ap.add_argument("--detail", action="store_true")
    ap.add_argument("--big", action="store_true", help="film_avit_big (E=768, 12 heads): BASELINE configs[4] with --res 1024 --batch 1")
    args = ap.parse_args()
    import torch
    import bench
    from bubbleformer_b200 import get_model, ops
    from bubbleformer_b200.parallel import GradSink
    from oracle.param_init import fluid_params
    dev = "cuda"
    model = get_model("filmavit", time_window=5, **(bench.CFG_BIG if args.big else bench.CFG)).to(dev)
    model.train()
    sink = GradSink(model)
    B = args.batch
    x = torch.randn(B, 5, 4, args.res, args.res, device=dev)
    tgt = torch.randn_like(x)
    cond = fluid_params(B).to(dev)

    def step():
        sink.begin_step()
        y = model(x, cond)
        loss = bench.rel_l2_loss(y, tgt)
        loss.backward()
        return y

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    step()
    e1.record()
    torch.cuda.synchronize()
    print(f"uninstrumented step: {e0.elapsed_time(e1):.3f} ms")
    ops.PROFILE = []
    e0.record()
    sink.begin_step()
    y = model(x, cond)
    e1.record()
    loss = bench.rel_l2_loss(y, tgt)
    loss.backward()
    e2.record()
    torch.cuda.synchronize()
    recs, ops.PROFILE = ops.PROFILE, None
    print(f"instrumented: fwd {e0.elapsed_time(e1):.3f} ms, loss+bwd {e1.elapsed_time(e2):.3f} ms")
    by = collections.OrderedDict()
    tot = 0.0
    for name, tag, a, b in recs:
        ms = a.elapsed_time(b)
        tot += ms
        key = (name, tag)
        n, t = by.get(key, (0, 0.0))
        by[key] = (n + 1, t + ms)
    print(f"sum over {len(recs)} instrumented launches: {tot:.3f} ms")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for (name, tag), (n, t) in by.items():
        agg[name][0] += n
        agg[name][1] += t
    print("\n== by op ==")
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{name:18s} n={n:4d}  {t:8.3f} ms  {100 * t / tot:5.1f}%")
    print("\n== by op and shape ==")
    for (name, tag), (n, t) in sorted(by.items(), key=lambda kv: -kv[1][1]):
        extra = ""
        if name == "gemm":
            M, N, K = (int(s[1:]) for s in tag.split()[:3])
            extra = f"  {2.0 * M * N * K * n / (t * 1e-3) / 1e12:7.1f} TFLOP/s"
        print(f"{name:12s} {tag:52s} n={n:3d}  {t:8.3f} ms  {t / n * 1e3:8.1f} us/launch{extra}")


if __name__ == "__main__":
    main()
