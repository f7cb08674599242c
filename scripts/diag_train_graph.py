"""Is the config-2 training step GPU-bound?  Times the eager step against the same step replayed from one CUDA
graph (fwd + rel-L2 loss + bwd captured together); the difference is host launch overhead the GPU waits on.

    python scripts/diag_train_graph.py [--batch 8] [--iters 10]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import bench
    from bubbleformer_b200 import get_model
    from bubbleformer_b200.parallel import GradSink
    from oracle.param_init import fluid_params

    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--iters", type=int, default=10)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.manual_seed(42)
    model = get_model("filmavit", time_window=bench.T, **bench.CFG).to(dev).train()
    g = torch.Generator(device="cpu").manual_seed(1234)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if "gamma" in n:
                p.copy_(0.05 * torch.randn(p.shape, generator=g))
    B = a.batch
    x = torch.randn(B, bench.T, bench.FIELDS, bench.RES, bench.RES, device=dev)
    tgt = torch.randn_like(x)
    cond = fluid_params(B).to(dev)
    sink = GradSink(model)

    def step():
        sink.begin_step()
        loss = bench.rel_l2_loss(model(x, cond), tgt)
        loss.backward()
        sink.finish()
        return loss

    def timeit(fn, n):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    for _ in range(3):
        step()
    print(f"eager step      {timeit(step, a.iters):8.3f} ms", flush=True)

    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        loss = step()
    graph.replay()
    torch.cuda.synchronize()
    l_graph = float(loss)
    print(f"graphed step    {timeit(graph.replay, a.iters):8.3f} ms   (loss {l_graph:.6f}, eager {float(step()):.6f})", flush=True)


if __name__ == "__main__":
    main()
