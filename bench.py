"""Benchmark of the FiLMAViT hot path (BASELINE.json metric: fwd+bwd samples/s and rollout steps/s on one node of B200s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload train|rollout]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Headline workload "train" (BASELINE configs[1] / [2]): film_avit_small (E=384, 6 heads, 12 blocks, patch 16), train mode
(drop-path 0.2), per-GPU batch 8 of synthetic N(0,1) tensors (T=5, 4 fields, 512x512), relative-L2 loss (upstream
modules.py:50), forward + backward (+ gradient all-reduce over NCCL when N > 1, overlapped with backward).  One step =
one pass of the hot path over one batch; the optimizer is not part of the path.

One JSON line is printed by rank 0:
  value        whole-job samples/s with inputs resident in HBM (CUDA events, max over ranks)
  e2e          same metric through the public module API with pinned HOST inputs: H2D of inputs + D2H of the loss inside
               the timed region
  roofline     the dominant kernel (tcgen05 GEMM): algorithmic FLOPs of its launches in one step / their summed
               CUDA-event durations, against the measured bf16 peak (MEASURED_PEAKS.json, sustained figure); step_frac =
               whole-step algorithmic FLOP/s over that peak; hbm = achieved GB/s per HBM-bound kernel family from the
               committed ncu capture
  parity       the very model that was timed, B=1, against the CPU oracle run on the host (fwd rel-L2, gradients)
  rollout      BASELINE configs[3]: 200-step autoregressive rollouts of >= 8 trajectories sharded over the ranks
               (512x512 and a 128x1024 strip), steps/s aggregate and per GPU
  config5      BASELINE configs[4]: film_avit_big (E=768) at 1024x1024, per-GPU batch 1, fwd+bwd samples/s
  cpu_baseline the CPU oracle port (oracle/filmavit_oracle.py, a restatement of the pure-Python reference) timed on the
               box's host cores on a bounded sample (B=1 micro-batches of the same workload)
`--impl reference` times that CPU port alone (rank 0 only), one B=1 micro-batch per step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(input_fields=4, output_fields=4, patch_size=16, embed_dim=384, num_heads=6, processor_blocks=12,
           drop_path=0.2, attn_scale=True, feat_scale=True, num_fluid_params=9)          # film_avit_small.yaml
CFG_BIG = dict(CFG, embed_dim=768, num_heads=12)                                          # film_avit_big.yaml
T, FIELDS, RES, BATCH = 5, 4, 512, 8
FLOPS_FWD_BWD_PER_SAMPLE = 948.66e9          # BASELINE.md section 3 (matmul/conv/bmm FLOPs of the reference graph)
FLOPS_FWD_PER_SAMPLE = 316.55e9
FLOPS_BIG_1024_FWD_BWD = 15126.93e9          # film_avit_big at 1024x1024
METRIC = "filmavit_fwd_bwd_samples_per_sec"
ROLLOUT_STEPS, ROLLOUT_TRAJ = 200, 8


def flops_fwd(cfg, t, h_px, w_px):
    """Forward matmul/conv FLOPs of one sample (SURVEY.md 8d formula; 316.55 GFLOP for film_avit_small at 512x512)."""
    E, L, p, C = cfg["embed_dim"], cfg["processor_blocks"], cfg["patch_size"], cfg["input_fields"]
    h, w = h_px // p, w_px // p
    P = h * w
    HW = h_px * w_px
    blocks = L * (32 * t * P * E * E + 4 * t * P * E * (t + h + w))
    embed = 2 * ((HW // 4) * 4 * C * (E // 4) + (HW // 16) * E * (E // 4) + (HW // 64) * E * (E // 4) + (HW // 256) * E * E)
    return blocks + t * 2 * embed


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(bf16=p["bf16_tflops_sustained"], bf16_burst=p["bf16_tflops"], hbm=p["hbm_gbs"], source="measured")
    return dict(bf16=1400.0, bf16_burst=1590.0, hbm=6650.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "200"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.25)
            self.proc.terminate()

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=mx or None, reasons=sorted(reasons),
                    samples=len(sm))


def rel_l2_loss(pred, tgt):
    """LpLoss(d=2, p=2, reduce_dims=[0,1,2], reductions=[mean,mean,sum]) -- upstream utils/losses.py:67-94,
    through the fused CUDA loss of this repo (bubbleformer_b200.losses)."""
    from bubbleformer_b200.losses import rel_l2_loss as fused
    return fused(pred, tgt)


# ---------------------------------------------------------------------------------------------------------------
# CPU legs (the oracle port on the host cores)
# ---------------------------------------------------------------------------------------------------------------
def cpu_port_step(sd, x, tgt, cond, train: bool):
    """One B=1 micro-batch of the workload through the CPU oracle port (fwd + loss + bwd)."""
    import torch
    from oracle import filmavit_oracle as O
    if not train:
        with torch.no_grad():
            return O.forward(sd, x, cond, patch_size=CFG["patch_size"], num_heads=CFG["num_heads"])
    y = O.forward(sd, x, cond, patch_size=CFG["patch_size"], num_heads=CFG["num_heads"])
    loss = O.rel_l2_loss(y, tgt)
    grads = torch.autograd.grad(loss, list(sd.values()))
    return grads


def time_cpu_port(steps: int, warmup: int, train: bool = True, res: int = RES):
    import torch
    from oracle.param_init import fluid_params, param_shapes, random_state_dict
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    shapes = param_shapes(**{k: v for k, v in CFG.items() if k != "drop_path"})
    sd = {k: v.requires_grad_(train) for k, v in random_state_dict(shapes, seed=42).items()}
    g = torch.Generator().manual_seed(42)
    x = torch.randn(1, T, FIELDS, res, res, generator=g)
    tgt = torch.randn(1, T, FIELDS, res, res, generator=g)
    cond = fluid_params(1)
    for _ in range(warmup):
        cpu_port_step(sd, x, tgt, cond, train)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_port_step(sd, x, tgt, cond, train)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return dict(ms_per_step=dt * 1e3, samples_per_s=1.0 / dt, cores=cores)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)     # torchrun exports 1; this arm runs on rank 0 alone
    steps = max(1, min(args.steps, 12))      # bounded sample: a B=1 micro-batch takes ~2 s on 16 host cores
    warm = max(1, min(args.warmup, 2))
    r = time_cpu_port(steps, warm, train=args.workload == "train")
    unit = "samples/s" if args.workload == "train" else "steps/s"
    line = {
        "impl": "reference", "metric": METRIC if args.workload == "train" else "filmavit_rollout_steps_per_sec",
        "value": r["samples_per_s"], "unit": unit, "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "film_avit_small %s, T=5, 4 fields, 512x512, one B=1 micro-batch per step on the host CPU"
                               % ("fwd+bwd" if args.workload == "train" else "fwd"),
                   "steps_requested": args.steps, "warmup_requested": args.warmup,
                   "steps_note": f"--steps / --warmup are clamped to {steps} / {warm} on this arm: every step is a bounded "
                                 f"sample (one B=1 micro-batch, seconds of host time), so the run ends within minutes"},
        "cpu_baseline": {"value": r["samples_per_s"], "unit": unit, "cores": r["cores"], "kind": "port",
                         "sample": f"{steps} B=1 micro-batches of the same workload (oracle/filmavit_oracle.py, torch CPU, "
                                   f"{r['cores']} threads); the reference is pure Python/PyTorch so the port runs the same ATen ops"},
        "e2e": {"value": r["samples_per_s"], "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------
# helpers of the GPU arm
# ---------------------------------------------------------------------------------------------------------------
def build_model(cfg, dev, seed=42):
    """Same weights on every rank (what DistributedDataParallel's construction broadcast gives upstream): the model is
    built from one seed; the layer-scale 1e-6 init is broken so that every block carries signal."""
    import torch
    from bubbleformer_b200 import get_model
    torch.manual_seed(seed)
    model = get_model("filmavit", time_window=T, **cfg).to(dev)
    g = torch.Generator(device="cpu").manual_seed(1234)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if "gamma" in n:
                p.copy_(0.05 * torch.randn(p.shape, generator=g))
            elif "freq_scalar" in n:
                p.copy_(0.2 * torch.randn(p.shape, generator=g))
    return model


def max_over_ranks(v: float, dev, world: int) -> float:
    import torch
    import torch.distributed as dist
    if world == 1:
        return v
    t = torch.tensor([v], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def timed(fn, n: int, dev, world: int) -> float:
    """ms per call of fn over n calls: barrier + synchronize on both sides, CUDA events, max over ranks."""
    import torch
    import torch.distributed as dist

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    barrier()
    return max_over_ranks(e0.elapsed_time(e1), dev, world) / n


def parity_check(model, dev):
    """B=1 check of the very model the bench times (same weights), train mode with injected stochastic-depth masks,
    against the CPU oracle on the host: forward rel-L2 per channel, dx, every parameter gradient."""
    import torch
    from oracle import parity
    sd = {k: v.detach().float().cpu().clone() for k, v in model.state_dict().items()}
    g = torch.Generator().manual_seed(777)
    x = torch.randn(1, T, FIELDS, RES, RES, generator=g)
    tgt = torch.randn(1, T, FIELDS, RES, RES, generator=g)
    from oracle.param_init import fluid_params
    cond = fluid_params(1)
    masks = parity.draw_masks(CFG["drop_path"], CFG["processor_blocks"], 1, T, seed=778)
    ref = parity.oracle_run(sd, x, tgt, cond, CFG, masks)
    got = parity.candidate_run(model, x, tgt, cond, masks, loss_fn=rel_l2_loss)
    res = parity.compare(ref, got)
    return {"fwd_rel_l2": res["fwd_rel_l2"], "fwd_rel_l2_per_channel": res["fwd_rel_l2_per_channel"],
            "grad_rel": res["grad_rel"], "dx_rel_l2": res["dx_rel_l2"], "loss_rel": res["loss_rel"],
            "tolerance": {"fwd_rel_l2": 1e-2, "grad_rel": 2e-2}, "pass": parity.passes(res),
            "what": "the timed film_avit_small weights, B=1, T=5, 4 fields, 512x512, train mode with injected drop-path masks, "
                    "fwd + rel-L2 loss + bwd on cuda vs oracle/filmavit_oracle.py in fp32 on the host "
                    f"({ref['seconds']:.1f} s); grad_rel = global-norm-relative error over all parameter gradients"}


def rollout_section(dev, rank, world, steps=ROLLOUT_STEPS, n_traj=ROLLOUT_TRAJ):
    """BASELINE configs[3]: `steps`-step autoregressive rollouts (upstream scripts/inference.py:239-252: pred = model(inp),
    inp = pred) of n_traj independent trajectories, sharded round-robin over the ranks with no data-path collective.  A
    rank advances its trajectories in lockstep as one batch through ONE captured CUDA graph of the forward step."""
    import torch
    from bubbleformer_b200.rollout import GraphedStep, shard_trajectories
    from oracle.param_init import fluid_params
    model = build_model(dict(CFG, drop_path=0.0), dev).eval()
    n_traj = max(n_traj, world)
    mine = list(shard_trajectories(n_traj, rank, world))
    out = {"steps": steps, "trajectories": n_traj, "sharding": "round-robin by trajectory, no collective",
           "model": "film_avit_small, eval, no_grad, CUDA graph per (domain, local batch)", "domains": {}}
    pk = peaks()
    for name, (H, W) in (("512x512", (512, 512)), ("128x1024", (128, 1024))):
        torch.manual_seed(100 + rank)
        x0 = torch.randn(len(mine), T, FIELDS, H, W, device=dev)
        cond = fluid_params(n_traj)[mine].to(dev)
        step = GraphedStep(model, x0, cond)
        state = {"x": x0}

        def advance():
            state["x"] = step(state["x"])          # static output buffer -> copied into the static input next call

        for _ in range(3):
            advance()
        ms = timed(advance, steps, dev, world)
        agg = n_traj / (ms * 1e-3)
        fl = flops_fwd(CFG, T, H, W)
        # latency of ONE trajectory (B = 1), the unit upstream's script runs
        if len(mine) > 1:
            step1 = GraphedStep(model, x0[:1].clone(), cond[:1].clone())
            s1 = {"x": x0[:1].clone()}

            def adv1():
                s1["x"] = step1(s1["x"])
            for _ in range(3):
                adv1()
            ms1 = timed(adv1, min(steps, 50), dev, world)
            del step1
        else:
            ms1 = ms
        finite = bool(torch.isfinite(state["x"]).all())
        out["domains"][name] = {
            "steps_per_s": agg, "steps_per_s_per_gpu": agg / world, "ms_per_step": ms, "traj_per_gpu": len(mine),
            "b1_steps_per_s": 1e3 / ms1, "b1_ms_per_step": ms1, "launches_per_step": step.launches_per_step,
            "frac_of_bf16_peak": agg / world * fl / 1e12 / pk["bf16"], "fwd_gflop_per_step": fl / 1e9, "finite": finite}
        del step
        torch.cuda.empty_cache()
    del model
    torch.cuda.empty_cache()
    return out


def config5_section(dev, rank, world, steps=5, warmup=2):
    """BASELINE configs[4]: film_avit_big (E=768, 12 heads, 12 blocks) at 1024x1024, bf16 fwd + rel-L2 loss + bwd,
    per-GPU batch 1 (global batch = N), gradient all-reduce (461 MB fp32) overlapped with backward, CUDA-graph step."""
    import torch
    from bubbleformer_b200.parallel import GradSink, GraphedTrainStep
    from oracle.param_init import fluid_params
    model = build_model(CFG_BIG, dev).train()
    sink = GradSink(model)
    torch.manual_seed(500 + rank)
    x = torch.randn(1, T, FIELDS, 1024, 1024, device=dev)
    tgt = torch.randn(1, T, FIELDS, 1024, 1024, device=dev)
    cond = fluid_params(1).to(dev)
    try:
        gstep = GraphedTrainStep(model, rel_l2_loss, sink, x, tgt, cond, warmup=2)
        for _ in range(warmup):
            gstep(x, tgt, cond)
        ms = timed(lambda: gstep(x, tgt, cond), steps, dev, world)
        loss = float(gstep.loss)
        val = world / (ms * 1e-3)
        pk = peaks()
        out = {"model": "film_avit_big (E=768, 12 heads, 12 blocks)", "resolution": "1024x1024", "per_gpu_batch": 1,
               "global_batch": world, "samples_per_s": val, "ms_per_step": ms, "steps": steps, "cuda_graph": True,
               "launches_per_step": gstep.launches_per_step,
               "algorithmic_tflops_per_gpu": val / world * FLOPS_BIG_1024_FWD_BWD / 1e12,
               "frac_of_bf16_peak": val / world * FLOPS_BIG_1024_FWD_BWD / 1e12 / pk["bf16"],
               "peak_mem_gib": torch.cuda.max_memory_allocated(dev) / 2**30, "loss": loss,
               "finite": loss == loss and abs(loss) != float("inf")}
        gstep.graph.reset()
        del gstep
    finally:
        sink.close()
    del model, sink
    torch.cuda.empty_cache()
    return out


def hbm_families():
    """Achieved HBM GB/s per kernel family of one replayed config-2 step, from the committed ncu capture
    (profiles/hbm_families.json: dram__bytes_read.sum + dram__bytes_write.sum and gpu__time_duration.sum per launch)."""
    path = os.path.join(ROOT, "profiles", "hbm_families.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        return json.load(f)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="train", choices=["train", "rollout"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the parity / rollout / config5 sections of the line")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of the captured training step")
    ap.add_argument("--profile", action="store_true",
                    help="ncu captures only (scripts/ncu_profiles.sh): one eager warm-up step, then --steps replays of the "
                         "captured step and exit; prints no bench line")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    # torchrun exports OMP_NUM_THREADS=1; the host-side legs (parity oracle, CPU baseline) want the box's cores
    _world = int(os.environ.get("WORLD_SIZE", "1"))
    os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // _world))
    import torch
    import torch.distributed as dist
    from bubbleformer_b200 import _lib, ops
    from bubbleformer_b200.parallel import GradSink
    from oracle.param_init import fluid_params

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # the gradient all-reduce moves 115 MB per ~28 ms step (4 GB/s): a few NCCL CTAs are plenty, and every SM they
        # do not occupy stays with the persistent GEMM grids they overlap with
        os.environ.setdefault("NCCL_MAX_CTAS", "4")
        dist.init_process_group("nccl", device_id=dev)
    W = max(args.warmup, 3)
    K = max(args.steps, 1)
    B = args.batch if args.workload == "train" else 1

    model = build_model(CFG, dev)                 # identical weights on every rank
    torch.manual_seed(4242 + rank)                # per-rank data and drop-path masks
    x = torch.randn(B, T, FIELDS, RES, RES, device=dev)
    tgt = torch.randn(B, T, FIELDS, RES, RES, device=dev)
    cond = fluid_params(B).to(dev)
    train = args.workload == "train"
    model.train() if train else model.eval()
    sink = GradSink(model) if train else None

    gstep = None
    if not train:
        from bubbleformer_b200.rollout import GraphedStep
        gstep = GraphedStep(model, x, cond)       # the whole B=1 step replayed from one CUDA graph

    gtrain = None
    if train and not args.no_graph:
        from bubbleformer_b200.parallel import GraphedTrainStep
        gtrain = GraphedTrainStep(model, rel_l2_loss, sink, x, tgt, cond,    # fwd + loss + bwd (+ all-reduce) in one graph
                                  warmup=1 if args.profile else 3)

    def eager_step(xd, td, cd):
        sink.begin_step()
        y = model(xd, cd)
        loss = rel_l2_loss(y, td)
        loss.backward()
        sink.finish()
        return loss

    if args.profile:
        for _ in range(K):
            step_fn = gtrain if gtrain is not None else eager_step
            step_fn(x, tgt, cond)
        torch.cuda.synchronize()
        return

    def step(xd, td, cd):
        if gtrain is not None:
            return gtrain(xd, td, cd)
        if train:
            return eager_step(xd, td, cd)
        return gstep(xd, cd)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(W):
        step(x, tgt, cond)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = _lib.launch_count()
    with ClockSampler(local) as clocks:
        barrier()
        e0.record()
        for _ in range(K):
            out = step(x, tgt, cond)
        e1.record()
        barrier()
    launches = _lib.launch_count() - n0
    graphed = gtrain if train else gstep
    if graphed is not None:                 # replays do not pass through the C ABI: count what the graph recorded
        launches = graphed.launches_per_step * K
    ms = max_over_ranks(e0.elapsed_time(e1), dev, world)
    ms_per_step = ms / K
    value = world * B / (ms_per_step * 1e-3)

    # ---- end to end through the public API with host buffers ----
    # Every step copies its own inputs from pinned host memory (H2D inside the timed region) and one step result
    # is read back to the host per step.  Like a data loader with one batch of prefetch, the copy of step i+1 is
    # issued on a copy stream while step i computes, and the loss read lags one step so it never stalls the GPU.
    xh, th, ch = (t.cpu().pin_memory() for t in (x, tgt, cond))
    Ke = max(3, min(K, 20))
    copy_stream = torch.cuda.Stream(device=dev)

    def fetch():
        with torch.cuda.stream(copy_stream):
            bufs = tuple(t.to(dev, non_blocking=True) for t in (xh, th, ch))
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return bufs, ev

    loss_host = torch.empty(2, dtype=torch.float32).pin_memory()          # two slots: the read lags one step
    loss_ev = [torch.cuda.Event(), torch.cuda.Event()]

    def run_e2e(n):
        nxt = fetch()
        host = 0.0
        for i in range(n):
            (xd, td, cd), ev = nxt
            nxt = fetch()                                    # prefetch the next step's inputs
            torch.cuda.current_stream().wait_event(ev)
            for t_ in (xd, td, cd):
                t_.record_stream(torch.cuda.current_stream())
            r = step(xd, td, cd)
            r = r.detach() if train else r[0, 0, 0, 0, :1]
            # D2H read of this step's result: an asynchronous copy into pinned memory behind the step (the step's result
            # is a static buffer of the captured graph, so it is copied out before the next replay can overwrite it) ...
            slot = i & 1
            loss_host[slot:slot + 1].copy_(r.reshape(1), non_blocking=True)
            loss_ev[slot].record()
            if i > 0:                                        # ... consumed on the host one step later
                loss_ev[slot ^ 1].synchronize()
                host = float(loss_host[slot ^ 1])
        loss_ev[(n - 1) & 1].synchronize()
        host = float(loss_host[(n - 1) & 1])
        torch.cuda.synchronize()
        return host

    run_e2e(2)
    barrier()
    t0 = time.perf_counter()
    run_e2e(Ke)
    e2e_s = max_over_ranks(time.perf_counter() - t0, dev, world)
    e2e_value = world * B * Ke / e2e_s
    h2d = (xh.numel() + th.numel() + ch.numel()) * 4 * (Ke + 1) // Ke      # one extra prefetch is issued per run
    d2h = 4
    del xh, th

    # ---- roofline of the dominant kernel: one instrumented step, CUDA events around every GEMM launch ----
    roof = None
    if rank == 0:
        ops.GEMM_TIMING = []
        ops.GEMM_RECORD = []
    # The eager step is host-bound (that is why the timed region replays a graph), so the GPU is first parked on a
    # ~60 ms spin: the whole step is enqueued behind it and every event pair brackets device time only.
    torch.cuda._sleep(int(0.06 * 1.9e9))
    (eager_step if train else step)(x, tgt, cond)     # every rank steps (the gradient all-reduce is collective)
    torch.cuda.synchronize()
    flops_step = FLOPS_FWD_BWD_PER_SAMPLE if train else FLOPS_FWD_PER_SAMPLE
    if rank == 0:
        recs, ops.GEMM_TIMING = ops.GEMM_TIMING, None
        launches_rec, ops.GEMM_RECORD = ops.GEMM_RECORD, None
        bracket_ms = sum(a.elapsed_time(b) for a, b, _ in recs)
        gemm_flops = sum(f for _, _, f in recs)
        pk = peaks()
        bracket = gemm_flops / (bracket_ms * 1e-3) / 1e12 if bracket_ms > 0 else 0.0
        # (a) above: one CUDA-event pair around every launch of the eager step -- each bracket also holds the launch
        #     latency of its kernel (events serialise the stream and switch programmatic dependent launch off).
        # (b) the very same launches (same argument structs, same buffers, same order) captured into one CUDA graph and
        #     replayed back to back, PDL-chained as in the timed step, between ONE event pair: the kernels' device time
        #     without the per-launch latency.  Operands are as cold as in the step (a step's GEMMs touch > 30 GB).
        gemm_ms, method = bracket_ms, "cuda events around every launch of one eager step"
        try:
            ops.gemm_replay(launches_rec)
            torch.cuda.synchronize()
            gg = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gg):
                ops.gemm_replay(launches_rec)
            gg.replay()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 3
            g0.record()
            for _ in range(reps):
                gg.replay()
            g1.record()
            torch.cuda.synchronize()
            gemm_ms = g0.elapsed_time(g1) / reps
            method = ("the step's %d bf_gemm launches (same arguments, buffers and order) replayed back to back from one "
                      "CUDA graph, one cuda event pair around %d replays" % (len(launches_rec), reps))
            del gg
        except Exception as e:           # keep the bracket numbers if the replay cannot be built
            method += " (graph replay failed: %r)" % (e,)
        del launches_rec
        ach = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
        traffic, traffic_src = None, None
        tf = os.path.join(ROOT, "profiles", "gemm_traffic.json")
        if train and os.path.exists(tf):             # DRAM bytes per GEMM launch from the committed ncu capture
            with open(tf) as f:
                tj = json.load(f)
            traffic, traffic_src = tj.get("dram_bytes_per_launch"), tj.get("source")
        step_tflops = value / world * flops_step / 1e12
        roof = {"bound": "tensor", "kernel": "gemm_tcgen05_kernel", "achieved": ach, "peak": pk["bf16"],
                "unit": "TFLOP/s", "frac": ach / pk["bf16"], "step_frac": step_tflops / pk["bf16"],
                "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": pk["source"] + " (sustained)",
                "launches_per_step": len(recs), "gemm_ms_per_step": gemm_ms, "method": method,
                "event_bracket": {"gemm_ms_per_step": bracket_ms, "achieved": bracket, "frac": bracket / pk["bf16"]},
                "gemm_share_of_step": gemm_ms / ms_per_step,
                "step_algorithmic_tflops": step_tflops, "step_frac_of_peak": step_tflops / pk["bf16"],
                "hbm": hbm_families()}

    # ---- extras: parity of the timed model, rollout (configs[3]) and film_avit_big at 1024x1024 (configs[4]) ----
    parity = rollout = config5 = None
    extras = train and not args.no_extras
    if extras:
        sink.close()                       # the parity backward must not enter the (collective) gradient sink
        if rank == 0:
            try:
                parity = parity_check(model, dev)
            except Exception as e:        # never lose the bench line to the checker
                parity = {"error": repr(e)}
    if gtrain is not None and world > 1:
        gtrain.graph.reset()
    del gtrain, gstep, graphed
    if extras:
        del model, sink, x, tgt
        torch.cuda.empty_cache()
        barrier()
        try:
            rollout = rollout_section(dev, rank, world)
        except Exception as e:
            rollout = {"error": repr(e)}
        barrier()
        try:
            config5 = config5_section(dev, rank, world)
        except Exception as e:
            config5 = {"error": repr(e)}
        barrier()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = time_cpu_port(3, 1, train=train)
        unit = "samples/s" if train else "steps/s"
        cpu = {"value": r["samples_per_s"], "unit": unit, "cores": r["cores"], "kind": "port",
               "sample": f"3 B=1 micro-batches of the same workload after 1 warm-up (oracle port, torch CPU, {r['cores']} threads)"}

    if rank == 0:
        line = {
            "metric": METRIC if train else "filmavit_rollout_steps_per_sec",
            "value": value, "unit": "samples/s" if train else "steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": ("film_avit_small fwd+bwd (train mode, drop_path 0.2, rel-L2 loss), per-GPU batch %d, "
                                    "T=5, 4 fields, 512x512" % B) if train else
                                   "film_avit_small autoregressive rollout step (eval, no_grad), B=1, T=5, 4 fields, 512x512",
                       "parallelism": f"dp{world}", "cuda_graph": (not args.no_graph) if train else True,
                       "l2_policy": "per-step working set (>10 GB of activations) far exceeds the 126 MB L2",
                       "precision": "bf16 block GEMMs / attention, fp16 patch embed+unembed forward, fp32 residual stream, statistics and gradients",
                       "weights": "identical on every rank (one seed); per-rank data and drop-path masks",
                       "comm": ("NCCL all-reduce (AVG) of the flat fp32 gradient buffer in 8 MB buckets on a side stream, "
                                "NCCL_MAX_CTAS=%s, %s SMs reserved for it (kernel grids sized for the rest)"
                                % (os.environ.get("NCCL_MAX_CTAS", "default"),
                                   os.environ.get("BF_RESERVED_SMS", os.environ.get("NCCL_MAX_CTAS", "0")))) if world > 1 else "none"},
            "e2e": {"value": e2e_value, "unit": "samples/s" if train else "steps/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps": Ke},
            "gpu_launches": launches, "gpu_launches_per_step": launches / K,
            "clocks": clocks.summary(), "roofline": roof, "parity": parity, "rollout": rollout, "config5": config5,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        # Tear-down: never let a stuck communicator destructor keep the job alive after the result line is out.
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        done = threading.Event()

        def _destroy():
            try:
                dist.destroy_process_group()
            finally:
                done.set()
        threading.Thread(target=_destroy, daemon=True).start()
        if not done.wait(20.0):
            os._exit(0)


if __name__ == "__main__":
    main()
