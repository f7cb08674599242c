"""Autoregressive rollout driver (upstream scripts/inference.py:239-252 semantics) with a CUDA-graph step.

Upstream loops `pred = model(inp); inp = pred` one trajectory at a time, B = 1, without `no_grad`.  At B = 1 the
forward pass is ~700 short kernels (5 120 tokens), i.e. launch bound; the whole step is therefore captured once
in a CUDA graph (static input / output buffers, eval mode, no autograd) and replayed.  Trajectories are
independent, so multi-GPU rollout shards them over ranks with no collective (`shard_trajectories`).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch


class GraphedStep:
    """One captured forward step `y = model(x[, fluid_params])` for fixed shapes."""

    def __init__(self, model: torch.nn.Module, x: torch.Tensor, fluid_params: Optional[torch.Tensor] = None,
                 warmup: int = 3):
        if not x.is_cuda:
            raise RuntimeError("bubbleformer_b200 runs on CUDA tensors only (no CPU fallback)")
        self.model = model.eval()
        self.x = x.detach().clone()
        self.cond = fluid_params.detach().clone() if fluid_params is not None else None
        args = (self.x,) if self.cond is None else (self.x, self.cond)
        side = torch.cuda.Stream(device=x.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(max(warmup, 1)):          # first calls set kernel attributes / allocate: keep them out of the capture
                self.model(*args)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize(x.device)
        from . import _lib
        self.graph = torch.cuda.CUDAGraph()
        n0 = _lib.launch_count()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.y = self.model(*args)
        self.launches_per_step = _lib.launch_count() - n0      # library kernels recorded in the graph

    def __call__(self, x: torch.Tensor, fluid_params: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Replays the step; the returned tensor is the graph's static output buffer (clone it to keep it)."""
        if x.data_ptr() != self.x.data_ptr():
            self.x.copy_(x, non_blocking=True)
        if fluid_params is not None and self.cond is not None and fluid_params.data_ptr() != self.cond.data_ptr():
            self.cond.copy_(fluid_params, non_blocking=True)
        self.graph.replay()
        return self.y


def rollout(model: torch.nn.Module, x0: torch.Tensor, steps: int, fluid_params: Optional[torch.Tensor] = None,
            graphed: bool = True, keep: bool = True) -> Optional[torch.Tensor]:
    """Free-running rollout: x_{k+1} = model(x_k).  x0: (B, T, C, H, W) with C_in == C_out.

    Returns (steps, B, T, C, H, W) predictions when `keep`, else only advances (benchmarking)."""
    step = GraphedStep(model, x0, fluid_params) if graphed else None
    inp = x0
    outs: List[torch.Tensor] = []
    with torch.no_grad():
        for _ in range(steps):
            if step is not None:
                y = step(inp)
            else:
                y = model(inp) if fluid_params is None else model(inp, fluid_params)
            if keep:
                outs.append(y.clone())
                inp = outs[-1]
            else:
                inp = y                      # static output buffer -> copied into the static input on the next call
    return torch.stack(outs) if keep else None


def shard_trajectories(n_traj: int, rank: int, world: int) -> Sequence[int]:
    """Round-robin assignment of independent trajectories to ranks (no data-path collective)."""
    return range(rank, n_traj, world)


def gather_trajectories(local: torch.Tensor, n_traj: int, rank: int, world: int, group=None) -> torch.Tensor:
    """Reassemble per-trajectory results after a sharded rollout (the optional gather at the end; the rollout itself
    has no collective).  `local`: (this rank's trajectories in `shard_trajectories` order, ...).  Returns
    (n_traj, ...) in trajectory order on every rank.  Shards are padded to the largest one for the all-gather."""
    if world == 1:
        return local
    import torch.distributed as dist
    mine = len(shard_trajectories(n_traj, rank, world))
    if local.shape[0] != mine:
        raise ValueError(f"rank {rank} holds {local.shape[0]} trajectories, its shard has {mine}")
    per = -(-n_traj // world)
    pad = local.new_zeros((per,) + tuple(local.shape[1:]))
    pad[:mine] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    out = local.new_empty((n_traj,) + tuple(local.shape[1:]))
    for r in range(world):
        idx = list(shard_trajectories(n_traj, r, world))
        if idx:
            out[idx] = bufs[r][:len(idx)]
    return out


def evaluate_rollout(preds: torch.Tensor, targets: torch.Tensor, sdf_channel: int = 0) -> dict:
    """Metrics of one trajectory on the device: preds / targets (frames, C, H, W) as `scripts/inference.py:254-255`
    concatenates them.  Relative L2 per field (the criterion inference.py prints) and the eikonal residual of the
    predicted and the true signed-distance field (utils/losses.py:5-15)."""
    from . import metrics
    return {
        "rel_l2_per_field": metrics.rel_l2_per_field(preds, targets),
        "eikonal_pred": metrics.eikonal_loss(preds[:, sdf_channel].unsqueeze(0)),
        "eikonal_target": metrics.eikonal_loss(targets[:, sdf_channel].unsqueeze(0)),
    }
