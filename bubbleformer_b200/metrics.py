"""Device-side rollout metrics: what upstream computes on predicted fields after `scripts/inference.py`.

  eikonal_loss      upstream bubbleformer/utils/losses.py:5-15   (SDF channel: mean (|grad phi| - 1)^2, dx = 1/32)
  heatflux          upstream bubbleformer/utils/heatflux.py:3-38 (wall heat flux of the FC-72 pool-boiling domain)
  rel_l2_per_field  LpLoss(d=2, p=2, reduce_dims=[0,1], reductions=["mean","mean"]) per field, inference.py:231
All reductions run in the library's kernels; only the final few scalars are combined with torch.
"""
from __future__ import annotations

from typing import Tuple

import torch

from . import _lib as L
from . import ops


def _check(t: torch.Tensor, nd: int, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError("bubbleformer_b200 runs on CUDA tensors only (no CPU fallback)")
    if t.dim() != nd:
        raise ValueError(f"{name}: expected {nd} dimensions, got {tuple(t.shape)}")
    return t.float().contiguous()


def eikonal_loss(phi: torch.Tensor, dx: float = 1.0 / 32) -> torch.Tensor:
    """phi: SDF (B, T, H, W).  Scalar tensor.  A rollout metric (upstream evaluates it under no_grad, inference.py):
    the kernel has no backward, so a gradient request fails loudly instead of returning a silently detached value."""
    if phi.requires_grad and torch.is_grad_enabled():
        raise NotImplementedError("bubbleformer_b200.eikonal_loss is a no-grad rollout metric (call it under "
                                  "torch.no_grad() or on a detached tensor); it has no backward kernel")
    phi = _check(phi, 4, "phi")
    B, T, H, W = phi.shape
    sums = torch.zeros(B * T, dtype=torch.float32, device=phi.device)
    L.check(L.lib.bf_eikonal_sums(phi.data_ptr(), sums.data_ptr(), B * T, H, W, float(dx),
                                  torch.cuda.current_stream().cuda_stream), "bf_eikonal_sums")
    return sums.sum() / (B * T * H * W)


def heatflux(dfun: torch.Tensor, temp: torch.Tensor, heater_temp: float, dx: float = 1.0 / 32, lc: float = 0.0007,
             x_min: float = -8.0) -> Tuple[torch.Tensor, torch.Tensor]:
    """dfun, temp: (T, H, W) un-normalised fields.  Returns (mean, max) over time of the wall heat flux."""
    dfun, temp = _check(dfun, 3, "dfun"), _check(temp, 3, "temp")
    if dfun.shape != temp.shape:
        raise ValueError("heatflux: dfun and temp must have the same shape")
    T, H, W = dfun.shape
    flux = torch.empty(T, dtype=torch.float32, device=dfun.device)
    L.check(L.lib.bf_heatflux_rows(dfun.data_ptr(), temp.data_ptr(), flux.data_ptr(), T, H * W, W, float(heater_temp),
                                   float(x_min), float(dx), float(lc), torch.cuda.current_stream().cuda_stream),
            "bf_heatflux_rows")
    return flux.mean(), flux.max()


def rel_l2_per_field(pred: torch.Tensor, tgt: torch.Tensor) -> torch.Tensor:
    """pred, tgt: (T, C, H, W) (one trajectory, as inference.py compares them).  Returns (C,): mean over time of
    ||pred - tgt||_2 / ||tgt||_2 per field; its mean is upstream's printed criterion value."""
    pred, tgt = _check(pred, 4, "pred"), _check(tgt, 4, "tgt")
    T, Cn = pred.shape[:2]
    sums = torch.zeros(T * Cn, 2, dtype=torch.float32, device=pred.device)
    ops.lploss_sums(pred.unsqueeze(0), tgt.unsqueeze(0), sums)
    return torch.sqrt(sums[:, 0] / sums[:, 1]).view(T, Cn).mean(dim=0)
