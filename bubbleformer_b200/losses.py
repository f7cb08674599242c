"""Relative Lp loss of the reference (upstream bubbleformer/utils/losses.py:16-94) on the fused CUDA passes.

`LpLoss` keeps upstream's constructor (d, p, reduce_dims, reductions) and semantics: the ratio ||pred - tgt|| / ||tgt||
over the last `d` dimensions, then sum / mean over `reduce_dims` in order (keepdim, squeezed at the end).  The two
streaming passes over the fields (`bf_lploss_sums`, `bf_lploss_bwd`) implement d = 2, p = 2 -- both configurations
upstream uses: training `LpLoss(d=2, p=2, reduce_dims=[0,1,2], reductions=["mean","mean","sum"])` (modules.py:50) and
inference `LpLoss(d=2, p=2, reduce_dims=[0,1], reductions=["mean","mean"])` (scripts/inference.py:231).  Other (d, p)
raise NotImplementedError (there is deliberately no eager fallback).
"""
from __future__ import annotations

from typing import List, Union

import torch

from . import ops


class _SlabRatio(torch.autograd.Function):
    """ratio[...] = ||pred - tgt||_2 / ||tgt||_2 over the last two dimensions; gradient to `pred` only."""

    @staticmethod
    def forward(ctx, pred, tgt):
        if pred.dim() < 3 or pred.shape != tgt.shape:
            raise ValueError("LpLoss expects prediction and target of equal shape (..., H, W) with at least one leading dim")
        if not pred.is_cuda:
            raise RuntimeError("bubbleformer_b200 runs on CUDA tensors only (no CPU fallback)")
        pred = pred.float().contiguous()
        tgt = tgt.float().contiguous()
        lead = pred.shape[:-2]
        H, W = pred.shape[-2:]
        slabs = 1
        for s in lead:
            slabs *= s
        sums = torch.zeros(slabs, 2, dtype=torch.float32, device=pred.device)
        ops.lploss_sums(pred.view(1, 1, slabs, H, W), tgt.view(1, 1, slabs, H, W), sums)
        ctx.save_for_backward(pred, tgt, sums)
        return torch.sqrt(sums[:, 0] / sums[:, 1]).view(lead)

    @staticmethod
    def backward(ctx, g):
        pred, tgt, sums = ctx.saved_tensors
        # d ratio / d pred = (pred - tgt) / (||pred - tgt|| ||tgt||); a slab with pred == tgt has the sub-gradient 0 (not
        # rsqrt(0) * 0 = NaN), which is also what upstream's torch.norm backward returns at zero
        prod = sums[:, 0] * sums[:, 1]
        coef = torch.where(prod > 0, g.reshape(-1).float() * torch.rsqrt(prod), torch.zeros_like(prod)).contiguous()
        dpred = torch.empty_like(pred)
        H, W = pred.shape[-2:]
        ops.lploss_bwd(pred.view(1, 1, -1, H, W), tgt.view(1, 1, -1, H, W), coef, dpred.view(1, 1, -1, H, W))
        return dpred, None


class LpLoss(torch.nn.Module):
    """Drop-in for upstream's LpLoss (utils/losses.py:16-94) for d = 2, p = 2."""

    def __init__(self, d: int = 1, p: int = 2, reduce_dims: Union[int, List[int], None] = 0,
                 reductions: Union[str, List[str]] = "sum"):
        super().__init__()
        self.d, self.p = d, p
        self.reduce_dims = [reduce_dims] if isinstance(reduce_dims, int) else reduce_dims
        if self.reduce_dims is not None:
            if isinstance(reductions, str):
                assert reductions == "sum" or reductions == "mean"
                self.reductions = [reductions] * len(self.reduce_dims)
            else:
                for reduction in reductions:
                    assert reduction == "sum" or reduction == "mean"
                self.reductions = reductions

    def forward(self, y_pred: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        if self.d != 2 or self.p != 2:
            raise NotImplementedError("bubbleformer_b200.LpLoss implements d=2, p=2 (the configurations upstream trains "
                                      "and evaluates with); there is no eager fallback")
        diff = _SlabRatio.apply(y_pred, y)
        if self.reduce_dims is not None:
            for j, dim in enumerate(self.reduce_dims):
                diff = torch.sum(diff, dim=dim, keepdim=True) if self.reductions[j] == "sum" \
                    else torch.mean(diff, dim=dim, keepdim=True)
            diff = diff.squeeze()
        return diff


_TRAIN = LpLoss(d=2, p=2, reduce_dims=[0, 1, 2], reductions=["mean", "mean", "sum"])


def rel_l2_loss(pred: torch.Tensor, tgt: torch.Tensor) -> torch.Tensor:
    """Upstream's training criterion (modules.py:50) on (B, T, C, H, W); gradient flows to `pred` only."""
    if pred.dim() != 5:
        raise ValueError("rel_l2_loss expects (B, T, C, H, W) prediction and target of equal shape")
    return _TRAIN(pred, tgt)
