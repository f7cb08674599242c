"""Relative-L2 training loss of the reference (upstream bubbleformer/utils/losses.py:67-94, configured at
modules.py:50 as LpLoss(d=2, p=2, reduce_dims=[0,1,2], reductions=["mean","mean","sum"])) as two fused CUDA passes.

loss = sum_c mean_b mean_t ||pred - tgt||_2 / ||tgt||_2   (norms over the H x W pixels of each field)
"""
from __future__ import annotations

import torch

from . import ops


class _RelL2(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, tgt):
        if pred.dim() != 5 or pred.shape != tgt.shape:
            raise ValueError("rel_l2_loss expects (B, T, C, H, W) prediction and target of equal shape")
        if not pred.is_cuda:
            raise RuntimeError("bubbleformer_b200 runs on CUDA tensors only (no CPU fallback)")
        pred = pred.float().contiguous()
        tgt = tgt.float().contiguous()
        B, T, C = pred.shape[:3]
        sums = torch.zeros(B * T * C, 2, dtype=torch.float32, device=pred.device)
        ops.lploss_sums(pred, tgt, sums)
        ratio = torch.sqrt(sums[:, 0] / sums[:, 1])                  # (B*T*C,)
        ctx.save_for_backward(pred, tgt, sums)
        ctx.bt = B * T
        return ratio.sum() / (B * T)

    @staticmethod
    def backward(ctx, g):
        pred, tgt, sums = ctx.saved_tensors
        coef = (g / ctx.bt) * torch.rsqrt(sums[:, 0] * sums[:, 1])
        coef = coef.to(torch.float32).contiguous()
        dpred = torch.empty_like(pred)
        ops.lploss_bwd(pred, tgt, coef, dpred)
        return dpred, None


def rel_l2_loss(pred: torch.Tensor, tgt: torch.Tensor) -> torch.Tensor:
    """Scalar training loss; gradient flows to `pred` only (the target is data)."""
    return _RelL2.apply(pred, tgt)
