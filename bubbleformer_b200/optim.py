"""Optimiser step for the hot path's parameters: one fused kernel over the flat fp32 weight / gradient buffers.

Upstream builds `torch.optim.AdamW` / `Adam` or `lion_pytorch.Lion` over `model.parameters()` and a linear-warm-up +
cosine schedule (bubbleformer/modules.py:132-160, utils/lr_schedulers.py:4-31, config/optim_cfg/*.yaml).  Here every
parameter is a view of one flat buffer (autograd.WeightBank) and every gradient a view of another with the same layout
(parallel.GradSink), so a step is ONE launch of `bf_optim_step` over 28.9 M elements, which also rewrites the bf16
operand mirror the GEMMs read (the next forward then skips its cast pass).
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch

from . import _lib as L
from .autograd import WeightBank
from .parallel import GradSink

_KINDS = {"lion": 0, "adamw": 1, "adam": 2}
_DEFAULTS = {  # upstream config/optim_cfg/{lion,adamw,adam}.yaml; betas / eps are the libraries' defaults
    "lion": dict(lr=0.5e-4, weight_decay=1.0e-1, betas=(0.9, 0.99), eps=0.0),
    "adamw": dict(lr=2.5e-4, weight_decay=1.0e-2, betas=(0.9, 0.999), eps=1e-8),
    "adam": dict(lr=2.5e-4, weight_decay=1.0e-5, betas=(0.9, 0.999), eps=1e-8),
}


def cosine_warmup_lr(step: int, base_lr: float, warmup_iters: int, max_iters: int, eta_min: float = 0.0) -> float:
    """Learning rate of upstream's CosineWarmupLR (SequentialLR of LambdaLR(step / warmup_iters) and
    CosineAnnealingLR(T_max=max_iters, eta_min), milestone at warmup_iters) after `step` scheduler steps."""
    if step < warmup_iters:
        return base_lr * step / warmup_iters
    t = step - warmup_iters
    return eta_min + (base_lr - eta_min) * (1.0 + math.cos(math.pi * t / max_iters)) / 2.0


class FlatOptimizer:
    """`name` in {"lion", "adamw", "adam"}; hyper-parameters default to upstream's optim_cfg YAMLs.

    Needs a GradSink on the model (gradients in one flat buffer).  `step(lr=None)` applies one update to every
    parameter of the model; `lr` overrides the base learning rate for that step (schedulers stay on the host).
    """

    def __init__(self, model: torch.nn.Module, sink: GradSink, name: str = "lion", lr: Optional[float] = None,
                 weight_decay: Optional[float] = None, betas: Optional[Tuple[float, float]] = None,
                 eps: Optional[float] = None):
        if name not in _KINDS:
            raise ValueError(f"Optimizer {name} not supported")          # upstream modules.py:142
        d = _DEFAULTS[name]
        self.name, self.kind = name, _KINDS[name]
        self.lr = d["lr"] if lr is None else lr
        self.weight_decay = d["weight_decay"] if weight_decay is None else weight_decay
        self.betas = d["betas"] if betas is None else betas
        self.eps = d["eps"] if eps is None else eps
        bank = getattr(model, "_bank", None)
        if bank is None:
            bank = WeightBank(model)
            model._bank = bank
        bank.ensure()
        if bank.flat.numel() != sink.flat.numel():
            raise RuntimeError("FlatOptimizer: the weight bank and the gradient sink cover different parameter sets")
        self.bank, self.sink = bank, sink
        self.m = torch.zeros_like(bank.flat)
        self.v = torch.zeros_like(bank.flat) if name != "lion" else None
        self.steps = 0

    def step(self, lr: Optional[float] = None) -> None:
        self.steps += 1
        b = self.bank
        b.ensure()
        n = b.flat.numel()
        L.check(L.lib.bf_optim_step(self.kind, b.flat.data_ptr(), self.sink.flat.data_ptr(), self.m.data_ptr(),
                                    self.v.data_ptr() if self.v is not None else None, b.flat16.data_ptr(), n,
                                    float(self.lr if lr is None else lr), float(self.betas[0]), float(self.betas[1]),
                                    float(self.eps), float(self.weight_decay), self.steps,
                                    torch.cuda.current_stream().cuda_stream), "bf_optim_step")
        b.mark_mirror_fresh()

    def zero_grad(self, set_to_none: bool = False) -> None:
        self.sink.flat.zero_()
