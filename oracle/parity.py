"""Whole-model parity harness: the CUDA path against the CPU oracle run on the host -- TEST INFRASTRUCTURE ONLY.

Used by `tests/` (through scripts/gpu_diag_model.py), `__graft_entry__.smoke()` and the `parity` leg of `bench.py`,
always as the checker.  The oracle (oracle/filmavit_oracle.py, pinned against the live reference by
oracle/make_golden.py) is evaluated in fp32 on the host cores with the same weights, inputs and stochastic-depth
masks as the candidate; compared are the forward fields per channel, the loss, the input gradient and every
parameter gradient (global-norm-relative, the measure SURVEY.md 7.3 prescribes: `knorm.bias` and `mlp.fc2.bias` have
identically-zero true gradients).

Reference behaviour being checked: bubbleformer/models/axial_vit.py:217-242 (FiLMConditionedAViT.forward) and its
autograd backward, bubbleformer/modules.py:300-312 (training_step: rel-L2 loss of utils/losses.py:67-94).
"""
from __future__ import annotations

import time
from typing import Dict, List, Optional

import numpy as np
import torch

from . import filmavit_oracle as O
from .param_init import fluid_params, param_shapes, random_state_dict


def draw_masks(drop_path: float, blocks: int, B: int, T: int, seed: int) -> Optional[List[tuple]]:
    """Stochastic-depth factors per block (timm DropPath: bernoulli(keep) / keep over dim 0; rates
    np.linspace(0, drop_path, blocks) as upstream axial_vit.py:199-204), drawn from a numpy stream."""
    if not drop_path:
        return None
    g = np.random.RandomState(seed)
    out = []
    for p in np.linspace(0, drop_path, blocks):
        if p == 0.0:
            out.append((None, None, None))
            continue
        keep = 1.0 - p
        out.append(tuple(torch.from_numpy((g.uniform(size=n) < keep).astype(np.float32) / np.float32(keep))
                         for n in (B, B * T, B * T)))
    return out


def make_case(cfg: Dict, B: int, T: int, H: int, W: int, seed: int, train: bool = True):
    """Seeded weights (randomised layer scales etc.), inputs, targets, fluid vectors and masks."""
    shapes = param_shapes(**{k: v for k, v in cfg.items() if k != "drop_path"})
    sd = random_state_dict(shapes, seed=seed)
    g = torch.Generator().manual_seed(seed + 1000)
    x = torch.randn(B, T, cfg["input_fields"], H, W, generator=g)
    tgt = torch.randn(B, T, cfg["output_fields"], H, W, generator=g)
    cond = fluid_params(B) if cfg.get("num_fluid_params") is not None else None
    masks = draw_masks(cfg.get("drop_path", 0.0), cfg["processor_blocks"], B, T, seed + 2000) if train else None
    return dict(sd=sd, x=x, tgt=tgt, cond=cond, masks=masks)


def oracle_run(sd, x, tgt, cond, cfg, masks, grads: bool = True, threads: Optional[int] = None) -> Dict:
    """Forward (+ loss + every gradient) of the CPU oracle in fp32.  Returns tensors on the host."""
    import os
    torch.set_num_threads(threads or os.cpu_count() or 1)
    t0 = time.perf_counter()
    fw = dict(patch_size=cfg["patch_size"], num_heads=cfg["num_heads"], attn_scale=cfg.get("attn_scale", True),
              feat_scale=cfg.get("feat_scale", True))
    if not grads:
        with torch.no_grad():
            y = O.forward(sd, x, cond, drop_masks=masks, **fw)
        return dict(y=y, seconds=time.perf_counter() - t0)
    sdg = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    xg = x.detach().clone().requires_grad_(True)
    y = O.forward(sdg, xg, cond, drop_masks=masks, **fw)
    loss = O.rel_l2_loss(y, tgt)
    loss.backward()
    out = dict(y=y.detach(), loss=float(loss), dx=xg.grad, grads={k: v.grad for k, v in sdg.items()},
               seconds=time.perf_counter() - t0)
    del sdg, xg, loss
    return out


def _rel(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


def candidate_run(model: torch.nn.Module, x, tgt, cond, masks, grads: bool = True, loss_fn=None) -> Dict:
    """The CUDA path on the model's device with injected masks.  Gradients are read from .grad (fresh tensors)."""
    dev = next(model.parameters()).device
    prev_override, was_training = model.drop_masks_override, model.training
    saved_grads = [p.grad for p in model.parameters()]
    try:
        if masks is not None:
            model.train()
            model.drop_masks_override = [tuple(None if m is None else m.to(dev) for m in trip) for trip in masks]
        else:
            model.eval()
        for p in model.parameters():
            p.grad = None
        xd = x.to(dev).requires_grad_(grads)
        args = (xd,) if cond is None else (xd, cond.to(dev))
        if not grads:
            with torch.no_grad():
                return dict(y=model(*args).float().cpu())
        y = model(*args)
        loss = (loss_fn or O.rel_l2_loss)(y, tgt.to(dev))
        loss.backward()
        torch.cuda.synchronize(dev)
        return dict(y=y.detach().float().cpu(), loss=float(loss), dx=xd.grad.float().cpu(),
                    grads={k: (p.grad.float().cpu() if p.grad is not None else None)
                           for k, p in model.named_parameters()})
    finally:
        model.drop_masks_override = prev_override
        model.train(was_training)
        for p, g in zip(model.parameters(), saved_grads):
            p.grad = g


def compare(ref: Dict, got: Dict, verbose_prefix: Optional[str] = None, top: int = 12) -> Dict:
    """rel-L2 per output channel, loss, dx, global-norm-relative gradient error (+ the worst tensors)."""
    res = {"fwd_rel_l2_per_channel": [_rel(got["y"][:, :, c], ref["y"][:, :, c]) for c in range(ref["y"].shape[2])]}
    res["fwd_rel_l2"] = max(res["fwd_rel_l2_per_channel"])
    if "grads" in ref and "grads" in got:
        res["loss_rel"] = abs(got["loss"] - ref["loss"]) / max(abs(ref["loss"]), 1e-30)
        res["dx_rel_l2"] = _rel(got["dx"], ref["dx"])
        gn = float(np.sqrt(sum(float((g.double() ** 2).sum()) for g in ref["grads"].values())))
        tot, worst, missing = 0.0, [], []
        for k, g in ref["grads"].items():
            c = got["grads"].get(k)
            if c is None:
                missing.append(k)
                continue
            d = float((c.double() - g.double()).norm())
            tot += d * d
            worst.append((d / gn, d / max(float(g.double().norm()), 1e-30), k))
        worst.sort(reverse=True)
        res["grad_rel"] = float(np.sqrt(tot)) / gn
        res["grad_missing"] = missing
        res["grad_worst"] = [(k, gr, r) for gr, r, k in worst[:top]]
    if verbose_prefix is not None:
        p = verbose_prefix
        print(f"[{p}] fwd rel-L2 per channel: " + " ".join(f"{e:.3e}" for e in res["fwd_rel_l2_per_channel"]))
        if "grad_rel" in res:
            print(f"[{p}] loss {got['loss']:.6f} ref {ref['loss']:.6f} (rel {res['loss_rel']:.2e})  dx rel-L2 "
                  f"{res['dx_rel_l2']:.3e}  parameter gradients global-norm-relative {res['grad_rel']:.3e}")
            for k, gr, r in res["grad_worst"]:
                print(f"   {k:62s} global-rel {gr:.3e}  rel {r:.3e}")
            if res["grad_missing"]:
                print(f"[{p}] NO GRADIENT for: {res['grad_missing']}")
    return res


def passes(res: Dict, fwd_tol: float = 1e-2, grad_tol: float = 2e-2) -> bool:
    ok = res["fwd_rel_l2"] < fwd_tol
    if "grad_rel" in res:
        ok = ok and res["dx_rel_l2"] < grad_tol and res["grad_rel"] < grad_tol and not res["grad_missing"]
    return bool(ok)
