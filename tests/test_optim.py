"""Optimiser step: oracle pinned against torch.optim on CPU; CUDA kernel against the oracle on GPU."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_oracle_adam_family_matches_torch_optim():
    from oracle import optim_oracle as O
    torch.manual_seed(0)
    for cls, decoupled, wd in ((torch.optim.AdamW, True, 1e-2), (torch.optim.Adam, False, 1e-5)):
        p = torch.randn(1000, dtype=torch.float64)
        ref = torch.nn.Parameter(p.clone())
        opt = cls([ref], lr=2.5e-4, weight_decay=wd)
        m, v = torch.zeros_like(p), torch.zeros_like(p)
        for t in range(1, 6):
            g = torch.randn_like(p)
            ref.grad = g.clone()
            opt.step()
            p, m, v = O.adamw_step(p, g, m, v, 2.5e-4, 0.9, 0.999, 1e-8, wd, t, decoupled)
        assert float((p - ref.detach()).abs().max()) < 1e-12


class _PaperLion(torch.optim.Optimizer):
    """Independent Lion written from the paper's Algorithm 1 (Chen et al. 2023), in-place torch ops, decoupled weight decay
    inside the update:  theta <- theta - lr * (sign(b1 m + (1 - b1) g) + wd * theta);  m <- b2 m + (1 - b2) g."""

    def __init__(self, params, lr, betas, weight_decay):
        super().__init__(params, dict(lr=lr, betas=betas, weight_decay=weight_decay))

    @torch.no_grad()
    def step(self):
        for grp in self.param_groups:
            b1, b2 = grp["betas"]
            for p in grp["params"]:
                st = self.state[p]
                if "m" not in st:
                    st["m"] = torch.zeros_like(p)
                c = torch.sign(torch.lerp(p.grad, st["m"], b1))            # b1*m + (1-b1)*g
                p.add_(c + grp["weight_decay"] * p, alpha=-grp["lr"])
                st["m"].lerp_(p.grad, 1 - b2)                               # b2*m + (1-b2)*g


def test_oracle_lion_matches_independent_torch_optimizer():
    """Lion is upstream's default optimiser (config/optim_cfg/lion.yaml:1-5: lr 0.5e-4, weight_decay 1e-1; lion_pytorch's
    default betas (0.9, 0.99)).  100 steps, fp64: the oracle's restatement of lion_pytorch's update_fn (decay first, then
    the signed step) against the paper's form (decay inside the step) -- same parameters and momenta up to rounding,
    including the order of the momentum update (after the parameter step, from the OLD momentum)."""
    from oracle import optim_oracle as O
    torch.manual_seed(3)
    lr, b1, b2, wd = 0.5e-4, 0.9, 0.99, 1e-1
    p = torch.randn(4096, dtype=torch.float64)
    ref = torch.nn.Parameter(p.clone())
    opt = _PaperLion([ref], lr, (b1, b2), wd)
    m = torch.zeros_like(p)
    for _ in range(100):
        g = torch.randn_like(p)
        # keep b1*m + (1-b1)*g away from 0 so that sign() cannot differ by rounding between the two forms
        interp = b1 * m + (1 - b1) * g
        g = torch.where(interp.abs() < 1e-9, g + 1e-6, g)
        ref.grad = g.clone()
        opt.step()
        p, m = O.lion_step(p, g, m, lr, b1, b2, wd)
    assert float((p - ref.detach()).abs().max()) < 1e-12
    assert float((m - opt.state[ref]["m"]).abs().max()) < 1e-12


def test_cosine_warmup_matches_sequential_lr():
    from torch.optim.lr_scheduler import CosineAnnealingLR, LambdaLR, SequentialLR
    from bubbleformer_b200.optim import cosine_warmup_lr
    w, T, eta, base = 5, 40, 1e-6, 2.5e-4
    opt = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=base)
    sched = SequentialLR(opt, schedulers=[LambdaLR(opt, lr_lambda=lambda s: s / w), CosineAnnealingLR(opt, T_max=T, eta_min=eta)],
                         milestones=[w])
    for step in range(30):
        assert abs(opt.param_groups[0]["lr"] - cosine_warmup_lr(step, base, w, T, eta)) < 1e-12, step
        opt.step()
        sched.step()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["lion", "adamw", "adam"])
def test_flat_optimizer_matches_oracle(name):
    from bubbleformer_b200 import get_model
    from bubbleformer_b200.optim import FlatOptimizer
    from bubbleformer_b200.parallel import GradSink
    from oracle import optim_oracle as O
    torch.manual_seed(1)
    m = get_model("filmavit", input_fields=4, output_fields=4, time_window=5, patch_size=16, embed_dim=128, num_heads=2,
                  processor_blocks=1, num_fluid_params=9).cuda()
    sink = GradSink(m)
    try:
        opt = FlatOptimizer(m, sink, name)
        p = opt.bank.flat.double().cpu()
        mm, vv = torch.zeros_like(p), torch.zeros_like(p)
        b1, b2 = opt.betas
        for t in range(1, 5):
            g = torch.randn_like(opt.bank.flat) * (opt.bank.flat != 0)      # padding slots keep zero gradients
            sink.flat.copy_(g)
            opt.step()
            gd = g.double().cpu()
            if name == "lion":
                p, mm = O.lion_step(p, gd, mm, opt.lr, b1, b2, opt.weight_decay)
            else:
                p, mm, vv = O.adamw_step(p, gd, mm, vv, opt.lr, b1, b2, opt.eps, opt.weight_decay, t, name == "adamw")
        got = opt.bank.flat.double().cpu()
        bad = int(((got - p).abs() > 2e-6).sum())
        # Lion: sign(b1*m + (1-b1)*g) may flip where the fp32 and fp64 interpolations straddle zero (a handful of the
        # ~1e6 elements at most); everything else, and every Adam element, agrees to fp32 rounding
        assert bad <= (8 if name == "lion" else 0), bad
        # the parameters are still views of the flat buffer and the bf16 mirror follows them
        w = m.blocks[0].spatial.mlp.fc1.weight
        assert w.data_ptr() >= opt.bank.flat.data_ptr()
        assert float((opt.bank.flat16.float() - opt.bank.flat).abs().max()) <= 2 ** -8 * float(opt.bank.flat.abs().max())
        from bubbleformer_b200 import _lib
        n0 = _lib.launch_count()
        opt.bank.refresh()                              # mirror is fresh: no cast launch
        assert _lib.launch_count() == n0
        with torch.no_grad():
            w.mul_(2.0)
        opt.bank.refresh()                              # any other in-place change of a parameter invalidates it
        assert _lib.launch_count() == n0 + 1
        assert float((opt.bank.flat16.float() - opt.bank.flat).abs().max()) <= 2 ** -8 * float(opt.bank.flat.abs().max())
    finally:
        sink.close()
