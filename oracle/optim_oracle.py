"""CPU restatement of the optimiser updates upstream applies to the hot path's parameters -- TEST INFRASTRUCTURE ONLY
(imported by tests/; the product path is bubbleformer_b200/csrc/optim.cu and fails without the CUDA library).

  AdamW / Adam: torch.optim (upstream bubbleformer/modules.py:134-138); pinned in tests against torch.optim itself.
  Lion: lion_pytorch.Lion (modules.py:139, env/requirements.txt, unpinned, NOT installed here) -- its published
        `update_fn`: p *= 1 - lr*wd; p -= lr*sign(b1*m + (1-b1)*g); m = b2*m + (1-b2)*g.  Unpinned against the
        lion_pytorch PACKAGE (no copy to run); pinned instead against an independent torch.optim.Optimizer written from
        the published algorithm (Chen et al. 2023, "Symbolic Discovery of Optimization Algorithms", Algorithm 1:
        c = sign(b1 m + (1-b1) g); theta <- theta - lr (c + wd theta); m <- b2 m + (1-b2) g), 100 steps with weight
        decay, in tests/test_optim.py::test_oracle_lion_matches_independent_torch_optimizer.
  Schedule: utils/lr_schedulers.py:4-31 (CosineWarmupLR), pinned in tests against torch's SequentialLR.
"""
import torch


def lion_step(p, g, m, lr, b1, b2, wd):
    p = p * (1 - lr * wd)
    p = p - lr * torch.sign(b1 * m + (1 - b1) * g)
    m = b2 * m + (1 - b2) * g
    return p, m


def adamw_step(p, g, m, v, lr, b1, b2, eps, wd, t, decoupled=True):
    if decoupled:
        p = p * (1 - lr * wd)
    else:
        g = g + wd * p
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    bc1, bc2 = 1 - b1 ** t, 1 - b2 ** t
    p = p - (lr / bc1) * m / (v.sqrt() / bc2 ** 0.5 + eps)
    return p, m, v
