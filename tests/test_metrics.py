"""Rollout metrics: oracle pinned to reference outputs on CPU, CUDA kernels against both on GPU."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden", "metrics.npz")


def test_metrics_oracle_matches_reference_fixture():
    from oracle import metrics_oracle as O
    g = np.load(GOLD)
    assert abs(O.eikonal_loss(g["phi"]) - float(g["eikonal"])) < 1e-6 * float(g["eikonal"])      # fixture fields are fp32
    m, mx = O.heatflux(g["dfun_row"], g["temp_row"], float(g["heater_temp"]))
    assert abs(m - float(g["hflux_mean"])) < 1e-6 * float(g["hflux_mean"])
    assert abs(mx - float(g["hflux_max"])) < 1e-6 * float(g["hflux_max"])
    assert abs(O.rel_l2_per_field(g["pred"], g["tgt"]).mean() - float(g["rel_l2"])) < 1e-6


@pytest.mark.gpu
def test_metrics_kernels_match_reference_fixture():
    import torch
    from bubbleformer_b200 import metrics
    from oracle import metrics_oracle as O
    g = np.load(GOLD)
    dev = "cuda"
    eik = float(metrics.eikonal_loss(torch.tensor(g["phi"], device=dev)))
    assert abs(eik - float(g["eikonal"])) < 1e-5 * float(g["eikonal"])
    m, mx = metrics.heatflux(torch.tensor(g["dfun_row"], device=dev), torch.tensor(g["temp_row"], device=dev),
                             float(g["heater_temp"]))
    assert abs(float(m) - float(g["hflux_mean"])) < 1e-5 * float(g["hflux_mean"])
    assert abs(float(mx) - float(g["hflux_max"])) < 1e-5 * float(g["hflux_max"])
    rel = metrics.rel_l2_per_field(torch.tensor(g["pred"], device=dev), torch.tensor(g["tgt"], device=dev))
    assert float((rel.cpu().double() - torch.tensor(O.rel_l2_per_field(g["pred"], g["tgt"]))).abs().max()) < 1e-6
    assert abs(float(rel.mean()) - float(g["rel_l2"])) < 1e-6
    # full-size fields (512 x 512, the domain heatflux.py assumes) against the oracle; ragged sizes for the stencil edges
    torch.manual_seed(0)
    phi = torch.randn(2, 3, 37, 53, device=dev)
    assert abs(float(metrics.eikonal_loss(phi)) - O.eikonal_loss(phi.cpu().numpy())) < 1e-5 * O.eikonal_loss(phi.cpu().numpy())
    d, t = torch.randn(7, 512, 512, device=dev), 60 + 30 * torch.rand(7, 512, 512, device=dev)
    m, mx = metrics.heatflux(d, t, 95.0)
    rm, rmx = O.heatflux(d.cpu().numpy(), t.cpu().numpy(), 95.0)
    assert abs(float(m) - rm) < 1e-5 * abs(rm) and abs(float(mx) - rmx) < 1e-5 * abs(rmx)


def test_lploss_inference_config_matches_reference_fixture(monkeypatch):
    """LpLoss(d=2, p=2, reduce_dims=[0,1], reductions=[mean, mean]) (scripts/inference.py:231): host side of the fused
    loss on the CPU emulation of the two kernels, against the value the reference printed for the fixture."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cpu_emulation
    cpu_emulation.install(monkeypatch, wide=True)
    monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda self: True))
    from bubbleformer_b200.losses import LpLoss
    g = np.load(GOLD)
    crit = LpLoss(d=2, p=2, reduce_dims=[0, 1], reductions=["mean", "mean"])
    assert abs(float(crit(torch.tensor(g["pred"]), torch.tensor(g["tgt"]))) - float(g["rel_l2"])) < 1e-6
    with pytest.raises(NotImplementedError):
        LpLoss(d=1, p=2)(torch.zeros(2, 3, 4), torch.ones(2, 3, 4))
