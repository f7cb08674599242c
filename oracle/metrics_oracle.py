"""CPU restatement of upstream's rollout metrics -- TEST INFRASTRUCTURE ONLY (imported by tests/).

  eikonal_loss: bubbleformer/utils/losses.py:5-15      heatflux: bubbleformer/utils/heatflux.py:3-38
  rel_l2: LpLoss(d=2, p=2, reduce_dims=[0,1], reductions=["mean","mean"]), utils/losses.py:67-94 as used at
  scripts/inference.py:231.
Pinned against the live reference functions by oracle/make_metrics_golden.py -> tests/golden/metrics.npz.
"""
import numpy as np


def eikonal_loss(phi, dx=1.0 / 32):
    phi = np.asarray(phi, dtype=np.float64)
    gy = np.empty_like(phi)
    gx = np.empty_like(phi)
    gy[..., 1:-1, :] = (phi[..., 2:, :] - phi[..., :-2, :]) / (2 * dx)
    gy[..., 0, :] = (phi[..., 1, :] - phi[..., 0, :]) / dx
    gy[..., -1, :] = (phi[..., -1, :] - phi[..., -2, :]) / dx
    gx[..., :, 1:-1] = (phi[..., :, 2:] - phi[..., :, :-2]) / (2 * dx)
    gx[..., :, 0] = (phi[..., :, 1] - phi[..., :, 0]) / dx
    gx[..., :, -1] = (phi[..., :, -1] - phi[..., :, -2]) / dx
    return float(((np.sqrt(gy ** 2 + gx ** 2) - 1.0) ** 2).mean())


def heatflux(dfun, temp, heater_temp, dx=1.0 / 32, lc=0.0007, x_min=-8.0):
    dfun, temp = np.asarray(dfun, dtype=np.float64), np.asarray(temp, dtype=np.float64)
    W = dfun.shape[-1]
    xc = x_min + (np.arange(W) + 0.5) * dx
    heater = (xc >= -5.0) & (xc <= 5.0)
    row = (heater[None, :] & (dfun[:, 0, :] < 0)) * (heater_temp - temp[:, 0, :])
    flux = (0.054 * row / (dx * lc)).mean(axis=1)
    return float(flux.mean()), float(flux.max())


def rel_l2_per_field(pred, tgt):
    pred, tgt = np.asarray(pred, dtype=np.float64), np.asarray(tgt, dtype=np.float64)
    num = np.sqrt(((pred - tgt) ** 2).sum(axis=(-2, -1)))
    den = np.sqrt((tgt ** 2).sum(axis=(-2, -1)))
    return (num / den).mean(axis=0)
