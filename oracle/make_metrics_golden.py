"""Generate tests/golden/metrics.npz by executing the UNMODIFIED reference metric functions -- TEST INFRASTRUCTURE ONLY.

Runs only where the upstream checkout is mounted (/root/reference); the two modules are loaded by file path because
the `bubbleformer` package __init__ pulls in Lightning.     python oracle/make_metrics_golden.py
"""
import importlib.util
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("BUBBLEFORMER_REFERENCE", "/root/reference")


def load(rel, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    losses = load("bubbleformer/utils/losses.py", "_ref_losses")
    hf = load("bubbleformer/utils/heatflux.py", "_ref_heatflux")
    rng = np.random.default_rng(7)
    yy, xx = np.meshgrid(np.arange(64) / 32.0, np.arange(96) / 32.0, indexing="ij")
    phi = np.stack([np.sqrt((xx - 1.3 - 0.1 * k) ** 2 + (yy - 0.9) ** 2) - 0.5 + 0.02 * rng.standard_normal(xx.shape)
                    for k in range(6)]).reshape(2, 3, 64, 96)
    eik = float(losses.eikonal_loss(torch.tensor(phi, dtype=torch.float64)))
    dfun = rng.standard_normal((5, 512, 512))
    temp = rng.uniform(50.0, 95.0, size=(5, 512, 512))
    hmean, hmax = hf.heatflux(dfun, temp, 91)
    pred = rng.standard_normal((4, 3, 32, 48))
    tgt = pred + 0.1 * rng.standard_normal(pred.shape)
    crit = losses.LpLoss(d=2, p=2, reduce_dims=[0, 1], reductions=["mean", "mean"])
    rel = float(crit(torch.tensor(pred), torch.tensor(tgt)))
    out = os.path.join(os.path.dirname(HERE), "tests", "golden", "metrics.npz")
    np.savez_compressed(out, phi=phi.astype(np.float32), eikonal=eik, dfun_row=dfun[:, :2].astype(np.float32),
                        temp_row=temp[:, :2].astype(np.float32), heater_temp=91.0, hflux_mean=hmean, hflux_max=hmax,
                        pred=pred.astype(np.float32), tgt=tgt.astype(np.float32), rel_l2=rel)
    print("wrote", out, dict(eikonal=eik, hflux=(hmean, hmax), rel_l2=rel))


if __name__ == "__main__":
    main()
