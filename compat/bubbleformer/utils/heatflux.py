"""`bubbleformer.utils.heatflux` (upstream utils/heatflux.py:3-38): numpy in, (mean, max) floats out, computed on the GPU."""
import numpy as np
import torch

from bubbleformer_b200 import metrics


def heatflux(dfun: np.ndarray, temp: np.ndarray, heater_temp: int):
    d = torch.as_tensor(np.asarray(dfun), dtype=torch.float32).cuda()
    t = torch.as_tensor(np.asarray(temp), dtype=torch.float32).cuda()
    mean, mx = metrics.heatflux(d, t, float(heater_temp))
    return float(mean), float(mx)
