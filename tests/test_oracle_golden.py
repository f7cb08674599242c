"""The CPU oracle must reproduce the fixtures that the live reference produced.

Fixtures: tests/golden/*.npz, written by oracle/make_golden.py in the build container by
executing the unmodified upstream modules (float64).  This pins the oracle: the reference
itself ships no golden vectors for this path (shape-only tests).
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle import filmavit_oracle as O
from oracle.param_init import fluid_params, param_shapes, random_state_dict
from tests.helpers import GOLD, load_case, case_masks


@pytest.mark.parametrize("name", ["film_eval_e128", "film_train_masks_e128", "avit_generic_e96"])
def test_oracle_matches_reference_fixture(name):
    case = load_case(name, dtype=torch.float64)
    sd = {k: v.clone().requires_grad_(True) for k, v in case["sd"].items()}
    x = case["x"].clone().requires_grad_(True)
    y = O.forward(sd, x, case["cond"], drop_masks=case["masks"], **case["fw"])
    loss = O.rel_l2_loss(y, case["tgt"])
    loss.backward()
    g = case["gold"]
    tol = 1e-12 if g["y"].dtype == np.float64 else 2e-7      # float32-stored fixtures
    assert O.rel_l2(y, torch.from_numpy(g["y"])) < tol
    assert abs(float(loss.detach()) - float(g["loss"])) < 1e-9 * abs(float(g["loss"]))
    assert O.rel_l2(x.grad, torch.from_numpy(g["dx"])) < tol
    gn = np.sqrt(sum(float((g["grad/" + k].astype(np.float64) ** 2).sum()) for k in sd))
    for k, p in sd.items():
        ref = torch.from_numpy(g["grad/" + k]).double()
        assert p.grad.shape == ref.shape, k
        assert float((p.grad - ref).norm()) / gn < tol, k


def test_oracle_fp32_rollout_fixture():
    """Config 1: film_avit_small on sample_1 frames, 10 autoregressive steps (teacher forced here)."""
    z = np.load(os.path.join(GOLD, "rollout_sample1_small.npz"))
    meta = json.loads(str(z["meta"]))
    cfg = meta["cfg"]
    shapes = param_shapes(**{k: v for k, v in cfg.items() if k != "drop_path"})
    sd = random_state_dict(shapes, seed=meta["seed"], dtype=torch.float32)
    cond = fluid_params(1, torch.float32)
    preds = torch.from_numpy(z["preds"])
    inp = torch.from_numpy(z["inp0"]).unsqueeze(0)
    with torch.no_grad():
        for s in (0, 1, 4):          # three of the ten steps keep the CPU suite short
            y = O.forward(sd, inp if s == 0 else preds[s - 1].unsqueeze(0), cond, patch_size=16, num_heads=6)
            for c in range(4):
                assert O.rel_l2(y[0, :, c], preds[s][:, c]) < 2e-5, (s, c)


def test_relpos_bucket_tables():
    """Bucket tables quoted in SURVEY.md 7.1 (probed from the reference)."""
    t5 = O.relpos_bucket_table(5)
    assert t5[0].tolist() == [0, 17, 18, 19, 20]
    assert t5[:, 0].tolist() == [0, 1, 2, 3, 4]
    t32 = O.relpos_bucket_table(32)
    upper = [0] + list(range(17, 25)) + [24, 25, 25, 26, 26, 27, 27] + [28] * 4 + [29] * 3 + [30] * 4 + [31] * 5
    assert t32[0].tolist() == upper
    assert t32[:, 0].tolist() == [0] + [u - 16 for u in upper[1:]]
    assert 16 not in t32.unique().tolist()
