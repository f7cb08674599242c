"""Host-side orchestration (bubbleformer_b200.engine + autograd glue + module tree) checked on CPU.

The CUDA kernels are replaced by torch emulations of their documented semantics (tests/cpu_emulation.py,
monkeypatched -- the product itself has no CPU path), with the 16-bit buffers widened to fp32, so that every
forward value and every parameter gradient of the engine can be compared with the fixtures the live reference
produced (tests/golden).  This pins the backward formulas (IN through raw sums, feature-scale constant,
layer-scale / drop-path bookkeeping, zero fc2-bias gradient, weight re-layouts of the patch stages).
"""
import numpy as np
import pytest
import torch

from bubbleformer_b200 import engine, get_model
from oracle import filmavit_oracle as O
from tests import cpu_emulation
from tests.helpers import load_case


def _wide_w16(p):
    return p.detach().reshape(p.shape[0], -1)


def _forward(model, case, x):
    B, T = case["B"], case["T"]
    _, _, C, H, W = x.shape
    p = model.patch_size
    geom = engine.Geom(B, T, H // p, W // p)
    xi = x.reshape(B * T, C, H, W)
    # the FiLM MLP runs inside the patch-embed Function (bf_film_fwd / bf_film_bwd), as in FiLMConditionedAViT.forward
    film = model.film_embed if case["cond"] is not None else None
    X = model.embed.tokens(xi, case["cond"], T, film=film)
    for i, blk in enumerate(model.blocks):
        masks = case["masks"][i] if case["masks"] is not None else None
        X = blk.tokens(X, geom, _wide_w16, masks)
    return model.debed.images(X, geom).view(B, T, -1, H, W)


@pytest.mark.parametrize("name", ["film_eval_e128", "film_train_masks_e128", "avit_generic_e96"])
def test_engine_orchestration_matches_reference_fixture(name, monkeypatch):
    cpu_emulation.install(monkeypatch, wide=True)
    case = load_case(name, dtype=torch.float32)
    model = get_model(case["model"], time_window=case["T"], **case["cfg"])
    model.load_state_dict(case["sd"], strict=True)
    model.train() if case["masks"] is not None else model.eval()
    x = case["x"].clone().requires_grad_(True)
    y = _forward(model, case, x)
    g = case["gold"]
    assert O.rel_l2(y, torch.from_numpy(g["y"])) < 2e-5
    O.rel_l2_loss(y, case["tgt"]).backward()
    assert O.rel_l2(x.grad, torch.from_numpy(g["dx"])) < 2e-4
    gn = np.sqrt(sum(float((g["grad/" + k].astype(np.float64) ** 2).sum()) for k, _ in model.named_parameters()))
    for k, p in model.named_parameters():
        ref = torch.from_numpy(g["grad/" + k]).double()
        assert p.grad is not None, k
        err = float((p.grad.double() - ref).norm()) / gn
        assert err < 2e-4, (k, err, float(ref.norm()) / gn)


def test_relpos_bucket_vector_matches_oracle():
    for Ln in (1, 2, 5, 8, 9, 16, 32, 33, 64):
        tab = O.relpos_bucket_table(Ln)
        vec = engine.relpos_bucket_vector(Ln, "cpu")
        assert vec.dtype == torch.int32 and vec.numel() == 2 * Ln - 1
        for i in range(Ln):
            for j in range(Ln):
                assert int(vec[j - i + Ln - 1]) == int(tab[i, j])


def test_pick_split_in_range():
    for tokens in (40960, 5120, 163840, 1000, 64, 63):
        for (m, n) in ((384, 384), (1152, 384), (1536, 384), (96, 384)):
            s = engine.pick_split(tokens, m, n)
            assert 1 <= s <= (tokens + 63) // 64


def test_fused_rel_l2_loss_matches_oracle(monkeypatch):
    """Host side of the fused loss (scalar assembly, gradient coefficients) against the oracle's LpLoss restatement."""
    cpu_emulation.install(monkeypatch, wide=True)
    from bubbleformer_b200 import losses
    monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda self: True))
    g = torch.Generator().manual_seed(0)
    pred = torch.randn(2, 3, 4, 8, 8, generator=g, requires_grad=True)
    tgt = torch.randn(2, 3, 4, 8, 8, generator=g)
    loss = losses.rel_l2_loss(pred, tgt)
    ref_in = pred.detach().clone().requires_grad_(True)
    ref = O.rel_l2_loss(ref_in, tgt)
    assert abs(float(loss) - float(ref)) < 1e-6 * abs(float(ref))
    loss.backward()
    ref.backward()
    assert O.rel_l2(pred.grad, ref_in.grad) < 1e-6
