"""Generate tests/golden/*.npz by executing the UNMODIFIED reference -- TEST INFRASTRUCTURE ONLY.

Runs only in the build container, where the upstream checkout is mounted read-only at
/root/reference (it does not exist on the GPU box; nothing at test/bench time reads it).
The reference imports `timm.layers.DropPath`, which is not installed here; a stand-in
with timm's semantics (mask ~ Bernoulli(keep)/keep over dim 0, identity in eval or p=0)
is registered in sys.modules.  For train-mode fixtures the stand-in pops pre-drawn masks
from a queue so the candidate can be given the identical masks.

    python oracle/make_golden.py            # rewrites tests/golden/*.npz

Every fixture is produced by the reference in float64 and stored as float32 (float64 for
the tiny ones), together with the configuration needed to regenerate inputs and weights
from `oracle/param_init.py`.
"""
from __future__ import annotations

import json
import os
import struct
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("BUBBLEFORMER_REFERENCE", "/root/reference")
GOLD = os.path.join(ROOT, "tests", "golden")

MASK_QUEUE: list = []          # filled per forward for train-mode fixtures
MASK_LOG: list = []


def install_timm_shim() -> None:
    class DropPath(torch.nn.Module):
        def __init__(self, drop_prob: float = 0.0, scale_by_keep: bool = True):
            super().__init__()
            self.drop_prob = drop_prob
            self.scale_by_keep = scale_by_keep

        def forward(self, x):
            if self.drop_prob == 0.0 or not self.training:
                return x
            m = MASK_QUEUE.pop(0).to(x.dtype)
            assert m.numel() == x.shape[0]
            MASK_LOG.append(m.clone())
            return x * m.reshape((x.shape[0],) + (1,) * (x.ndim - 1))

    timm = types.ModuleType("timm")
    layers = types.ModuleType("timm.layers")
    layers.DropPath = DropPath
    timm.layers = layers
    sys.modules["timm"] = timm
    sys.modules["timm.layers"] = layers


def load_reference():
    install_timm_shim()
    sys.path.insert(0, REF)
    from bubbleformer.models import get_model  # noqa: the reference package
    return get_model


sys.path.insert(0, ROOT)
from oracle.param_init import param_shapes, random_state_dict, fluid_params  # noqa: E402
from oracle import filmavit_oracle as O  # noqa: E402


def read_sample_hdf5(path: str) -> np.ndarray:
    """Raw reader for samples/sample_N.hdf5 (h5py is not installed).

    The files are HDF5 superblock v0 with four contiguous little-endian float32
    datasets of shape (50, 64, 64); byte offsets probed during the survey.
    Returns (50, 4, 64, 64) in field order dfun, temperature, velx, vely.
    """
    offs = {"dfun": 2048, "temperature": 821248, "velx": 1640448, "vely": 2461696}
    n = 50 * 64 * 64
    with open(path, "rb") as f:
        buf = f.read()
    assert buf[:8] == b"\x89HDF\r\n\x1a\n"
    out = [np.frombuffer(buf, dtype="<f4", count=n, offset=o).reshape(50, 64, 64) for o in offs.values()]
    return np.stack(out, axis=1).copy()


def draw_masks(cfg: dict, B: int, T: int, seed: int):
    """Per block (mask_b, mask_att, mask_mlp) like timm: bernoulli(keep)/keep. Block 0 has p=0."""
    g = np.random.RandomState(seed)
    rates = np.linspace(0, cfg["drop_path"], cfg["processor_blocks"])
    masks = []
    for p in rates:
        if p == 0.0:
            masks.append((None, None, None))
            continue
        keep = 1.0 - p
        trip = [torch.from_numpy((g.uniform(size=n) < keep).astype(np.float64) / keep)
                for n in (B, B * T, B * T)]
        masks.append(tuple(trip))
    return masks


def run_case(get_model, name: str, model_name: str, cfg: dict, B: int, T: int, H: int, W: int,
             seed: int, train_masks: bool, store64: bool) -> None:
    torch.manual_seed(0)
    is_film = model_name == "filmavit"
    kwargs = dict(cfg)
    model = get_model(model_name, time_window=T, **kwargs).double()
    shapes = param_shapes(
        input_fields=cfg["input_fields"], output_fields=cfg["output_fields"], patch_size=cfg["patch_size"],
        embed_dim=cfg["embed_dim"], num_heads=cfg["num_heads"], processor_blocks=cfg["processor_blocks"],
        attn_scale=cfg["attn_scale"], feat_scale=cfg["feat_scale"],
        num_fluid_params=cfg.get("num_fluid_params") if is_film else None)
    sd = random_state_dict(shapes, seed=seed, dtype=torch.float64)
    missing = model.load_state_dict(sd, strict=True)     # proves the inventory is exact
    assert list(model.state_dict().keys()) == list(sd.keys()), "registration order differs"
    g = np.random.RandomState(seed + 1000)
    x = torch.from_numpy(g.standard_normal((B, T, cfg["input_fields"], H, W)))
    tgt = torch.from_numpy(g.standard_normal((B, T, cfg["output_fields"], H, W)))
    cond = fluid_params(B, torch.float64) if is_film else None
    x.requires_grad_(True)
    masks = None
    if train_masks:
        model.train()
        masks = draw_masks(cfg, B, T, seed + 2000)
        MASK_QUEUE.clear()
        for trip in masks:
            MASK_QUEUE.extend([m for m in trip if m is not None])
    else:
        model.eval()
    y = model(x, cond) if is_film else model(x)
    assert not MASK_QUEUE
    loss = O.rel_l2_loss(y, tgt)
    loss.backward()
    # the oracle must agree with the live reference to fp64 round-off before we trust either
    xo = x.detach().clone().requires_grad_(True)
    sdo = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    yo = O.forward(sdo, xo, cond, patch_size=cfg["patch_size"], num_heads=cfg["num_heads"],
                   attn_scale=cfg["attn_scale"], feat_scale=cfg["feat_scale"], drop_masks=masks)
    O.rel_l2_loss(yo, tgt).backward()
    err_y = O.rel_l2(yo, y)
    err_x = O.rel_l2(xo.grad, x.grad)
    worst = 0.0
    gnorm = torch.sqrt(sum((p.grad ** 2).sum() for p in model.parameters()))
    for (k, p) in model.named_parameters():
        e = float((sdo[k].grad - p.grad).norm() / gnorm)
        worst = max(worst, e)
    print(f"[{name}] oracle vs live reference (fp64): out {err_y:.2e}  dx {err_x:.2e}  "
          f"worst param-grad (global-norm rel) {worst:.2e}  loss {float(loss):.6f}")
    assert err_y < 1e-12 and err_x < 1e-11 and worst < 1e-11, "oracle does not restate the reference"
    dt = np.float64 if store64 else np.float32
    out = {
        "meta": json.dumps(dict(model=model_name, cfg=cfg, B=B, T=T, H=H, W=W, seed=seed,
                                train_masks=train_masks, torch=torch.__version__)),
        "y": y.detach().numpy().astype(dt),
        "loss": np.float64(loss.item()),
        "dx": x.grad.numpy().astype(dt),
    }
    for k, p in model.named_parameters():
        out["grad/" + k] = p.grad.numpy().astype(dt)
    os.makedirs(GOLD, exist_ok=True)
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)


def run_rollout(get_model, name: str, steps: int = 10, seed: int = 42) -> None:
    """Config 1: film_avit_small on samples/sample_1.hdf5, autoregressive, fp32 like the reference runs it."""
    cfg = dict(input_fields=4, output_fields=4, patch_size=16, embed_dim=384, num_heads=6,
               processor_blocks=12, drop_path=0.2, attn_scale=True, feat_scale=True, num_fluid_params=9)
    model = get_model("filmavit", time_window=5, **cfg)
    shapes = param_shapes(**{k: v for k, v in cfg.items() if k != "drop_path"})
    sd = random_state_dict(shapes, seed=seed, dtype=torch.float32)
    model.load_state_dict(sd, strict=True)
    model.eval()
    data = read_sample_hdf5(os.path.join(REF, "samples", "sample_1.hdf5"))
    inp0 = torch.from_numpy(data[:5]).float().unsqueeze(0)          # (1, 5, 4, 64, 64)
    cond = fluid_params(1, torch.float32)
    preds = []
    inp = inp0
    with torch.no_grad():
        for _ in range(steps):
            pred = model(inp, cond)
            preds.append(pred[0].numpy().copy())
            inp = pred
        # divergence envelope: the reference's own free-running response to a bf16-sized input perturbation
        pert = inp0.to(torch.bfloat16).float()
        env = []
        inp = pert
        for s in range(steps):
            pred = model(inp, cond)
            env.append([O.rel_l2(pred[0, :, c], torch.from_numpy(preds[s][:, c])) for c in range(4)])
            inp = pred
        # oracle cross-check on step 1
        yo = O.forward(sd, inp0, cond, patch_size=16, num_heads=6)
    print(f"[{name}] oracle vs live reference (fp32) step1: {O.rel_l2(yo[0], torch.from_numpy(preds[0])):.2e}")
    print(f"[{name}] reference chaos envelope (bf16-rounded input), rel-L2 per step/channel:")
    for s, e in enumerate(env):
        print("   step", s + 1, " ".join(f"{v:.2e}" for v in e))
    np.savez_compressed(
        os.path.join(GOLD, name + ".npz"),
        meta=json.dumps(dict(model="filmavit", cfg=cfg, seed=seed, steps=steps, torch=torch.__version__,
                             source="samples/sample_1.hdf5 frames 0..4, fields dfun,temperature,velx,vely")),
        inp0=inp0[0].numpy(), preds=np.stack(preds).astype(np.float32),
        envelope=np.asarray(env, dtype=np.float64))


def main() -> None:
    get_model = load_reference()
    small = dict(input_fields=4, output_fields=4, patch_size=16, embed_dim=128, num_heads=2,
                 processor_blocks=2, drop_path=0.0, attn_scale=True, feat_scale=True, num_fluid_params=9)
    run_case(get_model, "film_eval_e128", "filmavit", small, B=2, T=3, H=64, W=64, seed=1,
             train_masks=False, store64=False)
    dp = dict(small, drop_path=0.5, processor_blocks=3)
    run_case(get_model, "film_train_masks_e128", "filmavit", dp, B=3, T=2, H=32, W=64, seed=2,
             train_masks=True, store64=False)
    generic = dict(input_fields=2, output_fields=1, patch_size=8, embed_dim=96, num_heads=2,
                   processor_blocks=1, drop_path=0.0, attn_scale=False, feat_scale=False)
    run_case(get_model, "avit_generic_e96", "avit", generic, B=1, T=2, H=32, W=48, seed=3,
             train_masks=False, store64=True)
    run_rollout(get_model, "rollout_sample1_small")


if __name__ == "__main__":
    main()
