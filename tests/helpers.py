"""Shared helpers for the parity tests (fixture loading, error metrics)."""
import json
import os

import numpy as np
import torch

from oracle.param_init import fluid_params, param_shapes, random_state_dict

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def case_masks(cfg, B, T, seed, dtype):
    """Re-draw the drop-path masks exactly as oracle/make_golden.py::draw_masks did."""
    g = np.random.RandomState(seed)
    rates = np.linspace(0, cfg["drop_path"], cfg["processor_blocks"])
    masks = []
    for p in rates:
        if p == 0.0:
            masks.append((None, None, None))
            continue
        keep = 1.0 - p
        masks.append(tuple(torch.from_numpy((g.uniform(size=n) < keep).astype(np.float64) / keep).to(dtype)
                           for n in (B, B * T, B * T)))
    return masks


def load_case(name, dtype=torch.float32):
    """Rebuild inputs + weights of a golden case and return them with the reference outputs."""
    z = np.load(os.path.join(GOLD, name + ".npz"))
    meta = json.loads(str(z["meta"]))
    cfg, B, T, H, W, seed = (meta[k] for k in ("cfg", "B", "T", "H", "W", "seed"))
    is_film = meta["model"] == "filmavit"
    shapes = param_shapes(
        input_fields=cfg["input_fields"], output_fields=cfg["output_fields"], patch_size=cfg["patch_size"],
        embed_dim=cfg["embed_dim"], num_heads=cfg["num_heads"], processor_blocks=cfg["processor_blocks"],
        attn_scale=cfg["attn_scale"], feat_scale=cfg["feat_scale"],
        num_fluid_params=cfg.get("num_fluid_params") if is_film else None)
    sd = random_state_dict(shapes, seed=seed, dtype=dtype)
    g = np.random.RandomState(seed + 1000)
    x = torch.from_numpy(g.standard_normal((B, T, cfg["input_fields"], H, W))).to(dtype)
    tgt = torch.from_numpy(g.standard_normal((B, T, cfg["output_fields"], H, W))).to(dtype)
    cond = fluid_params(B, dtype) if is_film else None
    masks = case_masks(cfg, B, T, seed + 2000, dtype) if meta["train_masks"] else None
    fw = dict(patch_size=cfg["patch_size"], num_heads=cfg["num_heads"],
              attn_scale=cfg["attn_scale"], feat_scale=cfg["feat_scale"])
    return dict(meta=meta, cfg=cfg, model=meta["model"], sd=sd, x=x, tgt=tgt, cond=cond, masks=masks,
                fw=fw, gold=z, B=B, T=T)
