"""Deterministic parameter inventory + randomiser for parity tests -- TEST INFRASTRUCTURE ONLY.

`param_shapes` restates the reference `state_dict` inventory (checked
by `oracle/make_golden.py` with a strict `load_state_dict` into the live reference).
`random_state_dict` draws every tensor from a numpy stream keyed by (seed, name) so
the same weights can be regenerated anywhere without shipping them.

Why randomise: at the reference's default init (layer scale 1e-6, freq scalars 0,
attn scale 1) the 12 blocks change the output by ~1e-5 relative, so a parity test on
default-init weights passes even when every block kernel is wrong.
"""
from __future__ import annotations

import math
import zlib
from collections import OrderedDict
from typing import Dict, Optional

import numpy as np
import torch


def param_shapes(*, input_fields: int, output_fields: int, patch_size: int, embed_dim: int,
                 num_heads: int, processor_blocks: int, attn_scale: bool = True,
                 feat_scale: bool = True, num_fluid_params: Optional[int] = 9) -> "OrderedDict[str, tuple]":
    """Name -> shape, in reference registration order. num_fluid_params=None -> AViT."""
    E, he = embed_dim, num_heads
    d = E // he
    n_layers = int(math.log2(patch_size))
    s: "OrderedDict[str, tuple]" = OrderedDict()
    cin = input_fields
    for i in range(n_layers):
        last = i == n_layers - 1
        cout = E if (last or n_layers == 1) else E // 4
        s[f"embed.in_proj.{3*i}.weight"] = (cout, cin, 2, 2)
        s[f"embed.in_proj.{3*i+1}.weight"] = (cout,)
        s[f"embed.in_proj.{3*i+1}.bias"] = (cout,)
        cin = cout
    if num_fluid_params is not None:
        s["film_embed.film_net.0.weight"] = (num_fluid_params,)
        s["film_embed.film_net.0.bias"] = (num_fluid_params,)
        s["film_embed.film_net.1.weight"] = (2 * E, num_fluid_params)
        s["film_embed.film_net.1.bias"] = (2 * E,)
    for b in range(processor_blocks):
        for kind in ("temporal", "spatial"):
            p = f"blocks.{b}.{kind}."
            if kind == "temporal":
                s[p + "gamma"] = (E,)
                if attn_scale:
                    s[p + "attn_scale_factor"] = (1, he, 1, 1)
            else:
                s[p + "gamma_att"] = (E,)
                s[p + "gamma_mlp"] = (E,)
                if attn_scale:
                    s[p + "attn_scale_factor_x"] = (1, he, 1, 1)
                    s[p + "attn_scale_factor_y"] = (1, he, 1, 1)
                if feat_scale:
                    s[p + "low_freq_scalar"] = (E,)
                    s[p + "high_freq_scalar"] = (E,)
            for nm in ("norm1", "norm2"):
                s[p + nm + ".weight"] = (E,)
                s[p + nm + ".bias"] = (E,)
            s[p + "input_head.weight"] = (3 * E, E, 1, 1)
            s[p + "input_head.bias"] = (3 * E,)
            s[p + "output_head.weight"] = (E, E, 1, 1)
            s[p + "output_head.bias"] = (E,)
            for nm in ("qnorm", "knorm"):
                s[p + nm + ".weight"] = (d,)
                s[p + nm + ".bias"] = (d,)
            s[p + "rel_pos_bias.relative_attention_bias.weight"] = (32, he)
            if kind == "spatial":
                s[p + "mlp.fc1.weight"] = (4 * E, E)
                s[p + "mlp.fc1.bias"] = (4 * E,)
                s[p + "mlp.fc2.weight"] = (E, 4 * E)
                s[p + "mlp.fc2.bias"] = (E,)
                s[p + "mlp_norm.weight"] = (E,)
                s[p + "mlp_norm.bias"] = (E,)
    cin = E
    for i in range(n_layers):
        last = i == n_layers - 1
        cout = output_fields if (last or n_layers == 1) else E // 4
        s[f"debed.out_proj.{3*i}.weight"] = (cin, cout, 2, 2)
        if not last:
            s[f"debed.out_proj.{3*i+1}.weight"] = (cout,)
            s[f"debed.out_proj.{3*i+1}.bias"] = (cout,)
        cin = cout
    return s


def _draw(name: str, shape: tuple, seed: int) -> np.ndarray:
    rng = np.random.RandomState((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    n = rng.standard_normal(shape)
    leaf = name.split(".")[-1]
    parent = name.split(".")[-2] if "." in name else ""
    if leaf.startswith("gamma"):
        return 0.05 * n
    if leaf in ("low_freq_scalar", "high_freq_scalar"):
        return 0.2 * n
    if leaf.startswith("attn_scale_factor"):
        return 1.0 + 0.3 * n
    if parent == "relative_attention_bias":
        return n
    is_norm = "norm" in parent or name.startswith("film_embed.film_net.0") or \
        (len(shape) == 1 and leaf == "weight")
    if leaf == "weight" and is_norm:
        return 1.0 + 0.1 * n
    if leaf == "bias":
        return 0.1 * n
    # conv / linear / conv-transpose weights: uniform(+-1/sqrt(fan_in)) like torch's default
    if len(shape) == 4 and name.startswith("debed."):
        fan_in = shape[1] * shape[2] * shape[3]      # ConvTranspose2d counts dim 1
    else:
        fan_in = int(np.prod(shape[1:]))
    bound = 1.0 / math.sqrt(fan_in)
    return rng.uniform(-bound, bound, size=shape)


def random_state_dict(shapes: "OrderedDict[str, tuple]", seed: int = 0,
                      dtype: torch.dtype = torch.float32) -> Dict[str, torch.Tensor]:
    return OrderedDict((k, torch.from_numpy(_draw(k, v, seed)).to(dtype)) for k, v in shapes.items())


FLUIDS = np.array([
    # inv_reynolds, cpgas, mugas, rhogas, thcogas, stefan, prandtl, nucWaitTime, wallTemp
    # (order of bubbleformer/data/dataset.py:170-178); FC-72 / R-515B / LN2 -like magnitudes
    [0.0042, 0.83, 0.023, 0.0083, 0.25, 0.50, 8.4, 0.4, 1.00],
    [0.0031, 0.74, 0.071, 0.0290, 0.21, 0.43, 3.6, 0.6, 0.85],
    [0.0025, 0.52, 0.035, 0.0057, 0.06, 0.13, 2.2, 0.2, 0.65],
], dtype=np.float64)


def fluid_params(batch: int, dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """Round-robin FC-72 / R-515B / LN2-like 9-vectors."""
    return torch.from_numpy(FLUIDS[np.arange(batch) % 3]).to(dtype)
