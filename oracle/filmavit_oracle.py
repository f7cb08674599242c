"""CPU oracle for the FiLMAViT / AViT hot path -- TEST INFRASTRUCTURE ONLY.

This file is a *restatement* (not a copy) of the reference algorithm in plain
torch tensor algebra on the CPU, written token-major / channels-last the way the
CUDA kernels see the data.  It is the checker for the CUDA path: only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import it.  The product (`bubbleformer_b200`) never does.

Pinning status: the reference ships NO golden vectors for this path (its tests
are shape-only).  The oracle is therefore pinned against
outputs of the reference itself, executed in the build container by
`oracle/make_golden.py` (fixtures in `tests/golden/`, checked by
`tests/test_oracle_golden.py`), forward and all parameter gradients, fp64.

Reference lines restated (relative to the upstream repository root):
  embed            bubbleformer/layers/patching.py:30-58
  debed            bubbleformer/layers/patching.py:86-115
  film             bubbleformer/layers/linear_layers.py:56-77
  relpos_bias      bubbleformer/layers/positional_encoding.py:76-162
  attention        bubbleformer/layers/attention.py:84-101 (hf-scaled softmax)
                   bubbleformer/layers/attention.py:105-117 (attn_scale=False)
  temporal_block   bubbleformer/layers/attention.py:66-124
  spatial_block    bubbleformer/layers/attention.py:199-319
  mlp              bubbleformer/layers/linear_layers.py:14-25
  model forward    bubbleformer/models/axial_vit.py:127-151, 217-242
  drop path        timm.layers.DropPath (third party, unpinned in
                   env/requirements.txt:10): mask ~ Bernoulli(keep)/keep over dim 0

Data layout used here: an *image* is one (b, t) frame; activations are
(I, h, w, C) channels-last, i.e. rows of a (tokens, C) matrix.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

Tensor = torch.Tensor
EPS = 1e-5  # nn.InstanceNorm2d / nn.LayerNorm default


# ----------------------------------------------------------------------------
# primitives
# ----------------------------------------------------------------------------
def instance_norm(x: Tensor, w: Tensor, b: Tensor) -> Tensor:
    """x: (I, h, w, C). Per image and channel over the h*w grid, biased variance."""
    mu = x.mean(dim=(1, 2), keepdim=True)
    var = ((x - mu) ** 2).mean(dim=(1, 2), keepdim=True)
    return (x - mu) * torch.rsqrt(var + EPS) * w + b


def layer_norm_last(x: Tensor, w: Tensor, b: Tensor) -> Tensor:
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) * torch.rsqrt(var + EPS) * w + b


def gelu_erf(x: Tensor) -> Tensor:
    return 0.5 * x * (1.0 + torch.erf(x * (1.0 / math.sqrt(2.0))))


def relpos_bucket(rel: int, num_buckets: int = 32, max_distance: int = 32) -> int:
    """T5 bidirectional bucket for rel = key_pos - query_pos.

    positional_encoding.py:101-129.  NB the static default max_distance=32 is what
    actually runs (compute_bias never forwards self.max_distance).
    The large-distance branch is evaluated in float32 like the reference
    (torch.log on a float32 tensor) so bucket edges agree bit for bit.
    """
    half = num_buckets // 2
    ret = half if rel > 0 else 0          # n = -rel; (n < 0) -> + half
    n = abs(rel)
    max_exact = half // 2
    if n < max_exact:
        return ret + n
    v = torch.log(torch.tensor(float(n), dtype=torch.float32) / max_exact) \
        / math.log(max_distance / max_exact) * (half - max_exact)
    large = max_exact + int(v.to(torch.long))
    return ret + min(large, half - 1)


def relpos_bucket_table(L: int) -> Tensor:
    """(L, L) long: bucket[i, j] for query i, key j."""
    t = torch.empty(L, L, dtype=torch.long)
    for i in range(L):
        for j in range(L):
            t[i, j] = relpos_bucket(j - i)
    return t


def relpos_bias(emb: Tensor, L: int) -> Tensor:
    """emb: (32, heads) -> (heads, L, L)."""
    return emb[relpos_bucket_table(L)].permute(2, 0, 1)


def attention_1d(qkv: Tensor, heads: int, qn_w, qn_b, kn_w, kn_b,
                 bias_emb: Tensor, scale_factor: Optional[Tensor]) -> Tensor:
    """qkv: (S, L, 3E) -- S independent sequences of L tokens.

    Column layout of the 3E axis: col = head*3d + j with j in [0,d) -> q,
    [d,2d) -> k, [2d,3d) -> v  (attention.py:80-81, 212-213).
    Returns (S, L, E) with col = head*d + j.
    scale_factor: (heads,) high-frequency scale, or None for attn_scale=False.
    """
    S, L, E3 = qkv.shape
    E = E3 // 3
    d = E // heads
    x = qkv.reshape(S, L, heads, 3, d).permute(0, 2, 3, 1, 4)      # (S, he, 3, L, d)
    q, k, v = x[:, :, 0], x[:, :, 1], x[:, :, 2]
    q = layer_norm_last(q, qn_w, qn_b)
    k = layer_norm_last(k, kn_w, kn_b)
    s = q @ k.transpose(-1, -2) * (d ** -0.5) + relpos_bias(bias_emb, L).to(q.dtype)
    p = torch.softmax(s, dim=-1)
    if scale_factor is not None:
        # the reference builds attn_low = torch.ones(L, L) / L in float32 whatever the model
        # dtype (attention.py:95), so 1/L carries float32 rounding even in a float64 run
        low = float(torch.ones((), dtype=torch.float32) / L)
        p = low + (p - low) * scale_factor.reshape(1, heads, 1, 1)
    o = p @ v                                                       # (S, he, L, d)
    return o.permute(0, 2, 1, 3).reshape(S, L, E)


def drop_mask(mask: Optional[Tensor], n: int, like: Tensor) -> Tensor:
    """mask: None (eval / p=0) or a (n,) tensor of already scaled keep factors."""
    if mask is None:
        return torch.ones(n, dtype=like.dtype)
    return mask.to(like.dtype)


# ----------------------------------------------------------------------------
# stem / head
# ----------------------------------------------------------------------------
def space_to_depth(x: Tensor) -> Tensor:
    """(I, H, W, C) -> (I, H/2, W/2, C*4) with inner order (c, ky, kx)."""
    I, H, W, C = x.shape
    x = x.reshape(I, H // 2, 2, W // 2, 2, C).permute(0, 1, 3, 5, 2, 4)
    return x.reshape(I, H // 2, W // 2, C * 4)


def depth_to_space(y: Tensor, cout: int) -> Tensor:
    """(I, H, W, cout*4) with inner order (co, ky, kx) -> (I, 2H, 2W, cout)."""
    I, H, W, _ = y.shape
    y = y.reshape(I, H, W, cout, 2, 2).permute(0, 1, 4, 2, 5, 3)
    return y.reshape(I, 2 * H, 2 * W, cout)


def embed(x: Tensor, sd: Dict[str, Tensor], prefix: str, patch: int) -> Tensor:
    """x: (I, H, W, Cin) -> (I, H/p, W/p, E).  patching.py:30-58."""
    n_layers = int(math.log2(patch))
    for i in range(n_layers):
        w = sd[f"{prefix}in_proj.{3 * i}.weight"]                    # (Cout, Cin, 2, 2)
        x = space_to_depth(x) @ w.reshape(w.shape[0], -1).t()
        x = instance_norm(x, sd[f"{prefix}in_proj.{3 * i + 1}.weight"],
                          sd[f"{prefix}in_proj.{3 * i + 1}.bias"])
        if i != n_layers - 1:
            x = gelu_erf(x)
    return x


def debed(x: Tensor, sd: Dict[str, Tensor], prefix: str, patch: int) -> Tensor:
    """x: (I, h, w, E) -> (I, h*p, w*p, Cout).  patching.py:86-115."""
    n_layers = int(math.log2(patch))
    for i in range(n_layers):
        w = sd[f"{prefix}out_proj.{3 * i}.weight"]                   # (Cin, Cout, 2, 2)
        cout = w.shape[1]
        x = depth_to_space(x @ w.reshape(w.shape[0], -1), cout)
        if i != n_layers - 1:
            x = instance_norm(x, sd[f"{prefix}out_proj.{3 * i + 1}.weight"],
                              sd[f"{prefix}out_proj.{3 * i + 1}.bias"])
            x = gelu_erf(x)
    return x


def film(x: Tensor, cond: Tensor, sd: Dict[str, Tensor], prefix: str, T: int) -> Tensor:
    """x: (B*T, h, w, E); cond: (B, F).  linear_layers.py:56-77."""
    E = x.shape[-1]
    c = layer_norm_last(cond, sd[f"{prefix}film_net.0.weight"], sd[f"{prefix}film_net.0.bias"])
    gb = c @ sd[f"{prefix}film_net.1.weight"].t() + sd[f"{prefix}film_net.1.bias"]
    gamma = gb[:, :E].repeat_interleave(T, dim=0)[:, None, None, :]
    beta = gb[:, E:].repeat_interleave(T, dim=0)[:, None, None, :]
    return gamma * x + beta


# ----------------------------------------------------------------------------
# blocks
# ----------------------------------------------------------------------------
def conv1x1(x: Tensor, w: Tensor, b: Tensor) -> Tensor:
    return x @ w.reshape(w.shape[0], w.shape[1]).t() + b


def temporal_block(x: Tensor, sd, p: str, B: int, T: int, heads: int, attn_scale: bool,
                   mask_b: Optional[Tensor] = None) -> Tensor:
    """x: (B*T, h, w, E).  attention.py:66-124.  mask_b: (B,) drop-path factors."""
    I, h, w, E = x.shape
    y = instance_norm(x, sd[p + "norm1.weight"], sd[p + "norm1.bias"])
    qkv = conv1x1(y, sd[p + "input_head.weight"], sd[p + "input_head.bias"])   # (I,h,w,3E)
    seq = qkv.reshape(B, T, h * w, 3 * E).permute(0, 2, 1, 3).reshape(B * h * w, T, 3 * E)
    sf = sd[p + "attn_scale_factor"].reshape(-1) if attn_scale else None
    o = attention_1d(seq, heads, sd[p + "qnorm.weight"], sd[p + "qnorm.bias"],
                     sd[p + "knorm.weight"], sd[p + "knorm.bias"],
                     sd[p + "rel_pos_bias.relative_attention_bias.weight"], sf)
    o = o.reshape(B, h * w, T, E).permute(0, 2, 1, 3).reshape(I, h, w, E)
    o = instance_norm(o, sd[p + "norm2.weight"], sd[p + "norm2.bias"])
    z = conv1x1(o, sd[p + "output_head.weight"], sd[p + "output_head.bias"])
    m = drop_mask(mask_b, B, x).repeat_interleave(T)[:, None, None, None]
    return x + m * (z * sd[p + "gamma"])


def spatial_block(x: Tensor, sd, p: str, heads: int, attn_scale: bool, feat_scale: bool,
                  mask_att: Optional[Tensor] = None, mask_mlp: Optional[Tensor] = None) -> Tensor:
    """x: (I, h, w, E).  attention.py:199-319.  masks: (I,) drop-path factors."""
    I, h, w, E = x.shape
    y = instance_norm(x, sd[p + "norm1.weight"], sd[p + "norm1.bias"])
    qkv = conv1x1(y, sd[p + "input_head.weight"], sd[p + "input_head.bias"])   # (I,h,w,3E)
    emb = sd[p + "rel_pos_bias.relative_attention_bias.weight"]
    ln = (sd[p + "qnorm.weight"], sd[p + "qnorm.bias"], sd[p + "knorm.weight"], sd[p + "knorm.bias"])
    sx = sd[p + "attn_scale_factor_x"].reshape(-1) if attn_scale else None
    sy = sd[p + "attn_scale_factor_y"].reshape(-1) if attn_scale else None
    # x direction: sequences run along w (one per image row)
    ox = attention_1d(qkv.reshape(I * h, w, 3 * E), heads, *ln, emb, sx).reshape(I, h, w, E)
    # y direction: sequences run along h (one per image column)
    oy = attention_1d(qkv.permute(0, 2, 1, 3).reshape(I * w, h, 3 * E), heads, *ln, emb, sy)
    oy = oy.reshape(I, w, h, E).permute(0, 2, 1, 3)
    o = (ox + oy) / 2
    o = instance_norm(o, sd[p + "norm2.weight"], sd[p + "norm2.bias"])
    z = conv1x1(o, sd[p + "output_head.weight"], sd[p + "output_head.bias"])
    if feat_scale:
        z_low = z.mean(dim=(1, 2), keepdim=True)
        z_high = z - z_low
        z = z + z_low * sd[p + "low_freq_scalar"] + z_high * sd[p + "high_freq_scalar"]
    ma = drop_mask(mask_att, I, x)[:, None, None, None]
    x = x + ma * (z * sd[p + "gamma_att"])
    hdn = gelu_erf(x @ sd[p + "mlp.fc1.weight"].t() + sd[p + "mlp.fc1.bias"])
    y2 = hdn @ sd[p + "mlp.fc2.weight"].t() + sd[p + "mlp.fc2.bias"]
    y2 = instance_norm(y2, sd[p + "mlp_norm.weight"], sd[p + "mlp_norm.bias"])
    mm = drop_mask(mask_mlp, I, x)[:, None, None, None]
    return x + mm * (sd[p + "gamma_mlp"] * y2)


# ----------------------------------------------------------------------------
# whole models
# ----------------------------------------------------------------------------
def num_blocks(sd) -> int:
    return 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("blocks."))


def forward(sd: Dict[str, Tensor], x: Tensor, fluid_params: Optional[Tensor], *,
            patch_size: int, num_heads: int, attn_scale: bool = True, feat_scale: bool = True,
            drop_masks: Optional[list] = None) -> Tensor:
    """FiLMConditionedAViT.forward (fluid_params given) or AViT.forward (None).

    x: (B, T, C, H, W) -> (B, T, C_out, H, W).  drop_masks: optional list, one
    (mask_b, mask_att, mask_mlp) triple per block, of already scaled factors.
    """
    B, T, C, H, W = x.shape
    xi = x.reshape(B * T, C, H, W).permute(0, 2, 3, 1)              # images, channels-last
    tok = embed(xi, sd, "embed.", patch_size)
    if fluid_params is not None:
        tok = film(tok, fluid_params, sd, "film_embed.", T)
    for i in range(num_blocks(sd)):
        mb, ma, mm = drop_masks[i] if drop_masks is not None else (None, None, None)
        tok = temporal_block(tok, sd, f"blocks.{i}.temporal.", B, T, num_heads, attn_scale, mb)
        tok = spatial_block(tok, sd, f"blocks.{i}.spatial.", num_heads, attn_scale, feat_scale, ma, mm)
    out = debed(tok, sd, "debed.", patch_size)                      # (I, H, W, Cout)
    return out.permute(0, 3, 1, 2).reshape(B, T, -1, H, W)


def rel_l2_loss(pred: Tensor, tgt: Tensor) -> Tensor:
    """LpLoss(d=2, p=2, reduce_dims=[0,1,2], reductions=[mean,mean,sum]).

    utils/losses.py:67-94 as configured at bubbleformer/modules.py:50.
    """
    diff = (pred - tgt).flatten(-2).norm(dim=-1)
    ynorm = tgt.flatten(-2).norm(dim=-1)
    return (diff / ynorm).mean(0).mean(0).sum()


def rel_l2(a: Tensor, b: Tensor) -> float:
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).norm() / b.norm().clamp_min(1e-300))
