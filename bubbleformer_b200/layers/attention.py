"""AttentionBlock / AxialAttentionBlock (API mirror of upstream bubbleformer/layers/attention.py).

Children (InstanceNorm2d, Conv2d 1x1, LayerNorm, RelativePositionBias, GeluMLP) hold the parameters under
the upstream names; the forward/backward passes run through bubbleformer_b200.engine on token-major data.
"""
from typing import Optional

import torch
import torch.nn as nn

from .. import engine
from ..autograd import adhoc_w16, run
from .linear_layers import GeluMLP
from .positional_encoding import ContinuousPositionBias1D, RelativePositionBias


def _drop_mask(n: int, p: float, training: bool, device) -> Optional[torch.Tensor]:
    """timm DropPath semantics: bernoulli(keep)/keep over dim 0; identity in eval mode or for p = 0."""
    if p == 0.0 or not training:
        return None
    keep = 1.0 - p
    m = torch.empty(n, dtype=torch.float32, device=device).bernoulli_(keep)
    if keep > 0.0:
        m.div_(keep)
    return m


class _TemporalSpec:
    def __init__(self, names, geom, heads, attn_scale, mask_img, w16):
        self.names, self.geom, self.heads, self.attn_scale, self.mask_img, self.w16 = names, geom, heads, attn_scale, mask_img, w16

    def forward(self, x, aux, pd, save):
        return engine.temporal_forward(x, self.geom, pd, lambda n: self.w16(pd[n]), self.heads, self.attn_scale,
                                       self.mask_img, save)

    def backward(self, dout, pd, saved, grads, need_dx):
        dx = engine.temporal_backward(dout, self.geom, pd, lambda n: self.w16(pd[n]), self.heads, self.attn_scale,
                                      self.mask_img, saved, grads)
        return dx, None


class AttentionBlock(nn.Module):
    """Self-attention across the time axis of (B, n, emb, H, W) tensors (upstream attention.py:10-124)."""

    def __init__(self, embed_dim: int = 768, num_heads: int = 12, drop_path: float = 0,
                 layer_scale_init_value: float = 1e-6, bias_type: str = "rel", attn_scale: bool = True):
        super().__init__()
        if bias_type != "rel" or not layer_scale_init_value > 0:
            raise NotImplementedError("only bias_type='rel' with a positive layer scale (what every upstream model "
                                      "config uses) is implemented on the B200 path")
        self.embed_dim = embed_dim
        self.num_heads = num_heads
        self.attn_scale = attn_scale
        self.drop_prob = float(drop_path)
        self.norm1 = nn.InstanceNorm2d(embed_dim, affine=True)
        self.norm2 = nn.InstanceNorm2d(embed_dim, affine=True)
        self.gamma = nn.Parameter(layer_scale_init_value * torch.ones((embed_dim)), requires_grad=True)
        self.input_head = nn.Conv2d(embed_dim, 3 * embed_dim, 1)
        self.output_head = nn.Conv2d(embed_dim, embed_dim, 1)
        self.qnorm = nn.LayerNorm(embed_dim // num_heads)
        self.knorm = nn.LayerNorm(embed_dim // num_heads)
        if attn_scale:
            self.attn_scale_factor = nn.Parameter(torch.ones((1, num_heads, 1, 1)), requires_grad=True)
        self.rel_pos_bias = RelativePositionBias(n_heads=num_heads)
        self.drop_path = nn.Identity()          # stochastic depth is applied inside the fused epilogue

    def tokens(self, X: torch.Tensor, geom: engine.Geom, w16=adhoc_w16, mask_b: Optional[torch.Tensor] = None):
        """token-major (B*T*P, E) fp32 -> same.  mask_b: optional (B,) drop-path factors (else drawn)."""
        if mask_b is None:
            mask_b = _drop_mask(geom.B, self.drop_prob, self.training, X.device)
        if mask_b is None:
            mask_img = None
        elif mask_b.numel() == geom.I and geom.T > 1:          # already expanded to one factor per (b, t) image
            mask_img = mask_b.to(torch.float32).contiguous()
        else:
            mask_img = mask_b.to(torch.float32).repeat_interleave(geom.T).contiguous()
        pd = dict(self.named_parameters())
        spec = _TemporalSpec(list(pd.keys()), geom, self.num_heads, self.attn_scale, mask_img, w16)
        return run(spec, X, None, pd)

    def forward(self, x):
        """(B, N, emb, H, W) -> (B, N, emb, H, W) like upstream."""
        B, n, E, h, w = x.shape
        if not x.is_cuda:
            raise RuntimeError("bubbleformer_b200 runs on CUDA tensors only (no CPU fallback)")
        X = x.float().permute(0, 1, 3, 4, 2).reshape(B * n * h * w, E).contiguous()
        Y = self.tokens(X, engine.Geom(B, n, h, w))
        return Y.view(B, n, h, w, E).permute(0, 1, 4, 2, 3)


class _SpatialSpec:
    def __init__(self, names, geom, heads, attn_scale, feat_scale, mask_att, mask_mlp, w16):
        self.names, self.geom, self.heads = names, geom, heads
        self.attn_scale, self.feat_scale, self.mask_att, self.mask_mlp, self.w16 = attn_scale, feat_scale, mask_att, mask_mlp, w16

    def forward(self, x, aux, pd, save):
        return engine.spatial_forward(x, self.geom, pd, lambda n: self.w16(pd[n]), self.heads, self.attn_scale,
                                      self.feat_scale, self.mask_att, self.mask_mlp, save)

    def backward(self, dout, pd, saved, grads, need_dx):
        dx = engine.spatial_backward(dout, self.geom, pd, lambda n: self.w16(pd[n]), self.heads, self.attn_scale,
                                     self.feat_scale, self.mask_att, self.mask_mlp, saved, grads)
        return dx, None


class AxialAttentionBlock(nn.Module):
    """Axial (x then y, averaged) attention + MLP on (B, emb, H, W) tensors (upstream attention.py:127-319)."""

    def __init__(self, embed_dim=768, num_heads=12, drop_path=0, layer_scale_init_value=1e-6, bias_type="rel",
                 attn_scale=True, feat_scale=True):
        super().__init__()
        if bias_type != "rel" or not layer_scale_init_value > 0:
            raise NotImplementedError("only bias_type='rel' with a positive layer scale (what every upstream model "
                                      "config uses) is implemented on the B200 path")
        self.embed_dim = embed_dim
        self.num_heads = num_heads
        self.attn_scale = attn_scale
        self.feat_scale = feat_scale
        self.drop_prob = float(drop_path)
        self.norm1 = nn.InstanceNorm2d(embed_dim, affine=True)
        self.norm2 = nn.InstanceNorm2d(embed_dim, affine=True)
        self.gamma_att = nn.Parameter(layer_scale_init_value * torch.ones((embed_dim)), requires_grad=True)
        self.gamma_mlp = nn.Parameter(layer_scale_init_value * torch.ones((embed_dim)), requires_grad=True)
        self.input_head = nn.Conv2d(embed_dim, 3 * embed_dim, 1)
        self.output_head = nn.Conv2d(embed_dim, embed_dim, 1)
        self.qnorm = nn.LayerNorm(embed_dim // num_heads)
        self.knorm = nn.LayerNorm(embed_dim // num_heads)
        self.rel_pos_bias = RelativePositionBias(n_heads=num_heads)
        if attn_scale:
            self.attn_scale_factor_x = nn.Parameter(torch.ones((1, num_heads, 1, 1)), requires_grad=True)
            self.attn_scale_factor_y = nn.Parameter(torch.ones((1, num_heads, 1, 1)), requires_grad=True)
        if feat_scale:
            self.low_freq_scalar = nn.Parameter(torch.zeros(embed_dim), requires_grad=True)
            self.high_freq_scalar = nn.Parameter(torch.zeros(embed_dim), requires_grad=True)
        self.drop_path = nn.Identity()          # stochastic depth is applied inside the fused epilogues
        self.mlp = GeluMLP(embed_dim)
        self.mlp_norm = nn.InstanceNorm2d(embed_dim, affine=True)

    def tokens(self, X: torch.Tensor, geom: engine.Geom, w16=adhoc_w16, mask_att=None, mask_mlp=None):
        """token-major (I*P, E) fp32 -> same.  masks: optional (I,) drop-path factors (else drawn)."""
        if mask_att is None:
            mask_att = _drop_mask(geom.I, self.drop_prob, self.training, X.device)
        if mask_mlp is None:
            mask_mlp = _drop_mask(geom.I, self.drop_prob, self.training, X.device)
        ma = mask_att.to(torch.float32).contiguous() if mask_att is not None else None
        mm = mask_mlp.to(torch.float32).contiguous() if mask_mlp is not None else None
        pd = dict(self.named_parameters())
        spec = _SpatialSpec(list(pd.keys()), geom, self.num_heads, self.attn_scale, self.feat_scale, ma, mm, w16)
        return run(spec, X, None, pd)

    def forward(self, x):
        """(B, emb, H, W) -> (B, emb, H, W) like upstream."""
        B, E, h, w = x.shape
        if not x.is_cuda:
            raise RuntimeError("bubbleformer_b200 runs on CUDA tensors only (no CPU fallback)")
        X = x.float().permute(0, 2, 3, 1).reshape(B * h * w, E).contiguous()
        Y = self.tokens(X, engine.Geom(B, 1, h, w))
        return Y.view(B, h, w, E).permute(0, 3, 1, 2)
