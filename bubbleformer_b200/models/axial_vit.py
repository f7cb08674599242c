"""SpaceTimeBlock / AViT / FiLMConditionedAViT (API mirror of upstream bubbleformer/models/axial_vit.py).

Same constructor kwargs, registered names ("avit", "filmavit"), parameter names/shapes (checkpoints load with
strict=True) and (B, T, C, H, W) tensors in and out.  Internally the whole network runs token-major through
the sm_100a kernels of libbubbleformer_b200.so; the only layout changes are at the fp32 NCHW boundaries,
inside the first / last patch kernels.
"""
from typing import Optional

import numpy as np
import torch
import torch.nn as nn

from .. import engine
from ..autograd import WeightBank
from ..layers import AttentionBlock, AxialAttentionBlock, FiLMMLP, HMLPDebed, HMLPEmbed
from ._api import register_model

__all__ = ["AViT"]


class SpaceTimeBlock(nn.Module):
    """Temporal attention followed by axial spatial attention + MLP (upstream axial_vit.py:13-65)."""

    def __init__(self, embed_dim: int = 768, num_heads: int = 12, drop_path: float = 0.0, attn_scale: bool = True,
                 feat_scale: bool = True):
        super().__init__()
        self.temporal = AttentionBlock(embed_dim=embed_dim, num_heads=num_heads, drop_path=drop_path,
                                       attn_scale=attn_scale)
        self.spatial = AxialAttentionBlock(embed_dim=embed_dim, num_heads=num_heads, drop_path=drop_path,
                                           attn_scale=attn_scale, feat_scale=feat_scale)

    def tokens(self, X, geom, w16, masks=None):
        mb, ma, mm = masks if masks is not None else (None, None, None)
        X = self.temporal.tokens(X, geom, w16, mb)
        return self.spatial.tokens(X, geom, w16, ma, mm)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """(B, T, emb, H, W) -> (B, T, emb, H, W) like upstream."""
        B, T, E, h, w = x.shape
        if not x.is_cuda:
            raise RuntimeError("bubbleformer_b200 runs on CUDA tensors only (no CPU fallback)")
        X = x.float().permute(0, 1, 3, 4, 2).reshape(B * T * h * w, E).contiguous()
        from ..autograd import adhoc_w16
        Y = self.tokens(X, engine.Geom(B, T, h, w), adhoc_w16)
        return Y.view(B, T, h, w, E).permute(0, 1, 4, 2, 3)


class _AViTBase(nn.Module):
    """Shared forward of AViT and FiLMConditionedAViT."""

    patch_size: int

    def _init_common(self, input_fields, output_fields, patch_size, embed_dim, num_heads, processor_blocks, drop_path,
                     attn_scale, feat_scale):
        self.patch_size = patch_size
        self.embed_dim = embed_dim
        self.drop_path = drop_path
        self.dp = np.linspace(0, drop_path, processor_blocks)
        self.embed = HMLPEmbed(patch_size=patch_size, in_channels=input_fields, embed_dim=embed_dim)

    def _init_tail(self, output_fields, patch_size, embed_dim, num_heads, processor_blocks, attn_scale, feat_scale):
        self.blocks = nn.ModuleList([
            SpaceTimeBlock(embed_dim=embed_dim, num_heads=num_heads, drop_path=float(self.dp[i]), attn_scale=attn_scale,
                           feat_scale=feat_scale)
            for i in range(processor_blocks)
        ])
        self.debed = HMLPDebed(patch_size=patch_size, embed_dim=embed_dim, out_channels=output_fields)
        self._bank: Optional[WeightBank] = None
        self.drop_masks_override = None      # tests: list of (mask_b, mask_att, mask_mlp) per block

    def _draw_drop_masks(self, B: int, T: int, device):
        """All stochastic-depth masks of one forward in one draw (timm DropPath semantics per block: bernoulli(keep) /
        keep over dim 0 -- per sample for the temporal block, per image for the two spatial residuals)."""
        if not self.training or not any(float(d) > 0.0 for d in self.dp):
            return None
        nb, I = len(self.blocks), B * T
        cache = getattr(self, "_keep_cache", None)
        if cache is None or cache.device != device:       # one host->device copy per model, not per step
            cache = torch.tensor([1.0 - float(d) for d in self.dp], dtype=torch.float32, device=device)[:, None]
            self._keep_cache = cache
        keep = cache
        u = torch.rand(nb, B + 2 * I, device=device)
        m = (u < keep).to(torch.float32) / keep.clamp_min(1e-12)
        mb = m[:, :B].repeat_interleave(T, dim=1).contiguous()        # (nb, I): per-sample mask expanded to images
        out = []
        for i in range(nb):
            if float(self.dp[i]) == 0.0:
                out.append((None, None, None))
            else:
                out.append((mb[i], m[i, B:B + I], m[i, B + I:]))
        return out

    def _run(self, x: torch.Tensor, cond: Optional[torch.Tensor]) -> torch.Tensor:
        if x.dim() != 5:
            raise ValueError(f"expected (B, T, C, H, W), got {tuple(x.shape)}")
        if not x.is_cuda:
            raise RuntimeError("bubbleformer_b200 runs on CUDA tensors only (no CPU fallback): move the model and "
                               "its inputs to a B200")
        B, T, C, H, W = x.shape
        p = self.patch_size
        if H % p or W % p:
            raise ValueError(f"spatial size {H}x{W} is not divisible by the patch size {p}")
        if self._bank is None:
            self._bank = WeightBank(self)
        self._bank.refresh()
        geom = engine.Geom(B, T, H // p, W // p)
        engine.reset_arena()
        xi = x.to(torch.float32).contiguous().view(B * T, C, H, W)
        X = self.embed.tokens(xi, cond, T, film=self.film_embed if cond is not None else None)
        drawn = self._draw_drop_masks(B, T, x.device) if self.drop_masks_override is None else None
        for i, blk in enumerate(self.blocks):
            masks = self.drop_masks_override[i] if self.drop_masks_override is not None else (drawn[i] if drawn else None)
            X = blk.tokens(X, geom, self._bank.w16, masks)
        out = self.debed.images(X, geom)
        return out.view(B, T, -1, H, W)


@register_model("avit")
class AViT(_AViTBase):
    """Factored space-time ViT (upstream axial_vit.py:68-151)."""

    def __init__(self, input_fields: int = 3, output_fields: int = 3, time_window: int = 12, patch_size: int = 16,
                 embed_dim: int = 768, num_heads: int = 12, processor_blocks: int = 12, drop_path: int = 0.2,
                 attn_scale: bool = True, feat_scale: bool = True):
        super().__init__()
        self._init_common(input_fields, output_fields, patch_size, embed_dim, num_heads, processor_blocks, drop_path,
                          attn_scale, feat_scale)
        self._init_tail(output_fields, patch_size, embed_dim, num_heads, processor_blocks, attn_scale, feat_scale)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """(B, T, C, H, W) -> (B, T, C_out, H, W)."""
        return self._run(x, None)


@register_model("filmavit")
class FiLMConditionedAViT(_AViTBase):
    """AViT with FiLM conditioning of the patch embedding on the fluid parameters (upstream axial_vit.py:154-242)."""

    def __init__(self, input_fields: int = 3, output_fields: int = 3, time_window: int = 12, patch_size: int = 16,
                 embed_dim: int = 768, num_heads: int = 12, processor_blocks: int = 12, drop_path: int = 0.2,
                 attn_scale: bool = True, feat_scale: bool = True, num_fluid_params: int = 8):
        super().__init__()
        self._init_common(input_fields, output_fields, patch_size, embed_dim, num_heads, processor_blocks, drop_path,
                          attn_scale, feat_scale)
        self.film_embed = FiLMMLP(num_fluid_params, embed_dim)
        self._init_tail(output_fields, patch_size, embed_dim, num_heads, processor_blocks, attn_scale, feat_scale)

    def forward(self, x: torch.Tensor, fluid_params: torch.Tensor) -> torch.Tensor:
        """x: (B, T, C, H, W); fluid_params: (B, num_fluid_params) -> (B, T, C_out, H, W)."""
        if fluid_params.shape[0] != x.shape[0]:
            raise ValueError("fluid_params must have one row per batch sample")
        return self._run(x, fluid_params.to(device=x.device, dtype=torch.float32).contiguous())
