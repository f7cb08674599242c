"""HMLPEmbed / HMLPDebed (API mirror of upstream bubbleformer/layers/patching.py).

The nn.Conv2d / nn.InstanceNorm2d / nn.ConvTranspose2d children are *parameter holders only* (same names,
shapes and default initialisation as upstream so checkpoints load strictly); the computation runs through
bubbleformer_b200.engine: a SIMT 2x2 patch kernel at the fp32 NCHW boundary, implicit-GEMM tcgen05 stages
(space-to-depth expressed as a 4-D TMA box, depth-to-space as a scatter epilogue) and fused IN(+GELU) passes.
"""
import math

import torch
import torch.nn as nn

from .. import engine, ops
from ..autograd import run


def _prefixed(module: nn.Module):
    return {k: v for k, v in module.named_parameters()}


_FILM_KEYS = ("film.film_net.0.weight", "film.film_net.0.bias", "film.film_net.1.weight", "film.film_net.1.bias")


class _EmbedSpec:
    """aux is the (B, 2E) FiLM matrix [gamma | beta] (film=False: produced elsewhere, gets a gradient) or the raw (B, F)
    fluid parameters (film=True: the FiLM MLP of upstream linear_layers.py:58-61 runs inside this Function, so its
    parameter gradients are part of the embed's segment of the flat gradient buffer)."""

    def __init__(self, names, n_layers, T, film=False):
        self.names, self.n_layers, self.T, self.film = names, n_layers, T, film

    def forward(self, x, aux, pd, save):
        gb = aux
        if self.film:
            gb = ops.film_fwd(aux, *(pd[k] for k in _FILM_KEYS))
        X, sv = engine.embed_forward(x, None, pd, self.n_layers, gb, self.T, save)
        if sv is not None and self.film:
            sv["cond"] = aux
        return X, sv

    def backward(self, dout, pd, saved, grads, need_dx):
        dx, dfilm = engine.embed_backward(dout, pd, self.n_layers, saved.get("film_gb"), self.T, saved, grads, need_dx)
        if self.film:
            lw, lb, W, _ = (pd[k] for k in _FILM_KEYS)
            ops.film_bwd(dfilm, saved["cond"], lw, lb, W, *(grads[k] for k in _FILM_KEYS))
            dfilm = None
        return dx, dfilm


class HMLPEmbed(nn.Module):
    """Image to patch embedding with hierarchical 2x2/stride-2 convs (upstream patching.py:6-59)."""

    def __init__(self, patch_size: int = 16, in_channels: int = 3, embed_dim: int = 768):
        super().__init__()
        self.patch_size = patch_size
        num_layers = int(math.log2(patch_size))
        assert (num_layers - math.log2(patch_size)) == 0, "Patch size must be a power of 2"
        self.in_channels = in_channels
        self.embed_dim = embed_dim
        self.num_layers = num_layers
        layers = []
        conv_in = in_channels
        for i in range(num_layers):
            is_last = i == num_layers - 1
            conv_out = embed_dim if (is_last or num_layers == 1) else embed_dim // 4
            layers.append(nn.Conv2d(conv_in, conv_out, kernel_size=2, stride=2, bias=False))
            layers.append(nn.InstanceNorm2d(conv_out, affine=True))
            if not is_last:
                layers.append(nn.GELU())
            conv_in = conv_out
        self.in_proj = nn.Sequential(*layers)

    def tokens(self, x: torch.Tensor, film_in, T: int, film: nn.Module = None) -> torch.Tensor:
        """x: (I, C, H, W) fp32 contiguous -> token-major (I*h*w, E) fp32.

        film_in: None, the (B, 2E) FiLM matrix [gamma | beta], or -- with `film` (a FiLMMLP) -- the raw (B, F) fluid
        parameters; FiLM is applied by the last InstanceNorm pass."""
        pd = _prefixed(self)
        if film is not None:
            pd.update({"film." + k: v for k, v in film.named_parameters()})
        spec = _EmbedSpec(list(pd.keys()), self.num_layers, T, film=film is not None)
        return run(spec, x, film_in, pd)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """(B, C, H, W) -> (B, E, H/p, W/p) like upstream."""
        B, C, H, W = x.shape
        _check_input(x, H, W, self.patch_size)
        X = self.tokens(x.float().contiguous(), None, 1)
        h, w = H // self.patch_size, W // self.patch_size
        return X.view(B, h, w, self.embed_dim).permute(0, 3, 1, 2)


def _check_input(x, H, W, patch):
    if not x.is_cuda:
        raise RuntimeError("bubbleformer_b200 runs on CUDA tensors only (no CPU fallback)")
    if H % patch or W % patch:
        raise ValueError(f"spatial size {H}x{W} is not divisible by the patch size {patch}")


class _DebedSpec:
    def __init__(self, names, n_layers, geom, out_fields):
        self.names, self.n_layers, self.geom, self.out_fields = names, n_layers, geom, out_fields

    def forward(self, x, aux, pd, save):
        return engine.debed_forward(x, self.geom, pd, self.n_layers, self.out_fields, save)

    def backward(self, dout, pd, saved, grads, need_dx):
        return engine.debed_backward(dout, self.geom, pd, self.n_layers, saved, grads), None


class HMLPDebed(nn.Module):
    """Patch to image de-embedding with hierarchical 2x2/stride-2 transposed convs (upstream patching.py:62-115)."""

    def __init__(self, patch_size: int = 16, out_channels: int = 3, embed_dim: int = 768):
        super().__init__()
        self.patch_size = patch_size
        num_layers = int(math.log2(patch_size))
        assert (num_layers - math.log2(patch_size)) == 0, "Patch size must be a power of 2"
        self.out_channels = out_channels
        self.embed_dim = embed_dim
        self.num_layers = num_layers
        layers = []
        conv_in = embed_dim
        for i in range(num_layers):
            is_last = i == num_layers - 1
            conv_out = out_channels if (is_last or num_layers == 1) else embed_dim // 4
            layers.append(nn.ConvTranspose2d(conv_in, conv_out, kernel_size=2, stride=2, bias=False))
            if not is_last:
                layers.append(nn.InstanceNorm2d(conv_out, affine=True))
                layers.append(nn.GELU())
            conv_in = conv_out
        self.out_proj = nn.Sequential(*layers)

    def images(self, X: torch.Tensor, geom: engine.Geom) -> torch.Tensor:
        """token-major (I*h*w, E) fp32 -> (I, C_out, H, W) fp32."""
        pd = _prefixed(self)
        return run(_DebedSpec(list(pd.keys()), self.num_layers, geom, self.out_channels), X, None, pd)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """(B, E, h, w) -> (B, C_out, h*p, w*p) like upstream."""
        B, E, h, w = x.shape
        if not x.is_cuda:
            raise RuntimeError("bubbleformer_b200 runs on CUDA tensors only (no CPU fallback)")
        X = x.float().permute(0, 2, 3, 1).reshape(B * h * w, E).contiguous()
        return self.images(X, engine.Geom(B, 1, h, w))
