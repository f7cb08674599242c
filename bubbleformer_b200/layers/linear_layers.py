"""GeluMLP / FiLMMLP / SirenMLP (API mirror of upstream bubbleformer/layers/linear_layers.py).

Inside AxialAttentionBlock the MLP runs as two tcgen05 GEMMs with fused bias+GELU / bias epilogues, and
FiLM is fused into the last InstanceNorm of the patch embed; these classes own the parameters.
"""
import torch
import torch.nn as nn

from .. import _lib as L
from .. import ops
from ..autograd import adhoc_w16


class GeluMLP(nn.Module):
    """fc1 -> exact-erf GELU -> fc2 (upstream linear_layers.py:5-25)."""

    def __init__(self, hidden_dim, exp_factor=4.0):
        super().__init__()
        self.fc1 = nn.Linear(hidden_dim, int(hidden_dim * exp_factor))
        self.fc2 = nn.Linear(int(hidden_dim * exp_factor), hidden_dim)
        self.act = nn.GELU()

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """Stand-alone inference path (..., E) -> (..., E); training goes through AxialAttentionBlock."""
        if torch.is_grad_enabled() and (x.requires_grad or self.fc1.weight.requires_grad):
            raise NotImplementedError("stand-alone GeluMLP is forward-only; wrap the call in torch.no_grad() "
                                      "(training runs the MLP inside AxialAttentionBlock)")
        E = x.shape[-1]
        xt = x.reshape(-1, E)
        N = xt.shape[0]
        xb = xt.to(torch.bfloat16).contiguous()
        Hd = self.fc1.weight.shape[0]
        G = torch.empty(N, Hd, dtype=torch.bfloat16, device=x.device)
        ops.gemm(xb, adhoc_w16(self.fc1.weight), N, Hd, E, epilogue=L.EPI_GELU, bias=self.fc1.bias.detach(), out16=G)
        out = torch.empty(N, E, dtype=torch.float32, device=x.device)
        ops.gemm(G, adhoc_w16(self.fc2.weight), N, E, Hd, epilogue=L.EPI_STORE32, bias=self.fc2.bias.detach(), out32=out)
        return out.reshape(x.shape)


class SirenMLP(nn.Module):
    """Unused by every model (upstream linear_layers.py:28-47); plain parameter holder."""

    def __init__(self, hidden_dim, w0=1.0):
        super().__init__()
        self.fc = nn.Linear(hidden_dim, hidden_dim)
        self.w0 = w0

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError("SirenMLP is outside the B200 hot path (no upstream model uses it)")


class FiLMMLP(nn.Module):
    """LayerNorm(param_dim) -> Linear(param_dim, 2E) -> gamma * x + beta (upstream linear_layers.py:49-77).

    In FiLMConditionedAViT the modulation is fused into the embed's last InstanceNorm pass; `gamma_beta`
    exposes the (B, 2E) conditioning vector for that.  `forward` keeps the stand-alone semantics.
    """

    def __init__(self, param_dim, embed_dim):
        super().__init__()
        self.film_net = nn.Sequential(nn.LayerNorm(param_dim), nn.Linear(param_dim, embed_dim * 2))

    def gamma_beta(self, cond: torch.Tensor) -> torch.Tensor:
        return self.film_net(cond.to(torch.float32))

    def forward(self, x: torch.Tensor, cond) -> torch.Tensor:
        gamma, beta = self.gamma_beta(cond).chunk(2, dim=1)
        gamma = gamma.view(-1, 1, x.shape[2], 1, 1)
        beta = beta.view(-1, 1, x.shape[2], 1, 1)
        return gamma * x + beta
