"""GeluMLP / FiLMMLP / SirenMLP (API mirror of upstream bubbleformer/layers/linear_layers.py).

Inside AxialAttentionBlock the MLP runs as two tcgen05 GEMMs with fused bias+GELU / bias epilogues, and
FiLM is fused into the last InstanceNorm of the patch embed; these classes own the parameters.
"""
import torch
import torch.nn as nn

from .. import _lib as L
from .. import ops
from ..autograd import adhoc_w16


class _GeluMLPFn(torch.autograd.Function):
    """Stand-alone fc1 -> GELU -> fc2 with its backward, on the same GEMM epilogues the axial block uses (bias + GELU
    storing the pre-activation, dGELU with the fc1-bias column sums, split-K weight gradients)."""

    @staticmethod
    def forward(ctx, x, W1, b1, W2, b2):
        from ..engine import pick_split
        N, E = x.shape
        Hd = W1.shape[0]
        xb = torch.empty(N, E, dtype=torch.bfloat16, device=x.device)
        ops.cast16(x.detach().float().contiguous().reshape(-1), xb.reshape(-1))
        w1, w2 = adhoc_w16(W1), adhoc_w16(W2)
        G = torch.empty(N, Hd, dtype=torch.bfloat16, device=x.device)
        need = any(ctx.needs_input_grad)
        Hpre = torch.empty(N, Hd, dtype=torch.bfloat16, device=x.device) if need else None
        ops.gemm(xb, w1, N, Hd, E, epilogue=L.EPI_GELU_D, bias=b1.detach(), out16=G, out16b=Hpre)
        out = torch.empty(N, E, dtype=torch.float32, device=x.device)
        ops.gemm(G, w2, N, E, Hd, epilogue=L.EPI_STORE32, bias=b2.detach(), out32=out)
        if need:
            ctx.save_for_backward(xb, G, Hpre, w1, w2)
            ctx.pick_split = pick_split
        return out

    @staticmethod
    def backward(ctx, dY):
        xb, G, Hpre, w1, w2 = ctx.saved_tensors
        N, E = xb.shape
        Hd = w1.shape[0]
        dev = dY.device
        dY16 = torch.empty(N, E, dtype=torch.bfloat16, device=dev)
        ops.cast16(dY.float().contiguous().reshape(-1), dY16.reshape(-1))
        dW1 = torch.zeros(Hd, E, dtype=torch.float32, device=dev)
        db1 = torch.zeros(Hd, dtype=torch.float32, device=dev)
        dW2 = torch.zeros(E, Hd, dtype=torch.float32, device=dev)
        db2 = torch.zeros(E, dtype=torch.float32, device=dev)
        dH = torch.empty(N, Hd, dtype=torch.bfloat16, device=dev)
        ops.gemm(dY16, w2, N, Hd, E, epilogue=L.EPI_DMUL, b_mode=L.B_KN, aux16=Hpre, out16=dH, colsum_out=db1)
        ops.gemm(dY16, G, E, Hd, N, epilogue=L.EPI_ATOMIC32, a_mode=L.A_KM, b_mode=L.B_KN,
                 split_k=ctx.pick_split(N, E, Hd), out32=dW2)
        ops.colsum16(dY16, db2)
        dX = torch.empty(N, E, dtype=torch.float32, device=dev)
        ops.gemm(dH, w1, N, E, Hd, epilogue=L.EPI_STORE32, b_mode=L.B_KN, out32=dX)
        ops.gemm(dH, xb, Hd, E, N, epilogue=L.EPI_ATOMIC32, a_mode=L.A_KM, b_mode=L.B_KN,
                 split_k=ctx.pick_split(N, Hd, E), out32=dW1)
        return dX, dW1, db1, dW2, db2


class GeluMLP(nn.Module):
    """fc1 -> GELU -> fc2 (upstream linear_layers.py:5-25; GELU form: see include/bubbleformer_b200.h, "GELU")."""

    def __init__(self, hidden_dim, exp_factor=4.0):
        super().__init__()
        self.fc1 = nn.Linear(hidden_dim, int(hidden_dim * exp_factor))
        self.fc2 = nn.Linear(int(hidden_dim * exp_factor), hidden_dim)
        self.act = nn.GELU()

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """Stand-alone path (..., E) -> (..., E), forward and backward (inside AxialAttentionBlock the same GEMMs run
        as part of the block's own Function).  The token count must be a multiple of 8 (16-byte rows of the operand
        that the weight-gradient GEMM reads transposed)."""
        if not x.is_cuda:
            raise RuntimeError("bubbleformer_b200 runs on CUDA tensors only (no CPU fallback)")
        E = x.shape[-1]
        xt = x.reshape(-1, E)
        out = _GeluMLPFn.apply(xt, self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias)
        return out.reshape(x.shape)


class SirenMLP(nn.Module):
    """Unused by every model (upstream linear_layers.py:28-47); plain parameter holder."""

    def __init__(self, hidden_dim, w0=1.0):
        super().__init__()
        self.fc = nn.Linear(hidden_dim, hidden_dim)
        self.w0 = w0

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError("SirenMLP is outside the B200 hot path (no upstream model uses it)")


class _FilmFn(torch.autograd.Function):
    """LayerNorm(F) -> Linear(F, 2E) on the device kernels (stand-alone use; inside FiLMConditionedAViT the same two
    kernels run within the patch-embed Function so the parameter gradients land in the flat gradient buffer)."""

    @staticmethod
    def forward(ctx, cond, ln_w, ln_b, W, b):
        ctx.save_for_backward(cond, ln_w, ln_b, W)
        return ops.film_fwd(cond, ln_w.detach(), ln_b.detach(), W.detach(), b.detach())

    @staticmethod
    def backward(ctx, dgb):
        cond, ln_w, ln_b, W = ctx.saved_tensors
        flat = torch.zeros(2 * ln_w.numel() + W.numel() + W.shape[0] + 24, dtype=torch.float32, device=dgb.device)
        F_, E2 = ln_w.numel(), W.shape[0]
        o = [0, (F_ + 7) // 8 * 8, 2 * ((F_ + 7) // 8 * 8)]
        d_lw, d_lb = flat[o[0]:o[0] + F_], flat[o[1]:o[1] + F_]
        d_W = flat[o[2]:o[2] + E2 * F_].view(E2, F_)
        ob = o[2] + (E2 * F_ + 7) // 8 * 8
        d_b = flat[ob:ob + E2]
        ops.film_bwd(dgb.contiguous(), cond, ln_w.detach(), ln_b.detach(), W.detach(), d_lw, d_lb, d_W, d_b)
        return None, d_lw, d_lb, d_W, d_b


class FiLMMLP(nn.Module):
    """LayerNorm(param_dim) -> Linear(param_dim, 2E) -> gamma * x + beta (upstream linear_layers.py:49-77).

    In FiLMConditionedAViT the modulation is fused into the embed's last InstanceNorm pass; `gamma_beta`
    exposes the (B, 2E) conditioning vector for that.  `forward` keeps the stand-alone semantics.
    """

    def __init__(self, param_dim, embed_dim):
        super().__init__()
        self.film_net = nn.Sequential(nn.LayerNorm(param_dim), nn.Linear(param_dim, embed_dim * 2))

    def gamma_beta(self, cond: torch.Tensor) -> torch.Tensor:
        """(B, F) fluid parameters -> (B, 2E) = [gamma | beta] through bf_film_fwd / bf_film_bwd."""
        ln, lin = self.film_net[0], self.film_net[1]
        return _FilmFn.apply(cond.to(torch.float32).contiguous(), ln.weight, ln.bias, lin.weight, lin.bias)

    def forward(self, x: torch.Tensor, cond) -> torch.Tensor:
        gamma, beta = self.gamma_beta(cond).chunk(2, dim=1)
        gamma = gamma.view(-1, 1, x.shape[2], 1, 1)
        beta = beta.view(-1, 1, x.shape[2], 1, 1)
        return gamma * x + beta
