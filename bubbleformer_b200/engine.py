"""Forward / backward orchestration of the FiLMAViT hot path over the C-ABI kernels.

Everything here is host-side sequencing: which kernel runs on which buffer.  All device work is done by
libbubbleformer_b200.so (bubbleformer_b200.ops); torch supplies memory, streams and autograd plumbing.

Data layout: activations are token-major.  One *image* is a (b, t) frame of P = h*w tokens; the
residual stream X is an fp32 (I*P, E) matrix, GEMM operands are bf16 copies (fp16 inside the patch
embed / unembed, whose errors are amplified by the InstanceNorms that follow them; their gradients are
bf16 again so they cannot underflow).

Reference behaviour restated here (upstream paths):
  temporal block   bubbleformer/layers/attention.py:66-124
  spatial block    bubbleformer/layers/attention.py:199-319
  embed / debed    bubbleformer/layers/patching.py:30-58, 86-115
  FiLM             bubbleformer/layers/linear_layers.py:56-77
"""
from __future__ import annotations

import math
import weakref
from dataclasses import dataclass
from typing import Dict, List, Optional

import torch

from . import _lib as L
from . import ops

F32, BF16, F16 = torch.float32, torch.bfloat16, torch.float16
EXACT = False
# A/B switches for same-box measurements (scripts/gpu/ab_env.sh); the defaults are the measured-faster forms
import os as _os
MLP_SAVES_DERIVATIVE = _os.environ.get("BF_MLP_DERIV", "1") != "0"     # fc1 stores gelu'(pre); 0: stores pre, dGELU epilogue
DMUL_BN = int(_os.environ.get("BF_DMUL_BN", "0"))                       # N tile of the fc2 input-gradient GEMM (0 = automatic)
# Cluster-fused InstanceNorm calls (statistics + apply, reduce + apply in ONE launch; csrc/norm.cu).  Measured on B200 at
# config 2 (same-box A/B of the replayed step, gpurun_out/r2t_*): 26.22 vs 26.22 ms -- the second launch of the two-launch
# form already finds its operands in L2 and overlaps its prologue with the first through PDL, so the fusion buys nothing,
# and 40 six-block clusters leave fewer blocks in flight than a full wave.  Opt-in.
NORM_FUSED = _os.environ.get("BF_NORM_FUSED", "0") == "1"
# input_head bias gradient (column sums of d qkv) inside the attention backward instead of a separate pass.  With per-warp
# private tables (no shared atomics) the fusion is correct and cheap per tile, but the attention backward is bound by its
# instruction count and the separate pass finds d qkv in L2: 26.29 vs 26.08 ms per step (same-box A/B, r2v).  Opt-in.
ATTN_FUSED_BIAS = _os.environ.get("BF_ATTN_FUSED_BIAS", "0") == "1"
IMPLICIT_GATHER = _os.environ.get("BF_IMPLICIT_GATHER", "1") != "0"     # head backward reads dZ in place (no s2d_gather copy)


def set_exact_mode(on: bool) -> None:
    """fp32 validation configuration (north star: parity within 1e-4 "for the fp32 path"; upstream runs fp32 / TF32,
    scripts/train.py:72): every activation that the production path stores in 16 bits becomes fp32, the GEMMs /
    attention / patch kernels run their plain-fp32 forms (csrc/exact.cu) and every GELU is the exact erf form.  The
    orchestration in this file is unchanged.  Process-wide; a checking mode, not a performance mode."""
    global BF16, F16, EXACT
    EXACT = bool(on)
    BF16 = torch.float32 if on else torch.bfloat16
    F16 = torch.float32 if on else torch.float16
    L.lib.bf_set_gelu_mode(1 if on else 0)


@dataclass(frozen=True)
class Geom:
    B: int
    T: int
    h: int
    w: int

    @property
    def I(self) -> int:  # noqa: E743
        return self.B * self.T

    @property
    def P(self) -> int:
        return self.h * self.w

    @property
    def N(self) -> int:
        return self.B * self.T * self.h * self.w


# ---------------------------------------------------------------------------------------------
# T5 relative-position buckets (upstream layers/positional_encoding.py:76-129), host computed once per L
# ---------------------------------------------------------------------------------------------
_BUCKETS: Dict[tuple, torch.Tensor] = {}


def relpos_bucket_vector(Ln: int, device) -> torch.Tensor:
    """int32 (2L-1,): bucket of rel = key - query for rel in [-(L-1), L-1].

    Bidirectional, 32 buckets, max_distance 32 (the static default that upstream actually runs with);
    the logarithmic branch is evaluated in float32 exactly like upstream so bucket edges agree.
    """
    key = (Ln, str(device))
    if key not in _BUCKETS:
        rel = torch.arange(-(Ln - 1), Ln, dtype=torch.long)
        n = -rel
        ret = (n < 0).to(torch.long) * 16
        n = n.abs()
        large = 8 + (torch.log(n.float() / 8) / math.log(32 / 8) * 8).to(torch.long)
        large = torch.minimum(large, torch.full_like(large, 15))
        ret = ret + torch.where(n < 8, n, large)
        _BUCKETS[key] = ret.to(torch.int32).to(device)
    return _BUCKETS[key]


def pick_split(k_tokens: int, m: int, n: int) -> int:
    """Split-K factor for a token-contraction (wgrad) GEMM: the (m-tiles x n-tiles x splits) work items should fill
    the 148 SMs in as few, as full, waves as possible.  Mirrors pick_bn() of gemm_sm100.cu for the n tile."""
    k_total = (k_tokens + 63) // 64
    bn = 256 if (n % 256 == 0 and n >= 512) else (192 if n % 192 == 0 else 128)
    if n <= 64:
        bn = 64
    elif n <= 128:
        bn = 128
    tiles = ((m + 127) // 128) * ((n + bn - 1) // bn)
    best, best_cost = 1, None
    for s in range(1, min(k_total, 64) + 1):
        waves = -(-tiles * s // 148)
        iters = -(-k_total // s)
        cost = waves * (iters + 6)            # k iterations per work item + a fixed prologue/epilogue share
        if best_cost is None or cost < best_cost:
            best, best_cost = s, cost
    return best


def _empty(shape, dtype, like):
    return torch.empty(shape, dtype=dtype, device=like.device)


# The block weight-gradient GEMMs have no consumer inside the backward pass (only the optimiser / the gradient
# all-reduce read them), so they leave the dependent chain: each is issued on a second stream that forks from the
# current one at the point of issue and is joined before the block's backward returns (every operand is still
# referenced until then, and GradSink's bucket logic sees a finished block exactly as before).  Their CTAs fill the
# partial last waves of the input-gradient / norm kernels on the main chain; inside a CUDA-graph capture the fork and
# join become graph edges.  Measured on B200 (same-box A/B of the replayed config-2 step, profiles/r6_wgrad_stream.txt):
# 25.08 -> 24.78 ms per step.  Also moving the bias column sums, the per-channel parameter kernels and the stem / head
# weight gradients there, or forking before the input-gradient GEMM, measured neutral and is not kept.  The join comes
# BEFORE the block's branch_param_grads launch, the one main-stream kernel that also writes a weight-gradient buffer.
# BF_WGRAD_STREAM=0 keeps everything on one stream.  (Module-level state: one backward pass per device at a time, which
# is how the autograd engine runs a device's nodes.)
WGRAD_STREAM = _os.environ.get("BF_WGRAD_STREAM", "1") != "0"
_SIDE_STREAMS: Dict[int, "torch.cuda.Stream"] = {}
_SIDE_DIRTY = False


class _WgradStream:
    def __enter__(self):
        global _SIDE_DIRTY
        self.on = WGRAD_STREAM and not EXACT and torch.cuda.is_available()   # (the CPU emulation of tests/ has no streams)
        if not self.on:
            return self
        main = torch.cuda.current_stream()
        dev = main.device.index
        side = _SIDE_STREAMS.get(dev)
        if side is None:
            side = _SIDE_STREAMS[dev] = torch.cuda.Stream(device=main.device)
        side.wait_stream(main)
        self._ctx = torch.cuda.stream(side)
        self._ctx.__enter__()
        _SIDE_DIRTY = True
        return self

    def __exit__(self, *exc):
        if self.on:
            self._ctx.__exit__(*exc)
        return False


def _wgrad_join() -> None:
    global _SIDE_DIRTY
    if _SIDE_DIRTY:
        main = torch.cuda.current_stream()
        main.wait_stream(_SIDE_STREAMS[main.device.index])
        _SIDE_DIRTY = False


class _ZeroArena:
    """Small zero-initialised fp32 buffers (statistics, reductions) carved from one zero-filled chunk, so a step
    issues a handful of fill kernels instead of one per buffer.  A region is handed out once and never reused;
    `reset()` (called at the start of every model forward) drops the chunk, which also keeps CUDA-graph captures
    self-contained (the fill of the chunk is recorded inside the capture)."""
    CHUNK = 1 << 22          # floats (16 MiB: a config-2 forward + backward takes ~4 M floats of statistics / reductions)

    def __init__(self):
        self.buf = None
        self.off = 0

    def reset(self):
        self.buf = None
        self.off = 0

    def take(self, n: int, device) -> torch.Tensor:
        n8 = (n + 7) // 8 * 8
        if self.buf is None or self.buf.device != device or self.off + n8 > self.buf.numel():
            self.buf = torch.zeros(max(self.CHUNK, n8), dtype=F32, device=device)
            self.off = 0
        v = self.buf[self.off:self.off + n]
        self.off += n8
        return v


_ARENA = _ZeroArena()


def reset_arena() -> None:
    _ARENA.reset()


def _zeros(shape, like, dtype=F32):
    n = 1
    for d in shape:
        n *= d
    if dtype == F32 and n <= (1 << 18):
        return _ARENA.take(n, like.device).view(shape)
    return torch.zeros(shape, dtype=dtype, device=like.device)


# Producer -> consumer handoff of InstanceNorm statistics: the kernel that writes a residual-stream tensor also
# accumulates its per-(image, channel) sums, so the next norm1 skips its statistics pass.  Keyed by tensor identity.
_HANDOFF = None


def _publish_stats(X: torch.Tensor, st: torch.Tensor) -> None:
    global _HANDOFF
    _HANDOFF = (weakref.ref(X), st)


def _take_stats(X: torch.Tensor) -> Optional[torch.Tensor]:
    global _HANDOFF
    h, _HANDOFF = _HANDOFF, None
    if h is not None and h[0]() is X:
        return h[1]
    return None


def _norm_fwd(x, out, I, P, weight, bias) -> torch.Tensor:
    """out = InstanceNorm(x) * weight + bias; returns the raw statistics (I, C, 2) for the backward pass."""
    C_ = x.shape[1]
    if NORM_FUSED:
        st = _empty((I, C_, 2), F32, x)
        ops.inorm_apply(x, out, I, P, st, weight, bias, compute_stats=True)
    else:
        st = _zeros((I, C_, 2), x)
        ops.inorm_stats(x, I, P, st)
        ops.inorm_apply(x, out, I, P, st, weight, bias)
    return st


def _norm_bwd(gin, x, I, P, stats, weight, bias, **kw) -> None:
    """InstanceNorm backward (reduce + apply, parameter gradients accumulated by the apply launch)."""
    C_ = x.shape[1]
    if NORM_FUSED:
        ops.inorm_bwd(3, gin, x, I, P, stats, weight, bias, _empty((I, C_, 2), F32, x), **kw)
    else:
        red = _zeros((I, C_, 2), x)
        ops.inorm_bwd(1, gin, x, I, P, stats, weight, bias, red)
        ops.inorm_bwd(2, gin, x, I, P, stats, weight, bias, red, **kw)


def _axis(g: Geom, axis: str) -> dict:
    P = g.P
    if axis == "t":
        return dict(L_=g.T, n_seq=g.B * P, inner=P, outer_stride=g.T * P, inner_stride=1, tok_stride=P)
    if axis == "x":
        return dict(L_=g.w, n_seq=g.I * g.h, inner=g.h, outer_stride=P, inner_stride=g.w, tok_stride=1)
    return dict(L_=g.h, n_seq=g.I * g.w, inner=g.w, outer_stride=P, inner_stride=1, tok_stride=g.w)


# ---------------------------------------------------------------------------------------------
# shared pieces: IN -> QKV -> attention(s) -> IN -> out-projection with residual epilogue
# ---------------------------------------------------------------------------------------------
def _attn_branch_fwd(X, g: Geom, p: Dict[str, torch.Tensor], w16, heads: int, axes: List[str], scale_keys,
                     mask_img, col_scale, col_shift, gamma, want_x16: bool, save: bool, want_stats: bool = False):
    I, P, N, E = g.I, g.P, g.N, X.shape[1]
    st1 = _take_stats(X)
    Xn = _empty((N, E), BF16, X)
    if st1 is None:      # no producer accumulated them
        st1 = _norm_fwd(X, Xn, I, P, p["norm1.weight"], p["norm1.bias"])
    else:
        ops.inorm_apply(X, Xn, I, P, st1, p["norm1.weight"], p["norm1.bias"])
    QKV = _empty((N, 3 * E), BF16, X)
    # head_dim 64 and axes of up to 64 tokens: LayerNorm(q), LayerNorm(k) are computed in the QKV GEMM epilogue (xhat +
    # rstd) and the attention kernels work on the pre-normalised rows; otherwise the generic kernels normalise in place
    prenorm = (E // heads == 64) and all(_axis(g, ax)["L_"] <= 64 for ax in axes) and BF16 == torch.bfloat16
    rstd = _empty((N, heads, 2), F32, X) if prenorm else None
    if prenorm:
        ops.gemm(Xn, w16("input_head.weight"), N, 3 * E, E, epilogue=L.EPI_QKV_LN, bias=p["input_head.bias"], out16=QKV,
                 ln_head_dim=64, ln_rstd=rstd)
    else:
        ops.gemm(Xn, w16("input_head.weight"), N, 3 * E, E, epilogue=L.EPI_STORE16, bias=p["input_head.bias"], out16=QKV)
    O = _empty((N, E), BF16, X)
    oscale = 1.0 / len(axes)
    for i, ax in enumerate(axes):
        geo = _axis(g, ax)
        sf = p[scale_keys[i]].reshape(-1) if scale_keys is not None else None
        ops.attention(QKV, O, heads=heads, qn_w=p["qnorm.weight"], qn_b=p["qnorm.bias"], kn_w=p["knorm.weight"],
                      kn_b=p["knorm.bias"], bias_emb=p["rel_pos_bias.relative_attention_bias.weight"],
                      bucket=relpos_bucket_vector(geo["L_"], X.device), scale_factor=sf, out_scale=oscale,
                      accumulate=i > 0, prenorm=prenorm, **geo)
    On = _empty((N, E), BF16, X)
    st2 = _norm_fwd(O, On, I, P, p["norm2.weight"], p["norm2.bias"])
    Xout = _empty((N, E), F32, X)
    Z = _empty((N, E), BF16, X) if save else None
    X16 = _empty((N, E), BF16, X) if want_x16 else None
    st_out = _zeros((I, E, 2), X) if (want_stats and P % 32 == 0) else None
    ops.gemm(On, w16("output_head.weight"), N, E, E, epilogue=L.EPI_RESID, bias=p["output_head.bias"],
             col_scale=col_scale, col_shift=col_shift, col_gamma=gamma, row_scale=mask_img, rows_per_group=P,
             in32=X, out32=Xout, out16=X16, out16b=Z, stats_out=st_out)
    if st_out is not None:
        _publish_stats(Xout, st_out)
    saved = dict(X=X, st1=st1, Xn=Xn, QKV=QKV, rstd=rstd, O=O, st2=st2, On=On, Z=Z) if save else None
    return Xout, X16, saved


def _attn_branch_bwd(dXout, g: Geom, p, w16, heads: int, axes, scale_keys, mask_img, coef, sv, grads):
    """Backward of X_out = X + mask*gamma*(Z*c1 + c0) through out-proj, IN, attention(s), QKV, IN.

    `coef` = gamma*c1 (the factor between dX_out and dZ).  Returns (dX, S01) with the per-image sums
    S01[0][img, c] = sum mask*dX_out, S01[1][img, c] = sum mask*dX_out*Z for the caller's gamma / feature-scale gradients.
    Accumulates the gradients of norm1/2, input_head, output_head.weight, qnorm/knorm, bias table, scales.
    """
    I, P, N = g.I, g.P, g.N
    E = dXout.shape[1]
    X, st1, Xn, QKV, O, st2, On, Z = (sv[k] for k in ("X", "st1", "Xn", "QKV", "O", "st2", "On", "Z"))
    S01 = _zeros((2, I, E), dXout)                    # per-image partial sums (few atomics per address)
    dZ = _empty((N, E), BF16, dXout)
    ops.resid_bwd(dXout, Z, dZ, I, P, mask_img, coef, S01[0], S01[1])
    # output_head: dgrad reads W (E_out, E_in) as the (K, N) operand, wgrad contracts over tokens
    dOn = _empty((N, E), BF16, dXout)
    ops.gemm(dZ, w16("output_head.weight"), N, E, E, epilogue=L.EPI_STORE16, b_mode=L.B_KN, out16=dOn)
    with _WgradStream():
        ops.gemm(dZ, On, E, E, N, epilogue=L.EPI_ATOMIC32, a_mode=L.A_KM, b_mode=L.B_KN, split_k=pick_split(N, E, E),
                 out32=grads["output_head.weight"].view(E, E))
    # norm2
    dO = _empty((N, E), BF16, dXout)
    _norm_bwd(dOn, O, I, P, st2, p["norm2.weight"], p["norm2.bias"], out=dO,
              dweight=grads["norm2.weight"], dbias=grads["norm2.bias"])
    # attention(s)
    dQKV = _empty((N, 3 * E), BF16, dXout)
    oscale = 1.0 / len(axes)
    # The pre-normalised attention backward (axes of up to 32 tokens) can accumulate the column sums of the d qkv it
    # writes (input_head bias gradient, `d_qkv_bias`) in per-warp private shared-memory tables; see ATTN_FUSED_BIAS.
    fused_bias = ATTN_FUSED_BIAS and sv["rstd"] is not None and all(_axis(g, ax)["L_"] <= 32 for ax in axes)
    for i, ax in enumerate(axes):
        geo = _axis(g, ax)
        sf = p[scale_keys[i]].reshape(-1) if scale_keys is not None else None
        gr = dict(d_qn_w=grads["qnorm.weight"], d_qn_b=grads["qnorm.bias"], d_kn_w=grads["knorm.weight"],
                  d_kn_b=grads["knorm.bias"], d_bias_emb=grads["rel_pos_bias.relative_attention_bias.weight"],
                  d_scale_factor=grads[scale_keys[i]].view(-1) if scale_keys is not None else None,
                  d_qkv_bias=grads["input_head.bias"] if fused_bias else None)
        ops.attention(QKV, dQKV, heads=heads, qn_w=p["qnorm.weight"], qn_b=p["qnorm.bias"], kn_w=p["knorm.weight"],
                      kn_b=p["knorm.bias"], bias_emb=p["rel_pos_bias.relative_attention_bias.weight"],
                      bucket=relpos_bucket_vector(geo["L_"], dXout.device), scale_factor=sf, out_scale=oscale,
                      accumulate=i > 0, dout=dO, grads=gr, prenorm=sv["rstd"] is not None, rstd=sv["rstd"], **geo)
    if not fused_bias:
        # (Measured: taking these column sums from one extra 16-column MMA against a tile of ones inside the weight-
        # gradient GEMM costs 8 us there and saves nothing in the step -- dQKV is still L2 resident for this pass.)
        ops.colsum16(dQKV, grads["input_head.bias"])
    # input_head
    dXn = _empty((N, E), BF16, dXout)
    ops.gemm(dQKV, w16("input_head.weight"), N, E, 3 * E, epilogue=L.EPI_STORE16, b_mode=L.B_KN, out16=dXn)
    with _WgradStream():
        ops.gemm(dQKV, Xn, 3 * E, E, N, epilogue=L.EPI_ATOMIC32, a_mode=L.A_KM, b_mode=L.B_KN,
                 split_k=pick_split(N, 3 * E, E), out32=grads["input_head.weight"].view(3 * E, E))
    # norm1 (+ the identity path of the residual)
    dX = _empty((N, E), F32, dXout)
    _norm_bwd(dXn, X, I, P, st1, p["norm1.weight"], p["norm1.bias"], out=dX, add32=dXout,
              dweight=grads["norm1.weight"], dbias=grads["norm1.bias"])
    return dX, S01


# ---------------------------------------------------------------------------------------------
# temporal block
# ---------------------------------------------------------------------------------------------
def temporal_forward(X, g: Geom, p, w16, heads: int, attn_scale: bool, mask_img, save: bool):
    keys = ["attn_scale_factor"] if attn_scale else None
    Xout, _, saved = _attn_branch_fwd(X, g, p, w16, heads, ["t"], keys, mask_img, None, None, p["gamma"], False, save,
                                      want_stats=True)
    return Xout, saved


def temporal_backward(dXout, g: Geom, p, w16, heads: int, attn_scale: bool, mask_img, sv, grads):
    keys = ["attn_scale_factor"] if attn_scale else None
    dX, S01 = _attn_branch_bwd(dXout, g, p, w16, heads, ["t"], keys, mask_img, p["gamma"], sv, grads)
    # d_gamma += S1 (d/dgamma of mask*gamma*Z), d_output_head.bias += gamma*S0
    _wgrad_join()
    ops.branch_param_grads(S01, p["gamma"], grads["gamma"], grads["output_head.bias"])
    return dX


# ---------------------------------------------------------------------------------------------
# spatial block (axial attention + MLP)
# ---------------------------------------------------------------------------------------------
def _feat_consts(p, feat_scale: bool):
    """Feature scaling  z + mean_img(z)*low + (z - mean_img(z))*high  (attention.py:302-307).

    mean_img(z) over an image of z = IN(o) W^T + b is exactly W b_norm2 + b_out =: c (the normalised part of
    IN has zero mean per image and channel), so the op is the per-channel affine z*(1+high) + c*(low-high).
    """
    if not feat_scale:
        return None, None, None, None
    return ops.feat_consts(p["output_head.weight"], p["norm2.bias"], p["output_head.bias"], p["low_freq_scalar"],
                           p["high_freq_scalar"], gamma=p["gamma_att"])


def spatial_forward(X, g: Geom, p, w16, heads: int, attn_scale: bool, feat_scale: bool, mask_att, mask_mlp, save: bool):
    I, P, N, E = g.I, g.P, g.N, X.shape[1]
    keys = ["attn_scale_factor_x", "attn_scale_factor_y"] if attn_scale else None
    c, c1, c0, coef = _feat_consts(p, feat_scale)
    Xmid, Xb, sv = _attn_branch_fwd(X, g, p, w16, heads, ["x", "y"], keys, mask_att, c1, c0, p["gamma_att"], True, save)
    if save:
        sv["feat"] = (c, c1, c0, coef)
    G = _empty((N, 4 * E), BF16, X)
    Hpre = _empty((N, 4 * E), BF16, X) if save else None
    # the second output is gelu'(pre), not pre: the tanh is evaluated once and the backward epilogue is a multiply
    ops.gemm(Xb, w16("mlp.fc1.weight"), N, 4 * E, E, epilogue=L.EPI_GELU_D if MLP_SAVES_DERIVATIVE else L.EPI_GELU,
             bias=p["mlp.fc1.bias"], out16=G, out16b=Hpre)
    Y2 = _empty((N, E), BF16, X)
    st3 = _zeros((I, E, 2), X)
    if P % 32 == 0:      # the fc2 epilogue accumulates the statistics of what it stores
        ops.gemm(G, w16("mlp.fc2.weight"), N, E, 4 * E, epilogue=L.EPI_STORE16, bias=p["mlp.fc2.bias"], out16=Y2,
                 rows_per_group=P, stats_out=st3)
    else:
        ops.gemm(G, w16("mlp.fc2.weight"), N, E, 4 * E, epilogue=L.EPI_STORE16, bias=p["mlp.fc2.bias"], out16=Y2)
        ops.inorm_stats(Y2, I, P, st3)
    Xout = _empty((N, E), F32, X)
    st_out = _zeros((I, E, 2), X)
    ops.inorm_apply(Y2, Xout, I, P, st3, p["mlp_norm.weight"], p["mlp_norm.bias"], resid_in=Xmid, row_scale=mask_mlp,
                    col_gamma=p["gamma_mlp"], stats_out=st_out)
    _publish_stats(Xout, st_out)
    if save:
        sv.update(Xb=Xb, G=G, Hpre=Hpre, Y2=Y2, st3=st3)
    return Xout, sv


def spatial_backward(dXout, g: Geom, p, w16, heads: int, attn_scale: bool, feat_scale: bool, mask_att, mask_mlp, sv,
                     grads):
    I, P, N = g.I, g.P, g.N
    E = dXout.shape[1]
    Xb, G, Hpre, Y2, st3 = (sv[k] for k in ("Xb", "G", "Hpre", "Y2", "st3"))
    # ---- MLP branch: X_out = X_mid + mask*gamma_mlp*IN(Y2) ----
    dY2 = _empty((N, E), BF16, dXout)
    _norm_bwd(dXout, Y2, I, P, st3, p["mlp_norm.weight"], p["mlp_norm.bias"], out=dY2,
                  row_scale=mask_mlp, col_scale=p["gamma_mlp"], dweight=grads["mlp_norm.weight"],
                  dbias=grads["mlp_norm.bias"], dcol_scale=grads["gamma_mlp"])
    # fc2 (its bias feeds an InstanceNorm, so its gradient is identically zero and stays zero)
    dH = _empty((N, 4 * E), BF16, dXout)
    ops.gemm(dY2, w16("mlp.fc2.weight"), N, 4 * E, E, epilogue=L.EPI_DMUL if MLP_SAVES_DERIVATIVE else L.EPI_DGELU,
             b_mode=L.B_KN, aux16=Hpre, out16=dH, bn=DMUL_BN,
             colsum_out=grads["mlp.fc1.bias"])
    with _WgradStream():
        ops.gemm(dY2, G, E, 4 * E, N, epilogue=L.EPI_ATOMIC32, a_mode=L.A_KM, b_mode=L.B_KN,
                 split_k=pick_split(N, E, 4 * E), out32=grads["mlp.fc2.weight"])
    # fc1: the input gradient joins the residual-stream gradient in the epilogue
    dXmid = _empty((N, E), F32, dXout)
    ops.gemm(dH, w16("mlp.fc1.weight"), N, E, 4 * E, epilogue=L.EPI_ACC32, b_mode=L.B_KN, in32=dXout, out32=dXmid)
    with _WgradStream():
        ops.gemm(dH, Xb, 4 * E, E, N, epilogue=L.EPI_ATOMIC32, a_mode=L.A_KM, b_mode=L.B_KN,
                 split_k=pick_split(N, 4 * E, E), out32=grads["mlp.fc1.weight"])
    # ---- attention branch ----
    keys = ["attn_scale_factor_x", "attn_scale_factor_y"] if attn_scale else None
    c, c1, c0, coef = sv["feat"] if "feat" in sv else _feat_consts(p, feat_scale)
    ga = p["gamma_att"]
    if not feat_scale:
        coef = ga
    dX, S01 = _attn_branch_bwd(dXmid, g, p, w16, heads, ["x", "y"], keys, mask_att, coef, sv, grads)
    feat = None
    if feat_scale:
        # gamma_att, low / high_freq_scalar, output_head.bias and -- through c = W b_norm2 + b_out -- output_head.weight
        # and norm2.bias (formulas in include/bubbleformer_b200.h, bf_branch_param_grads)
        feat = dict(c=c, c1=c1, c0=c0, low=p["low_freq_scalar"], high=p["high_freq_scalar"], W=p["output_head.weight"],
                    norm2_bias=p["norm2.bias"], d_low=grads["low_freq_scalar"], d_high=grads["high_freq_scalar"],
                    d_W=grads["output_head.weight"], d_norm2_bias=grads["norm2.bias"])
    # (joined first: branch_param_grads adds the feature-scale term to d output_head.weight with plain read-modify-writes,
    # which must not run beside the weight-gradient GEMM's reduce-adds into the same buffer)
    _wgrad_join()
    ops.branch_param_grads(S01, ga, grads["gamma_att"], grads["output_head.bias"], feat)
    return dX


# ---------------------------------------------------------------------------------------------
# hierarchical patch embed (+ FiLM) and unembed
# ---------------------------------------------------------------------------------------------
def _conv_w_fwd(w: torch.Tensor, dtype) -> torch.Tensor:
    """(Cout, Cin, 2, 2) -> (Cout, (ky, kx, ci)) operand copy."""
    return w.detach().permute(0, 2, 3, 1).reshape(w.shape[0], -1).to(dtype).contiguous()


def _s2d_ok(wo: int) -> bool:
    return (wo % 128 == 0) if wo >= 128 else (128 % wo == 0)


def embed_forward(x, g_in, p, n_layers: int, film_gb: Optional[torch.Tensor], T: int, save: bool):
    """x: (I, F, H, W) fp32.  p: embed params keyed 'in_proj.k.weight/bias'.  Returns X (I*h*w, E) fp32."""
    I, Fd, H, W = x.shape
    sv = dict(x=x, Y=[], st=[], A=[], dims=[], film_gb=film_gb) if save else None
    A = None
    h_, w_ = H, W
    X = None
    for i in range(n_layers):
        last = i == n_layers - 1
        wt = p[f"in_proj.{3 * i}.weight"]
        Cout, Cin = wt.shape[0], wt.shape[1]
        ho, wo = h_ // 2, w_ // 2
        M = I * ho * wo
        st = _zeros((I, Cout, 2), x)
        Y = _empty((M, Cout), F16, x)
        if i == 0:
            Wkn = wt.detach().reshape(Cout, 4 * Fd).t().contiguous()
            ops.patch_in(x, Wkn, Y.view(I, ho, wo, Cout), st)
        else:
            W16 = _conv_w_fwd(wt, F16)
            fuse = (ho * wo) % 32 == 0          # the GEMM epilogue accumulates the statistics of what it stores
            skw = dict(rows_per_group=ho * wo, stats_out=st) if fuse else {}
            if _s2d_ok(wo):
                ops.gemm(A, W16, M, Cout, 4 * Cin, epilogue=L.EPI_STORE16, a_mode=L.A_S2D, ldb=4 * Cin,
                         s2d=(I, h_, w_, Cin), out16=Y, **skw)
            else:
                Ag = _empty((M, 4 * Cin), F16, x)
                ops.s2d_gather(A.view(I, h_, w_, Cin), Ag)
                ops.gemm(Ag, W16, M, Cout, 4 * Cin, epilogue=L.EPI_STORE16, out16=Y, **skw)
            if not fuse:
                ops.inorm_stats(Y, I, ho * wo, st)
        nw, nb = p[f"in_proj.{3 * i + 1}.weight"], p[f"in_proj.{3 * i + 1}.bias"]
        if not last:
            An = _empty((M, Cout), F16, x)
            ops.inorm_apply(Y, An, I, ho * wo, st, nw, nb, gelu=True)
        else:
            X = _empty((M, Cout), F32, x)
            st_out = _zeros((I, Cout, 2), x)
            if film_gb is not None:
                ops.inorm_apply(Y, X, I, ho * wo, st, nw, nb, film_gb=film_gb, film_T=T, stats_out=st_out)
            else:
                ops.inorm_apply(Y, X, I, ho * wo, st, nw, nb, stats_out=st_out)
            _publish_stats(X, st_out)
            An = None
        if save:
            sv["Y"].append(Y); sv["st"].append(st); sv["A"].append(A); sv["dims"].append((h_, w_, Cin, Cout))
        A = An
        h_, w_ = ho, wo
    return X, sv


def embed_backward(dX, p, n_layers: int, film_gb, T: int, sv, grads, need_dx: bool):
    """Returns (dx or None, d_film_gb or None).  Gradients inside the stem are bf16 (range), activations fp16."""
    x = sv["x"]
    I, Fd, H, W = x.shape
    dfilm = None
    gin = dX                      # gradient w.r.t. the output of stage i's IN(+GELU / +FiLM)
    for i in reversed(range(n_layers)):
        last = i == n_layers - 1
        h_, w_, Cin, Cout = sv["dims"][i]
        ho, wo = h_ // 2, w_ // 2
        M = I * ho * wo
        Y, st = sv["Y"][i], sv["st"][i]
        nw, nb = p[f"in_proj.{3 * i + 1}.weight"], p[f"in_proj.{3 * i + 1}.bias"]
        red = _zeros((I, Cout, 2), dX)
        ops.inorm_bwd(1, gin, Y, I, ho * wo, st, nw, nb, red, gelu=not last)
        dY = _empty((M, Cout), BF16, dX)
        pk = dict(dweight=grads[f"in_proj.{3 * i + 1}.weight"], dbias=grads[f"in_proj.{3 * i + 1}.bias"])
        if last and film_gb is not None:
            dfilm = _zeros((I // T, 2 * Cout), dX)          # [d gamma | d beta], the layout of bf_film_fwd's output
            ops.inorm_bwd(2, gin, Y, I, ho * wo, st, nw, nb, red, gelu=not last, out=dY, film_gb=film_gb, film_T=T,
                          dfilm_gb=dfilm, **pk)
        else:
            ops.inorm_bwd(2, gin, Y, I, ho * wo, st, nw, nb, red, gelu=not last, out=dY, **pk)
        wt = p[f"in_proj.{3 * i}.weight"]
        if i == 0:
            ops.patch_wgrad(dY.view(I, ho, wo, Cout), x, grads[f"in_proj.0.weight"])
            if need_dx:
                dx = _empty((I, Fd, H, W), F32, dX)
                ops.patch_out(dY.view(I, ho, wo, Cout), wt.detach().reshape(Cout, 4 * Fd).contiguous(), dx)
                return dx, dfilm
            return None, dfilm
        # GEMM stage: wgrad over the gathered (bf16) patches, dgrad scattered back depth-to-space
        A = sv["A"][i]
        Ag = _empty((M, 4 * Cin), BF16, dX)
        ops.s2d_gather(A.view(I, h_, w_, Cin), Ag)
        dWp = _zeros((Cout, 4 * Cin), dX)
        ops.gemm(dY, Ag, Cout, 4 * Cin, M, epilogue=L.EPI_ATOMIC32, a_mode=L.A_KM, b_mode=L.B_KN,
                 split_k=pick_split(M, Cout, 4 * Cin), out32=dWp)
        grads[f"in_proj.{3 * i}.weight"] += dWp.view(Cout, 2, 2, Cin).permute(0, 3, 1, 2)
        del Ag
        dA = _empty((I * h_ * w_, Cin), BF16, dX)
        ops.gemm(dY, _conv_w_fwd(wt, BF16), M, 4 * Cin, Cout, epilogue=L.EPI_D2S, b_mode=L.B_KN, d2s=(ho, wo, Cin),
                 out16=dA, ldo=4 * Cin)
        gin = dA
    raise AssertionError("unreachable")


def _convT_w(w: torch.Tensor, dtype) -> torch.Tensor:
    """(Cin, Cout, 2, 2) -> (Cin, (ky, kx, co)) operand copy."""
    return w.detach().permute(0, 2, 3, 1).reshape(w.shape[0], -1).to(dtype).contiguous()


def debed_forward(X, g: Geom, p, n_layers: int, out_fields: int, save: bool):
    """X: (I*h*w, E) fp32 -> (I, F, H, W) fp32."""
    I = g.I
    h_, w_ = g.h, g.w
    Xh = _empty(tuple(X.shape), F16, X)
    ops.cast16(X, Xh)
    A = Xh
    sv = dict(A=[], Z=[], st=[], dims=[]) if save else None
    out = None
    for i in range(n_layers):
        last = i == n_layers - 1
        wt = p[f"out_proj.{3 * i}.weight"]
        Cin, Cout = wt.shape[0], wt.shape[1]
        M = I * h_ * w_
        if last:
            out = _empty((I, Cout, 2 * h_, 2 * w_), F32, X)
            ops.patch_out(A.view(I, h_, w_, Cin), wt.detach().reshape(Cin, 4 * Cout).contiguous(), out)
            if save:
                sv["A"].append(A); sv["dims"].append((h_, w_, Cin, Cout))
            break
        Z = _empty((4 * M, Cout), F16, X)
        ops.gemm(A, _convT_w(wt, F16), M, 4 * Cout, Cin, epilogue=L.EPI_D2S, b_mode=L.B_KN, d2s=(h_, w_, Cout),
                 out16=Z, ldo=4 * Cout)
        st = _zeros((I, Cout, 2), X)
        ops.inorm_stats(Z, I, 4 * h_ * w_, st)
        An = _empty((4 * M, Cout), F16, X)
        ops.inorm_apply(Z, An, I, 4 * h_ * w_, st, p[f"out_proj.{3 * i + 1}.weight"], p[f"out_proj.{3 * i + 1}.bias"],
                        gelu=True)
        if save:
            sv["A"].append(A); sv["Z"].append(Z); sv["st"].append(st); sv["dims"].append((h_, w_, Cin, Cout))
        A = An
        h_, w_ = 2 * h_, 2 * w_
    return out, sv


def debed_backward(dOut, g: Geom, p, n_layers: int, sv, grads):
    """dOut: (I, F, H, W) fp32 -> dX (I*h*w, E) fp32."""
    I = g.I
    dOut = dOut.contiguous()
    gin = None
    dX = None
    for i in reversed(range(n_layers)):
        last = i == n_layers - 1
        h_, w_, Cin, Cout = sv["dims"][i]
        M = I * h_ * w_
        A = sv["A"][i]
        wt = p[f"out_proj.{3 * i}.weight"]
        if last:
            ops.patch_wgrad(A.view(I, h_, w_, Cin), dOut, grads[f"out_proj.{3 * i}.weight"])
            dA = _empty((M, Cin), BF16, dOut)
            ops.patch_in(dOut, wt.detach().reshape(Cin, 4 * Cout).t().contiguous(), dA.view(I, h_, w_, Cin), None)
            if n_layers == 1:
                return dA.float()
            gin = dA
            continue
        Z, st = sv["Z"][i], sv["st"][i]
        nw, nb = p[f"out_proj.{3 * i + 1}.weight"], p[f"out_proj.{3 * i + 1}.bias"]
        red = _zeros((I, Cout, 2), dOut)
        ops.inorm_bwd(1, gin, Z, I, 4 * h_ * w_, st, nw, nb, red, gelu=True)
        dZ = _empty((4 * M, Cout), BF16, dOut)
        ops.inorm_bwd(2, gin, Z, I, 4 * h_ * w_, st, nw, nb, red, gelu=True, out=dZ,
                      dweight=grads[f"out_proj.{3 * i + 1}.weight"], dbias=grads[f"out_proj.{3 * i + 1}.bias"])
        # dZ (I, 2h, 2w, Cout) enters both GEMMs as the 2x2 patch gather (M, (ky, kx, co)): the A operand of the dgrad and
        # the B operand of the wgrad.  With 64-pixel row segments both read it in place through 4-D tensor maps
        # (BF_A_S2D / BF_B_KN_S2D); otherwise one gathered copy is made.
        implicit = IMPLICIT_GATHER and not EXACT and _s2d_ok(w_) and w_ % 64 == 0 and (2 * Cout) % 64 == 0
        Ab = _empty(tuple(A.shape), BF16, dOut)
        ops.convert16(A, Ab)
        dWp = _zeros((Cin, 4 * Cout), dOut)
        Wb = _convT_w(wt, BF16)                                   # (Cin, 4*Cout) = (N, K) of the dgrad
        if implicit:
            s2d = (I, 2 * h_, 2 * w_, Cout)
            dZ2 = dZ.view(-1, Cout)
            ops.gemm(Ab, dZ2, Cin, 4 * Cout, M, epilogue=L.EPI_ATOMIC32, a_mode=L.A_KM, b_mode=L.B_KN_S2D, s2d=s2d,
                     split_k=pick_split(M, Cin, 4 * Cout), out32=dWp)
            akw = dict(a_mode=L.A_S2D, s2d=s2d, ldb=4 * Cout)
            dZg = dZ2
        else:
            dZg = _empty((M, 4 * Cout), BF16, dOut)
            ops.s2d_gather(dZ.view(I, 2 * h_, 2 * w_, Cout), dZg)
            ops.gemm(Ab, dZg, Cin, 4 * Cout, M, epilogue=L.EPI_ATOMIC32, a_mode=L.A_KM, b_mode=L.B_KN,
                     split_k=pick_split(M, Cin, 4 * Cout), out32=dWp)
            akw = {}
        grads[f"out_proj.{3 * i}.weight"] += dWp.view(Cin, 2, 2, Cout).permute(0, 3, 1, 2)
        if i == 0:
            dX = _empty((M, Cin), F32, dOut)
            ops.gemm(dZg, Wb, M, Cin, 4 * Cout, epilogue=L.EPI_STORE32, out32=dX, **akw)
            return dX
        dA = _empty((M, Cin), BF16, dOut)
        ops.gemm(dZg, Wb, M, Cin, 4 * Cout, epilogue=L.EPI_STORE16, out16=dA, **akw)
        gin = dA
    return dX
