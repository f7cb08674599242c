"""Torch-on-CPU emulation of the C-ABI kernels' *semantics* -- TEST INFRASTRUCTURE ONLY.

Used by tests/test_engine_cpu.py to exercise the host-side orchestration in bubbleformer_b200.engine
(which buffer goes where, the backward formulas, gradient bookkeeping) in the GPU-less build container, by
monkeypatching bubbleformer_b200.ops.  The product never imports this file and has no CPU path.
Each function documents the kernel contract it mirrors (include/bubbleformer_b200.h).
"""
import torch
import torch.nn.functional as F

from bubbleformer_b200 import _lib as L

EPS = 1e-5


def _gelu_grad(x):
    cdf = 0.5 * (1 + torch.erf(x / 2 ** 0.5))
    pdf = torch.exp(-0.5 * x * x) / (2 * torch.pi) ** 0.5
    return cdf + x * pdf


def gemm(A, B, M, N, K, *, epilogue, a_mode=L.A_ROWMAJOR, b_mode=L.B_NK, split_k=1, bn=0, lda=None, ldb=None,
         s2d=None, d2s=None, rows_per_group=1, bias=None, col_scale=None, col_shift=None, col_gamma=None,
         row_scale=None, in32=None, aux16=None, out16=None, out16b=None, out32=None, stats_out=None,
         ln_head_dim=0, ln_rstd=None, colsum_out=None, ldo=None, ld32=None):
    a = A.float()
    if a_mode == L.A_KM:
        a = a.reshape(K, M).t()
    elif a_mode == L.A_S2D:
        I, Hin, Win, Cin = s2d
        a = a.reshape(I, Hin // 2, 2, Win // 2, 2, Cin).permute(0, 1, 3, 2, 4, 5).reshape(M, K)
    else:
        a = a.reshape(M, K)
    b = B.float()
    if b_mode == L.B_KN_S2D:          # implicit patch gather, K = output pixels
        I, Hin, Win, Cin = s2d
        b = b.reshape(I, Hin // 2, 2, Win // 2, 2, Cin).permute(0, 1, 3, 2, 4, 5).reshape(K, N)
    else:
        b = b.reshape(K, N) if b_mode == L.B_KN else b.reshape(N, K).t()
    acc = a @ b
    if bias is not None:
        acc = acc + bias
    if epilogue == L.EPI_QKV_LN:
        d = ln_head_dim
        x = acc.reshape(M, N // (3 * d), 3, d).clone()
        qk = x[:, :, :2]
        mean = qk.mean(-1, keepdim=True)
        rstd = torch.rsqrt(qk.var(-1, unbiased=False, keepdim=True) + EPS)
        x[:, :, :2] = (qk - mean) * rstd
        out16.reshape(M, N).copy_(x.reshape(M, N))
        ln_rstd.reshape(M, N // (3 * d), 2).copy_(rstd[..., 0])
    elif epilogue == L.EPI_STORE16:
        out16.reshape(M, N).copy_(acc)
        if stats_out is not None:
            xi = out16.reshape(-1, rows_per_group, N).float()
            stats_out.reshape(-1, N, 2).add_(torch.stack([xi.sum(1), (xi * xi).sum(1)], dim=-1))
    elif epilogue == L.EPI_STORE32:
        out32.reshape(M, N).copy_(acc)
    elif epilogue == L.EPI_GELU:
        if out16b is not None:
            out16b.reshape(M, N).copy_(acc)
        out16.reshape(M, N).copy_(F.gelu(acc))
    elif epilogue == L.EPI_GELU_D:
        if out16b is not None:
            out16b.reshape(M, N).copy_(_gelu_grad(acc))
        out16.reshape(M, N).copy_(F.gelu(acc))
    elif epilogue == L.EPI_DMUL:
        out16.reshape(M, N).copy_(acc * aux16.reshape(M, N).float())
        if colsum_out is not None:
            colsum_out.add_(out16.reshape(M, N).float().sum(0))
    elif epilogue == L.EPI_RESID:
        if out16b is not None:
            out16b.reshape(M, N).copy_(acc)
        v = acc
        if col_scale is not None:
            v = v * col_scale + col_shift
        rs = 1.0
        if row_scale is not None:
            rs = row_scale.repeat_interleave(rows_per_group)[:M, None]
        xo = in32.reshape(M, N) + rs * col_gamma * v
        out32.reshape(M, N).copy_(xo)
        if out16 is not None:
            out16.reshape(M, N).copy_(xo)
        if stats_out is not None:
            xi = out32.reshape(-1, rows_per_group, N)
            stats_out.reshape(-1, N, 2).add_(torch.stack([xi.sum(1), (xi * xi).sum(1)], dim=-1))
    elif epilogue == L.EPI_DGELU:
        out16.reshape(M, N).copy_(acc * _gelu_grad(aux16.reshape(M, N).float()))
        if colsum_out is not None:
            colsum_out.add_(out16.reshape(M, N).float().sum(0))
    elif epilogue == L.EPI_ACC32:
        out32.reshape(M, N).copy_(in32.reshape(M, N) + acc)
    elif epilogue == L.EPI_ATOMIC32:
        out32.reshape(M, N).add_(acc)
    elif epilogue == L.EPI_D2S:
        h, w, co = d2s
        I = M // (h * w)
        y = acc.reshape(I, h, w, 2, 2, co).permute(0, 1, 3, 2, 4, 5).reshape(I * 2 * h * 2 * w, co)
        out16.reshape(-1, co).copy_(y)
    else:
        raise AssertionError(epilogue)


def inorm_stats(x, I, P, stats):
    xi = x.float().reshape(I, P, -1)
    stats.add_(torch.stack([xi.sum(1), (xi * xi).sum(1)], dim=-1))


def _mean_rstd(stats, P):
    mean = stats[..., 0] / P
    var = (stats[..., 1] / P - mean * mean).clamp_min(0)
    return mean, torch.rsqrt(var + EPS)


def inorm_apply(x, out, I, P, stats, weight, bias, *, gelu=False, film_gamma=None, film_beta=None, film_T=0,
                film_gb=None, resid_in=None, row_scale=None, col_gamma=None, stats_out=None, compute_stats=False):
    C = x.shape[1]
    if compute_stats:                # `stats` is an output
        stats.zero_()
        inorm_stats(x, I, P, stats)
    if film_gb is not None:          # (B, 2C) = [gamma | beta], the layout bf_film_fwd writes
        film_gamma, film_beta = film_gb[:, :C], film_gb[:, C:]
    mean, rstd = _mean_rstd(stats, P)
    xi = x.float().reshape(I, P, C)
    y = (xi - mean[:, None]) * rstd[:, None] * weight + bias
    if gelu:
        y = F.gelu(y)
    if film_gamma is not None:
        y = y * film_gamma.repeat_interleave(film_T, 0)[:, None] + film_beta.repeat_interleave(film_T, 0)[:, None]
    if resid_in is not None:
        rs = row_scale[:, None, None] if row_scale is not None else 1.0
        y = resid_in.reshape(I, P, C) + rs * col_gamma * y
    out.copy_(y.reshape(I * P, C))
    if stats_out is not None:
        yo = out.float().reshape(I, P, C)
        stats_out.add_(torch.stack([yo.sum(1), (yo * yo).sum(1)], dim=-1))


def inorm_bwd(phase, gin, x, I, P, stats, weight, bias, red, *, gelu=False, out=None, row_scale=None, col_scale=None,
              film_gamma=None, film_T=0, add32=None, dweight=None, dbias=None, dcol_scale=None, dfilm_gamma=None,
              dfilm_beta=None, film_gb=None, dfilm_gb=None):
    C = x.shape[1]
    if film_gb is not None:
        film_gamma = film_gb[:, :C]
        if dfilm_gb is not None:
            dfilm_gamma, dfilm_beta = dfilm_gb[:, :C], dfilm_gb[:, C:]
    mean, rstd = _mean_rstd(stats, P)
    xh = (x.float().reshape(I, P, C) - mean[:, None]) * rstd[:, None]
    g = gin.float().reshape(I, P, C)
    if gelu:
        g = g * _gelu_grad(xh * weight + bias)
    if phase == 3:                   # both phases, `red` is an output
        red.zero_()
    if phase in (1, 3):
        red.add_(torch.stack([g.sum(1), (g * xh).sum(1)], dim=-1))
        if phase == 1:
            return
    cs = torch.ones(I, C)
    if row_scale is not None:
        cs = cs * row_scale[:, None]
    if col_scale is not None:
        cs = cs * col_scale
    if film_gamma is not None:
        cs = cs * film_gamma.repeat_interleave(film_T, 0)
    k = (rstd * weight * cs)[:, None]
    o = k * (g - red[:, None, :, 0] / P - xh * red[:, None, :, 1] / P)
    o = o.reshape(I * P, C)
    if add32 is not None:
        o = o + add32
    out.copy_(o)
    if dweight is not None:
        fg = torch.zeros_like(dfilm_gamma) if dfilm_gamma is not None else None
        fb = torch.zeros_like(dfilm_beta) if dfilm_beta is not None else None
        inorm_bwd_params(red, I, P, C, weight, bias, row_scale=row_scale, col_scale=col_scale, film_gamma=film_gamma,
                         film_T=film_T, dweight=dweight, dbias=dbias, dcol_scale=dcol_scale, dfilm_gamma=fg, dfilm_beta=fb)
        if fg is not None:
            dfilm_gamma.add_(fg)
            dfilm_beta.add_(fb)


def inorm_bwd_params(red, I, P, Cn, weight, bias, *, row_scale=None, col_scale=None, film_gamma=None, film_T=0,
                     dweight=None, dbias=None, dcol_scale=None, dfilm_gamma=None, dfilm_beta=None):
    R1, R2 = red[..., 0], red[..., 1]
    rs = row_scale[:, None] if row_scale is not None else torch.ones(I, 1)
    cs = rs * (col_scale if col_scale is not None else 1.0)
    if film_gamma is not None:
        cs = cs * film_gamma.repeat_interleave(film_T, 0)
    if dweight is not None:
        dweight.add_((cs * R2).sum(0))
    if dbias is not None:
        dbias.add_((cs * R1).sum(0))
    if dcol_scale is not None:
        dcol_scale.add_((rs * (weight * R2 + bias * R1)).sum(0))
    if dfilm_gamma is not None:
        t = (weight * R2 + bias * R1).reshape(I // film_T, film_T, Cn)
        dfilm_gamma.copy_(t.sum(1))
        dfilm_beta.copy_(R1.reshape(I // film_T, film_T, Cn).sum(1))


def resid_bwd(dx, z16, dz16, I, P, row_scale, coef, S0, S1):
    C = dx.shape[1]
    rs = row_scale.repeat_interleave(P)[:, None] if row_scale is not None else 1.0
    if dz16 is not None:
        dz16.copy_(rs * coef * dx)
    S0.reshape(I, C).add_((rs * dx).reshape(I, P, C).sum(1))
    if z16 is not None:
        S1.reshape(I, C).add_((rs * dx * z16.float()).reshape(I, P, C).sum(1))


def feat_consts(W, norm2_bias, out_bias, low, high, gamma=None):
    Wm = W.reshape(W.shape[0], -1)
    c = (Wm * norm2_bias[None, :]).sum(dim=1) + out_bias
    if gamma is not None:
        return c, 1.0 + high, c * (low - high), gamma * (1.0 + high)
    return c, 1.0 + high, c * (low - high)


def branch_param_grads(S01, gamma, d_gamma, d_out_bias, feat=None):
    S0, S1 = S01.sum(dim=1)
    if feat is None:
        d_gamma.add_(S1)
        d_out_bias.add_(gamma * S0)
        return
    c, c1, c0, lo, hi = feat["c"], feat["c1"], feat["c0"], feat["low"], feat["high"]
    E = S0.shape[0]
    d_gamma.add_(c1 * S1 + c0 * S0)
    feat["d_high"].add_(gamma * (S1 - c * S0))
    feat["d_low"].add_(gamma * c * S0)
    dc = gamma * (lo - hi) * S0
    feat["d_W"].view(E, E).add_(dc[:, None] * feat["norm2_bias"][None, :])
    feat["d_norm2_bias"].add_((feat["W"].reshape(E, E) * dc[:, None]).sum(dim=0))
    d_out_bias.add_(gamma * c1 * S0 + dc)


def film_fwd(cond, ln_w, ln_b, W, bias):
    """bf_film_fwd: gb = Linear(LayerNorm(cond)) (upstream linear_layers.py:58-61)."""
    return F.linear(F.layer_norm(cond.float(), (cond.shape[1],), ln_w, ln_b, EPS), W, bias).detach()


def film_bwd(dgb, cond, ln_w, ln_b, W, d_ln_w, d_ln_b, d_W, d_bias):
    """bf_film_bwd: accumulates the FiLM MLP parameter gradients."""
    xh = F.layer_norm(cond.float(), (cond.shape[1],), None, None, EPS)
    c = xh * ln_w + ln_b
    d_W.add_(dgb.t() @ c)
    d_bias.add_(dgb.sum(0))
    dc = dgb @ W
    d_ln_w.add_((dc * xh).sum(0))
    d_ln_b.add_(dc.sum(0))


def colsum16(x, out):
    out.add_(x.float().sum(0))


def _seq_index(n_seq, L_, inner, outer_stride, inner_stride, tok_stride):
    s = torch.arange(n_seq)
    base = (s // inner) * outer_stride + (s % inner) * inner_stride
    return base[:, None] + torch.arange(L_)[None, :] * tok_stride          # (n_seq, L)


def attention(qkv, out, *, heads, L_, n_seq, inner, outer_stride, inner_stride, tok_stride, qn_w, qn_b, kn_w, kn_b,
              bias_emb, bucket, scale_factor=None, out_scale=1.0, accumulate=False, dout=None, grads=None,
              prenorm=False, rstd=None):
    with torch.enable_grad():       # called from inside autograd.Function.backward (grad mode off)
        return _attention(qkv, out, heads, L_, n_seq, inner, outer_stride, inner_stride, tok_stride, qn_w, qn_b, kn_w,
                          kn_b, bias_emb, bucket, scale_factor, out_scale, accumulate, dout, grads, prenorm, rstd)


def _attention(qkv, out, heads, L_, n_seq, inner, outer_stride, inner_stride, tok_stride, qn_w, qn_b, kn_w, kn_b,
               bias_emb, bucket, scale_factor, out_scale, accumulate, dout, grads, prenorm=False, rstd=None):
    idx = _seq_index(n_seq, L_, inner, outer_stride, inner_stride, tok_stride)
    E3 = qkv.shape[1]
    E = E3 // 3
    d = E // heads
    leaves = [t.detach().clone().requires_grad_(dout is not None) for t in (qkv.float(), qn_w, qn_b, kn_w, kn_b, bias_emb)]
    sf = scale_factor.detach().clone().requires_grad_(dout is not None) if scale_factor is not None else None
    q32 = leaves[0]
    x = q32[idx].reshape(n_seq, L_, heads, 3, d).permute(0, 2, 3, 1, 4)
    q, k, v = x[:, :, 0], x[:, :, 1], x[:, :, 2]
    if prenorm:                     # rows already hold xhat: only the affine part is left
        q = q * leaves[1] + leaves[2]
        k = k * leaves[3] + leaves[4]
    else:
        q = F.layer_norm(q, (d,), leaves[1], leaves[2], EPS)
        k = F.layer_norm(k, (d,), leaves[3], leaves[4], EPS)
    i = torch.arange(L_)
    bias = leaves[5][bucket.long()[i[None, :] - i[:, None] + L_ - 1]].permute(2, 0, 1)
    p = torch.softmax(q @ k.transpose(-1, -2) * d ** -0.5 + bias, dim=-1)
    if sf is not None:
        low = float(torch.ones((), dtype=torch.float32) / L_)
        p = low + (p - low) * sf.reshape(1, heads, 1, 1)
    o = (p @ v).permute(0, 2, 1, 3).reshape(n_seq, L_, E) * out_scale
    if dout is None:
        res = torch.zeros(out.shape[0], E)
        res[idx.reshape(-1)] = o.reshape(-1, E).detach()
        if accumulate:
            out.copy_(out.float() + res)
        else:
            out.copy_(res)
        return
    (o * dout.float()[idx]).sum().backward()
    dq = leaves[0].grad
    if prenorm:                     # gradient w.r.t. xhat -> gradient w.r.t. the raw rows (LayerNorm backward)
        T = qkv.shape[0]
        gx = dq.reshape(T, heads, 3, d).clone()
        xh = qkv.float().reshape(T, heads, 3, d)[:, :, :2]
        gqk = gx[:, :, :2]
        gx[:, :, :2] = rstd.reshape(T, heads, 2, 1) * (gqk - gqk.mean(-1, keepdim=True)
                                                      - xh * (gqk * xh).mean(-1, keepdim=True))
        dq = gx.reshape(T, 3 * heads * d)
    if accumulate:
        out.copy_(out.float() + dq)
    else:
        out.copy_(dq)
    grads["d_qn_w"].add_(leaves[1].grad)
    grads["d_qn_b"].add_(leaves[2].grad)
    grads["d_kn_w"].add_(leaves[3].grad)
    grads["d_kn_b"].add_(leaves[4].grad)
    if grads.get("d_qkv_bias") is not None:
        grads["d_qkv_bias"].add_(dq.reshape(-1, E3).sum(0))
    if grads.get("d_bias_emb") is not None:
        grads["d_bias_emb"].add_(leaves[5].grad)
    if sf is not None and grads.get("d_scale_factor") is not None:
        grads["d_scale_factor"].add_(sf.grad.reshape(grads["d_scale_factor"].shape))


def patch_in(x, Wkn, out, stats):
    I, Fd, H, W = x.shape
    N = out.shape[-1]
    xs = x.reshape(I, Fd, H // 2, 2, W // 2, 2).permute(0, 2, 4, 1, 3, 5).reshape(-1, 4 * Fd)
    y = (xs @ Wkn).reshape(I, H // 2, W // 2, N)
    out.copy_(y)
    if stats is not None:
        yo = out.float().reshape(I, -1, N)
        stats.add_(torch.stack([yo.sum(1), (yo * yo).sum(1)], dim=-1))


def patch_out(a, Wck, out):
    I, h, w, C = a.shape
    Fd = out.shape[1]
    y = (a.float().reshape(-1, C) @ Wck).reshape(I, h, w, Fd, 2, 2).permute(0, 3, 1, 4, 2, 5)
    out.copy_(y.reshape(I, Fd, 2 * h, 2 * w))


def patch_wgrad(a, x, dW):
    I, Fd, H, W = x.shape
    N = a.shape[-1]
    xs = x.reshape(I, Fd, H // 2, 2, W // 2, 2).permute(0, 2, 4, 1, 3, 5).reshape(-1, 4 * Fd)
    dW.add_((a.float().reshape(-1, N).t() @ xs).reshape(dW.shape))


def s2d_gather(img, out):
    I, H, W, C = img.shape
    out.copy_(img.reshape(I, H // 2, 2, W // 2, 2, C).permute(0, 1, 3, 2, 4, 5).reshape(out.shape))


def cast16(src, dst):
    dst.copy_(src.reshape(dst.shape))


def convert16(src, dst):
    dst.copy_(src)


def lploss_sums(pred, tgt, sums):
    d = (pred - tgt).flatten(-2)
    sums.reshape(-1, 2).add_(torch.stack([(d * d).sum(-1).reshape(-1), (tgt.flatten(-2) ** 2).sum(-1).reshape(-1)], dim=-1))


def lploss_bwd(pred, tgt, coef, dpred):
    dpred.copy_(coef.reshape(*pred.shape[:-2], 1, 1) * (pred - tgt))


ALL = ["film_fwd", "film_bwd", "gemm", "inorm_stats", "inorm_apply", "inorm_bwd", "inorm_bwd_params", "resid_bwd", "colsum16", "feat_consts",
       "branch_param_grads", "attention",
       "patch_in", "patch_out", "patch_wgrad", "s2d_gather", "cast16", "convert16", "lploss_sums", "lploss_bwd"]


def install(monkeypatch, wide: bool = True):
    """Patch bubbleformer_b200.ops with the emulations; with `wide` the 16-bit buffers become fp32 so that the
    orchestration can be checked against the fp64 fixtures to ~1e-5 instead of bf16 accuracy."""
    import sys
    from bubbleformer_b200 import engine, ops
    me = sys.modules[__name__]
    for n in ALL:
        monkeypatch.setattr(ops, n, getattr(me, n))
    if wide:
        monkeypatch.setattr(engine, "BF16", torch.float32)
        monkeypatch.setattr(engine, "F16", torch.float32)
