"""Run each kernel twice on identical inputs and report max |a - b| (bitwise-deterministic kernels must give 0)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bubbleformer_b200 import _lib as L, engine, ops
dev = "cuda"
torch.manual_seed(0)
bf = torch.bfloat16
I, P, E, he = 5, 16, 128, 2
N = I * P
def rep(name, fn, outs):
    res = []
    for _ in range(3):
        for o in outs: o.zero_()
        fn(); torch.cuda.synchronize()
        res.append([o.clone() for o in outs])
    d = max(float((a.float() - b.float()).abs().max()) for r in res[1:] for a, b in zip(r, res[0]))
    print(f"[det] {name:28s} max diff {d:.3e}", flush=True)
for (I, h, w) in [(5, 4, 4), (10, 32, 32)]:
    P = h * w; N = I * P
    g = engine.Geom(I // 5, 5, h, w)
    print(f"--- I={I} P={P}")
    X32 = torch.randn(N, E, device=dev); Xb = X32.to(bf)
    Win = (torch.randn(3 * E, E, device=dev) * E ** -0.5).to(bf)
    Wo = (torch.randn(E, E, device=dev) * E ** -0.5).to(bf)
    W1 = (torch.randn(4 * E, E, device=dev) * E ** -0.5).to(bf)
    vE, v3E, v4E = torch.randn(E, device=dev), torch.randn(3 * E, device=dev), torch.randn(4 * E, device=dev)
    rs = torch.rand(I, device=dev)
    out3 = torch.empty(N, 3 * E, device=dev, dtype=bf); rstd = torch.empty(N, he, 2, device=dev)
    rep("gemm qkv_ln", lambda: ops.gemm(Xb, Win, N, 3 * E, E, epilogue=L.EPI_QKV_LN, bias=v3E, out16=out3, ln_head_dim=64, ln_rstd=rstd), [out3, rstd])
    o32 = torch.empty(N, E, device=dev); O = torch.empty(N, E, device=dev, dtype=bf); O2 = torch.empty_like(O)
    st = torch.zeros(I, E, 2, device=dev)
    if P % 32 == 0:
        rep("gemm resid+stats", lambda: ops.gemm(Xb, Wo, N, E, E, epilogue=L.EPI_RESID, bias=vE, col_gamma=vE, row_scale=rs, rows_per_group=P, in32=X32, out32=o32, out16=O, out16b=O2, stats_out=st), [o32, O, O2, st])
    rep("gemm resid", lambda: ops.gemm(Xb, Wo, N, E, E, epilogue=L.EPI_RESID, bias=vE, col_gamma=vE, row_scale=rs, rows_per_group=P, in32=X32, out32=o32, out16=O, out16b=O2), [o32, O, O2])
    G = torch.empty(N, 4 * E, device=dev, dtype=bf); Hp = torch.empty_like(G)
    rep("gemm gelu", lambda: ops.gemm(Xb, W1, N, 4 * E, E, epilogue=L.EPI_GELU, bias=v4E, out16=G, out16b=Hp), [G, Hp])
    ops.gemm(Xb, Win, N, 3 * E, E, epilogue=L.EPI_QKV_LN, bias=v3E, out16=out3, ln_head_dim=64, ln_rstd=rstd)
    ln = [torch.ones(64, device=dev), torch.zeros(64, device=dev) + 0.1, torch.ones(64, device=dev), torch.zeros(64, device=dev) + 0.1]
    emb = torch.randn(32, he, device=dev); sf = torch.rand(he, device=dev) + 0.5
    for ax in "xyt":
        geo = engine._axis(g, ax)
        kw = dict(heads=he, qn_w=ln[0], qn_b=ln[1], kn_w=ln[2], kn_b=ln[3], bias_emb=emb, bucket=engine.relpos_bucket_vector(geo["L_"], dev), scale_factor=sf, out_scale=0.5, **geo)
        rep(f"attn fwd {ax}", lambda: ops.attention(out3, O, prenorm=True, **kw), [O])
        dq = torch.empty(N, 3 * E, device=dev, dtype=bf)
        gr = dict(d_qn_w=torch.zeros(64, device=dev), d_qn_b=torch.zeros(64, device=dev), d_kn_w=torch.zeros(64, device=dev), d_kn_b=torch.zeros(64, device=dev), d_bias_emb=torch.zeros(32, he, device=dev), d_scale_factor=torch.zeros(he, device=dev))
        rep(f"attn bwd {ax}", lambda: ops.attention(out3, dq, dout=Xb, grads=gr, prenorm=True, rstd=rstd, **kw), [dq])
    st.zero_(); ops.inorm_stats(X32, I, P, st)
    rep("inorm_stats", lambda: ops.inorm_stats(X32, I, P, st), [st])
    ops.inorm_stats(X32, I, P, st)
    rep("inorm_apply", lambda: ops.inorm_apply(X32, O, I, P, st, vE, vE), [O])
    st2 = torch.zeros_like(st)
    rep("inorm_apply resid+stats", lambda: ops.inorm_apply(Xb, o32, I, P, st, vE, vE, resid_in=X32, row_scale=rs, col_gamma=vE, stats_out=st2), [o32, st2])
x = torch.randn(5, 4, 64, 64, device=dev)
Wkn = torch.randn(16, 32, device=dev)
out = torch.empty(5, 32, 32, 32, device=dev, dtype=torch.float16); stt = torch.zeros(5, 32, 2, device=dev)
rep("patch_in", lambda: ops.patch_in(x, Wkn, out, stt), [out, stt])
