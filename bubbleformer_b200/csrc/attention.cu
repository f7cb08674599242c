// Fused 1-D attention along one axis (time, image rows, image columns) of the token grid.
//
// Replaces upstream layers/attention.py:80-101 (temporal) and :212-238, :258-277 (axial x / y):
// head split, LayerNorm(head_dim) on q and k, T5 relative-position bias, softmax, the
// "high-frequency" attention scaling  attn = 1/L + (softmax - 1/L) * s_head , attn @ v, and the inverse
// rearrange -- forward and backward -- without materialising any permuted copy: sequences are
// addressed in place in the token-major (tokens, 3E) QKV matrix through (base, stride).
//
// The problems are tiny (L <= 128, head_dim <= 128, thousands of independent (sequence, head) pairs) and
// HBM-bound, so one warp owns one (sequence, head) pair: operands are staged in shared memory with
// 16-byte coalesced loads and multiplied with warp-level mma.sync.m16n8k16 (bf16 in, fp32 accumulate).
// tcgen05 tiles are 64/128 rows tall and would need block-diagonal packing of unrelated sequences.
// Everything is recomputed in the backward pass from the raw QKV; nothing but QKV is saved.
#include "common.cuh"

namespace bf {

using bf16 = __nv_bfloat16;

struct AttnParams {
  const bf16* qkv; long ld_qkv;      // (tokens, 3E): per head [q | k | v], each D wide
  bf16* out; long ld_out;            // fwd: (tokens, E).  bwd: dqkv (tokens, 3E)
  const bf16* dout; long ld_dout;    // bwd: (tokens, E)
  int heads, L;
  int G;                             // sequences packed per warp tile (block-diagonal attention), G*L <= LP
  long n_seq;
  long inner;                        // seq -> base token = (seq / inner) * outer_stride + (seq % inner) * inner_stride
  long outer_stride, inner_stride, tok_stride;
  const float* qn_w; const float* qn_b; const float* kn_w; const float* kn_b;   // [D]
  const float* bias_emb;             // [32][heads]
  const int* bucket;                 // [2L-1]: T5 bucket of rel = j - i, index rel + L - 1
  const float* scale_factor;         // [heads] or null
  float out_scale;                   // fwd: multiplies the output; bwd: multiplies dout
  int accumulate;                    // add into the destination instead of overwriting
  // bwd parameter gradients (fp32, atomically accumulated)
  float* d_qn_w; float* d_qn_b; float* d_kn_w; float* d_kn_b; float* d_bias_emb; float* d_scale_factor;
};

// ---------------------------------------------------------------------------------------------
// warp-level MMA helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const bf16* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const bf16* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// A fragment of the 16x16 block at (m0, k0).  TRANS=false: S stored [m][k]; true: S stored [k][m].
template <bool TRANS>
__device__ __forceinline__ void load_a(uint32_t (&a)[4], const bf16* S, int ld, int m0, int k0, int lane) {
  if (!TRANS) ldsm_x4(a, S + (m0 + (lane & 15)) * ld + k0 + (lane >> 4) * 8);
  else        ldsm_x4_t(a, S + (k0 + (lane & 7) + (lane >> 4) * 8) * ld + m0 + ((lane >> 3) & 1) * 8);
}
// B fragments of two adjacent 8-wide n tiles (n0, n0+8) at k0: b[0..1] tile 0, b[2..3] tile 1.
// TRANS=false: S stored [n][k]; true: S stored [k][n].
template <bool TRANS>
__device__ __forceinline__ void load_b2(uint32_t (&b)[4], const bf16* S, int ld, int n0, int k0, int lane) {
  if (!TRANS) ldsm_x4(b, S + (n0 + (lane & 7) + (lane >> 4) * 8) * ld + k0 + ((lane >> 3) & 1) * 8);
  else        ldsm_x4_t(b, S + (k0 + (lane & 7) + ((lane >> 3) & 1) * 8) * ld + n0 + (lane >> 4) * 8);
}
// acc[NT][4] (16 x NT*8) += A(16 x K) * B(K x NT*8)
template <int NT, int KSTEPS, bool A_T, bool B_T>
__device__ __forceinline__ void warp_gemm(float (&acc)[NT][4], const bf16* A, int lda, int m0, const bf16* B, int ldb,
                                          int lane) {
#pragma unroll
  for (int ks = 0; ks < KSTEPS; ++ks) {
    uint32_t a[4];
    load_a<A_T>(a, A, lda, m0, ks * 16, lane);
#pragma unroll
    for (int np = 0; np < NT / 2; ++np) {
      uint32_t b[4];
      load_b2<B_T>(b, B, ldb, np * 16, ks * 16, lane);
      mma_bf16(acc[2 * np], a, b[0], b[1]);
      mma_bf16(acc[2 * np + 1], a, b[2], b[3]);
    }
  }
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
  return v;
}
// sum over the 8 row-groups (lanes with equal lane%4)
__device__ __forceinline__ float col_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 16);
  return v;
}

template <int D, int LP>
struct AttnSmem {
  static constexpr int DS = D + 8;
  static constexpr int LS = LP + 8;
  static constexpr int kTile = LP * DS;          // elements of one (LP x D) tile
  static constexpr int kSq = LP * LS;            // elements of one (LP x LP) tile
  static constexpr int kMaxG = 8;                // sequences per tile
  // per-warp tables: meanv/cdo[kMaxG][D], brel[2*LP], stat[2][LP][2] (fp32), rowtok[LP] (int64), rowgp[LP] (int32)
  static constexpr int kMiscBytes = (kMaxG * D + 2 * LP + 4 * LP) * 4 + LP * 8 + LP * 4;
  // forward: q, k, v, p (bf16)
  static constexpr int kFwdBytes = (3 * kTile + kSq) * 2 + kMiscBytes;
  // backward: q, k, v, do (bf16), p, ds (bf16)
  static constexpr int kBwdBytes = (4 * kTile + 2 * kSq) * 2 + kMiscBytes;
};

__device__ __forceinline__ long seq_base(const AttnParams& p, long seq) {
  return (seq / p.inner) * p.outer_stride + (seq % p.inner) * p.inner_stride;
}

// Row tables of one work item: tile row r holds token rowtok[r] (or -1), belonging to packed sequence
// rowgp[r] >> 8 at position rowgp[r] & 255 (group 255 = unused row).
template <int LP>
__device__ __forceinline__ void fill_row_tables(const AttnParams& p, long seq0, long* rowtok, int* rowgp, int lane) {
  for (int r = lane; r < LP; r += 32) {
    const int g = r / p.L, i = r - g * p.L;
    const long sq = seq0 + g;
    const bool ok = g < p.G && sq < p.n_seq;
    rowtok[r] = ok ? seq_base(p, sq) + (long)i * p.tok_stride : -1;
    rowgp[r] = ok ? ((g << 8) | i) : (255 << 8);
  }
}

// Stage raw q, k, v rows of one (tile, head) into shared memory; unused rows are zeroed.
template <int D, int LP>
__device__ __forceinline__ void stage_qkv(const AttnParams& p, const long* rowtok, int head, bf16* sQ, bf16* sK, bf16* sV,
                                          int lane) {
  constexpr int DS = D + 8;
  constexpr int CPR = 3 * D / 8;                   // 16-byte chunks per row
  const bf16* src = p.qkv + (long)head * 3 * D;
  for (int idx = lane; idx < LP * CPR; idx += 32) {
    const int r = idx / CPR, ch = idx - r * CPR;
    const int which = ch / (D / 8), off = (ch - which * (D / 8)) * 8;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    const long tok = rowtok[r];
    if (tok >= 0) v = *reinterpret_cast<const uint4*>(src + tok * p.ld_qkv + ch * 8);
    bf16* dst = (which == 0 ? sQ : (which == 1 ? sK : sV)) + r * DS + off;
    *reinterpret_cast<uint4*>(dst) = v;
  }
}

// LayerNorm over D of the q rows then the k rows in place (4 lanes per row); q is also multiplied by
// D^-1/2.  Row statistics (mean, rstd) of the raw rows go to stat[which][row][2].
template <int D, int LP>
__device__ __forceinline__ void layernorm_qk(const AttnParams& p, bf16* sQ, bf16* sK, float* stat, const long* rowtok,
                                             int lane) {
  constexpr int DS = D + 8;
  constexpr int EPL = D / 4;                       // elements per lane
  const float qscale = rsqrtf((float)D);
  const int sub = lane & 3;
  for (int it = 0; it < (2 * LP) / 8; ++it) {
    const int rid = it * 8 + (lane >> 2);
    const int which = rid >= LP ? 1 : 0;
    const int r = rid - which * LP;
    bf16* row = (which ? sK : sQ) + r * DS + sub * EPL;
    const float* w = (which ? p.kn_w : p.qn_w) + sub * EPL;
    const float* b = (which ? p.kn_b : p.qn_b) + sub * EPL;
    float v[EPL];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < EPL; j += 2) {
      float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(row + j));
      v[j] = f.x; v[j + 1] = f.y;
      s += f.x + f.y;
    }
    const float mean = quad_sum(s) * (1.f / D);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < EPL; ++j) { const float d = v[j] - mean; q = fmaf(d, d, q); }
    const float rstd = rsqrtf(quad_sum(q) * (1.f / D) + 1e-5f);
    if (sub == 0) { stat[(which * LP + r) * 2] = mean; stat[(which * LP + r) * 2 + 1] = rstd; }
    const float post = which ? 1.f : qscale;
    if (rowtok[r] >= 0) {
#pragma unroll
      for (int j = 0; j < EPL; j += 2) {
        const float y0 = ((v[j] - mean) * rstd * __ldg(w + j) + __ldg(b + j)) * post;
        const float y1 = ((v[j + 1] - mean) * rstd * __ldg(w + j + 1) + __ldg(b + j + 1)) * post;
        *reinterpret_cast<__nv_bfloat162*>(row + j) = __floats2bfloat162_rn(y0, y1);
      }
    }
  }
}

// scores of one 16-row tile -> probabilities in place (fp32 fragments)
template <int LP>
__device__ __forceinline__ void softmax_tile(float (&acc)[LP / 8][4], const float* brel, const int* rowgp, int L, int mt,
                                             int lane) {
  constexpr int NT = LP / 8;
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int i = mt * 16 + g + half * 8;
    const int gi = rowgp[i];
    float mx = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = nt * 8 + 2 * t + e;
        const int gj = rowgp[j];
        float s = acc[nt][half * 2 + e];
        if ((gi >> 8) == 255) s = 0.f;                                   // unused query row: harmless uniform row
        else if ((gj >> 8) == (gi >> 8)) s += brel[(gj & 255) - (gi & 255) + L - 1];
        else s = -INFINITY;                                              // other sequence / padding
        acc[nt][half * 2 + e] = s;
        mx = fmaxf(mx, s);
      }
    }
    mx = quad_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float pe = __expf(acc[nt][half * 2 + e] - mx);
        acc[nt][half * 2 + e] = pe;
        sum += pe;
      }
    }
    const float inv = 1.f / quad_sum(sum);
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      acc[nt][half * 2] *= inv;
      acc[nt][half * 2 + 1] *= inv;
    }
  }
}

template <int NT>
__device__ __forceinline__ void store_tile_bf16(bf16* S, int ld, int m0, const float (&acc)[NT][4], int lane) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    *reinterpret_cast<__nv_bfloat162*>(S + (m0 + g) * ld + nt * 8 + 2 * t) = __floats2bfloat162_rn(acc[nt][0], acc[nt][1]);
    *reinterpret_cast<__nv_bfloat162*>(S + (m0 + g + 8) * ld + nt * 8 + 2 * t) = __floats2bfloat162_rn(acc[nt][2], acc[nt][3]);
  }
}

// coalesced copy of L rows x D columns from shared memory to a strided global destination
template <int D>
__device__ __forceinline__ void store_rows(const bf16* S, int ld, bf16* dst, long ld_dst, const long* rowtok, int rows,
                                           int accumulate, int lane) {
  constexpr int CPR = D / 8;
  for (int idx = lane; idx < rows * CPR; idx += 32) {
    const int r = idx / CPR, ch = idx - r * CPR;
    const long tok = rowtok[r];
    if (tok < 0) continue;
    uint4 v = *reinterpret_cast<const uint4*>(S + r * ld + ch * 8);
    bf16* gp = dst + tok * ld_dst + ch * 8;
    if (accumulate) {
      uint4 o = *reinterpret_cast<const uint4*>(gp);
      uint32_t* vv = reinterpret_cast<uint32_t*>(&v);
      const uint32_t* oo = reinterpret_cast<const uint32_t*>(&o);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float2 a = unpack2<bf16>(vv[k]), b = unpack2<bf16>(oo[k]);
        vv[k] = pack2<bf16>(a.x + b.x, a.y + b.y);
      }
    }
    *reinterpret_cast<uint4*>(gp) = v;
  }
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
template <int D, int LP, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
attn_fwd_kernel(AttnParams p) {
  pdl_prologue_done();
  using SM = AttnSmem<D, LP>;
  constexpr int DS = SM::DS, LS = SM::LS, NT = LP / 8;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* my = smem_raw + (size_t)warp * SM::kFwdBytes;
  bf16* sQ = reinterpret_cast<bf16*>(my);
  bf16* sK = sQ + SM::kTile;
  bf16* sV = sK + SM::kTile;
  bf16* sP = sV + SM::kTile;
  long* rowtok = reinterpret_cast<long*>(sP + SM::kSq);
  float* meanv = reinterpret_cast<float*>(rowtok + LP);     // [G][D]
  float* brel = meanv + SM::kMaxG * D;
  float* stat = brel + 2 * LP;
  int* rowgp = reinterpret_cast<int*>(stat + 4 * LP);
  const int L = p.L, G = p.G;
  const int rows = G * L;
  const long n_tiles = (p.n_seq + G - 1) / G;
  const long n_work = n_tiles * p.heads;
  const int t = lane & 3, g8 = lane >> 2;

  for (long wi = (long)blockIdx.x * WARPS + warp; wi < n_work; wi += (long)gridDim.x * WARPS) {
    const long tile = wi / p.heads;
    const int head = (int)(wi - tile * p.heads);
    fill_row_tables<LP>(p, tile * G, rowtok, rowgp, lane);
    for (int r = lane; r < 2 * L - 1; r += 32) brel[r] = __ldg(p.bias_emb + __ldg(p.bucket + r) * p.heads + head);
    __syncwarp();
    stage_qkv<D, LP>(p, rowtok, head, sQ, sK, sV, lane);
    __syncwarp();
    layernorm_qk<D, LP>(p, sQ, sK, stat, rowtok, lane);
    for (int idx = lane; idx < G * D; idx += 32) {
      const int gq = idx / D, c = idx - gq * D;
      float s = 0.f;
      for (int i = 0; i < L; ++i) s += __bfloat162float(sV[(gq * L + i) * DS + c]);
      meanv[idx] = s / (float)L;
    }
    __syncwarp();
    const float sf = p.scale_factor != nullptr ? __ldg(p.scale_factor + head) : 1.f;
    for (int mt = 0; mt * 16 < rows; ++mt) {
      float acc[NT][4];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) { acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f; }
      warp_gemm<NT, D / 16, false, false>(acc, sQ, DS, mt * 16, sK, DS, lane);
      softmax_tile<LP>(acc, brel, rowgp, L, mt, lane);
      store_tile_bf16<NT>(sP, LS, mt * 16, acc, lane);
      __syncwarp();
      float o[D / 8][4];
#pragma unroll
      for (int nt = 0; nt < D / 8; ++nt) { o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f; }
      warp_gemm<D / 8, LP / 16, false, true>(o, sP, LS, mt * 16, sV, DS, lane);
      // attn = 1/L + (P - 1/L) * s  =>  out = s * (P V) + (1 - s) * mean_L(V)
      const int ga = min(rowgp[mt * 16 + g8] >> 8, G - 1), gb = min(rowgp[mt * 16 + g8 + 8] >> 8, G - 1);
#pragma unroll
      for (int nt = 0; nt < D / 8; ++nt) {
        const int c = nt * 8 + 2 * t;
        o[nt][0] = (sf * o[nt][0] + (1.f - sf) * meanv[ga * D + c]) * p.out_scale;
        o[nt][1] = (sf * o[nt][1] + (1.f - sf) * meanv[ga * D + c + 1]) * p.out_scale;
        o[nt][2] = (sf * o[nt][2] + (1.f - sf) * meanv[gb * D + c]) * p.out_scale;
        o[nt][3] = (sf * o[nt][3] + (1.f - sf) * meanv[gb * D + c + 1]) * p.out_scale;
      }
      store_tile_bf16<D / 8>(sQ, DS, mt * 16, o, lane);     // the Q rows of this tile are dead
    }
    __syncwarp();
    store_rows<D>(sQ, DS, p.out + (long)head * D, p.ld_out, rowtok, rows, p.accumulate, lane);
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------
// LayerNorm backward on one 16-row output tile held in fragments.  acc = dL/d(normalised*w+b) [* pre],
// raw rows (bf16) in sRaw, row stats in stat.  Writes dL/d(raw) into acc; accumulates dw, db.
template <int D>
__device__ __forceinline__ void ln_bwd_tile(float (&acc)[D / 8][4], float pre, const bf16* sRaw, int ld, const float* stat,
                                            int m0, const float* w, float* s_dw, float* s_db, int lane) {
  const int g = lane >> 2, t = lane & 3;
  float dwp[D / 8][2], dbp[D / 8][2];
#pragma unroll
  for (int nt = 0; nt < D / 8; ++nt) { dwp[nt][0] = dwp[nt][1] = dbp[nt][0] = dbp[nt][1] = 0.f; }
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int r = m0 + g + half * 8;
    const float mean = stat[2 * r], rstd = stat[2 * r + 1];
    float nrm[D / 8][2];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int nt = 0; nt < D / 8; ++nt) {
      const float2 raw = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(sRaw + r * ld + nt * 8 + 2 * t));
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int c = nt * 8 + 2 * t + e;
        const float n = ((e ? raw.y : raw.x) - mean) * rstd;
        const float dy = acc[nt][half * 2 + e] * pre;
        nrm[nt][e] = n;
        dwp[nt][e] = fmaf(dy, n, dwp[nt][e]);
        dbp[nt][e] += dy;
        const float dn = dy * __ldg(w + c);
        acc[nt][half * 2 + e] = dn;
        s1 += dn;
        s2 = fmaf(dn, n, s2);
      }
    }
    s1 = quad_sum(s1) * (1.f / D);
    s2 = quad_sum(s2) * (1.f / D);
#pragma unroll
    for (int nt = 0; nt < D / 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e)
        acc[nt][half * 2 + e] = rstd * (acc[nt][half * 2 + e] - s1 - nrm[nt][e] * s2);
    }
  }
#pragma unroll
  for (int nt = 0; nt < D / 8; ++nt) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const float dw = col_sum(dwp[nt][e]), db = col_sum(dbp[nt][e]);
      if (g == 0) {
        atomicAdd(s_dw + nt * 8 + 2 * t + e, dw);
        atomicAdd(s_db + nt * 8 + 2 * t + e, db);
      }
    }
  }
}

template <int D, int LP>
__device__ __forceinline__ void stage_rows(const bf16* src, long ld_src, const long* rowtok, bf16* S, float mul, int lane) {
  constexpr int DS = D + 8, CPR = D / 8;
  for (int idx = lane; idx < LP * CPR; idx += 32) {
    const int r = idx / CPR, ch = idx - r * CPR;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    const long tok = rowtok[r];
    if (tok >= 0) {
      v = *reinterpret_cast<const uint4*>(src + tok * ld_src + ch * 8);
      if (mul != 1.f) {
        uint32_t* vv = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
        for (int k = 0; k < 4; ++k) { float2 a = unpack2<bf16>(vv[k]); vv[k] = pack2<bf16>(a.x * mul, a.y * mul); }
      }
    }
    *reinterpret_cast<uint4*>(S + r * DS + ch * 8) = v;
  }
}

template <int D, int LP, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
attn_bwd_kernel(AttnParams p) {
  pdl_prologue_done();
  using SM = AttnSmem<D, LP>;
  constexpr int DS = SM::DS, LS = SM::LS, NT = LP / 8;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  // block-level parameter-gradient accumulators live after the per-warp regions
  float* s_acc = reinterpret_cast<float*>(smem_raw + (size_t)WARPS * SM::kBwdBytes);
  float* s_dqw = s_acc;            // [D]
  float* s_dqb = s_dqw + D;
  float* s_dkw = s_dqb + D;
  float* s_dkb = s_dkw + D;
  float* s_demb = s_dkb + D;       // [32 * heads]
  float* s_dsf = s_demb + 32 * p.heads;   // [heads]
  const int n_acc = 4 * D + 33 * p.heads;
  for (int i = threadIdx.x; i < n_acc; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* my = smem_raw + (size_t)warp * SM::kBwdBytes;
  bf16* sQ = reinterpret_cast<bf16*>(my);
  bf16* sK = sQ + SM::kTile;
  bf16* sV = sK + SM::kTile;
  bf16* sG = sV + SM::kTile;       // dO, later the raw k / q rows
  bf16* sP = sG + SM::kTile;
  bf16* sS = sP + SM::kSq;         // dS
  long* rowtok = reinterpret_cast<long*>(sS + SM::kSq);
  float* cdo = reinterpret_cast<float*>(rowtok + LP);    // [G][D] column sums of dO per packed sequence
  float* brel = cdo + SM::kMaxG * D;
  float* stat = brel + 2 * LP;
  int* rowgp = reinterpret_cast<int*>(stat + 4 * LP);
  const int L = p.L, G = p.G;
  const int rows = G * L;
  const float invL = 1.f / (float)L;
  const long n_tiles = (p.n_seq + G - 1) / G;
  const long n_work = n_tiles * p.heads;
  const int g8 = lane >> 2, t = lane & 3;
  const float qscale = rsqrtf((float)D);

  for (long wi = (long)blockIdx.x * WARPS + warp; wi < n_work; wi += (long)gridDim.x * WARPS) {
    const long tile = wi / p.heads;
    const int head = (int)(wi - tile * p.heads);
    fill_row_tables<LP>(p, tile * G, rowtok, rowgp, lane);
    for (int r = lane; r < 2 * L - 1; r += 32) brel[r] = __ldg(p.bias_emb + __ldg(p.bucket + r) * p.heads + head);
    __syncwarp();
    stage_qkv<D, LP>(p, rowtok, head, sQ, sK, sV, lane);
    stage_rows<D, LP>(p.dout + (long)head * D, p.ld_dout, rowtok, sG, p.out_scale, lane);
    __syncwarp();
    layernorm_qk<D, LP>(p, sQ, sK, stat, rowtok, lane);
    for (int idx = lane; idx < G * D; idx += 32) {
      const int gq = idx / D, c = idx - gq * D;
      float s = 0.f;
      for (int i = 0; i < L; ++i) s += __bfloat162float(sG[(gq * L + i) * DS + c]);
      cdo[idx] = s;
    }
    __syncwarp();
    const float sf = p.scale_factor != nullptr ? __ldg(p.scale_factor + head) : 1.f;
    float dsf = 0.f;
    // ---- phase 1: P and dS, tile by tile ----
    for (int mt = 0; mt < LP / 16; ++mt) {
      float acc[NT][4], dp[NT][4];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
        dp[nt][0] = dp[nt][1] = dp[nt][2] = dp[nt][3] = 0.f;
      }
      if (mt * 16 < rows) {
        warp_gemm<NT, D / 16, false, false>(acc, sQ, DS, mt * 16, sK, DS, lane);
        softmax_tile<LP>(acc, brel, rowgp, L, mt, lane);
        warp_gemm<NT, D / 16, false, false>(dp, sG, DS, mt * 16, sV, DS, lane);   // dP_raw = dO V^T
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int i = mt * 16 + g8 + half * 8;
          const int gi = rowgp[i] >> 8;
          float dot = 0.f;
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int j = nt * 8 + 2 * t + e;
              const bool valid = gi != 255 && (rowgp[j] >> 8) == gi;
              const float pv = valid ? acc[nt][half * 2 + e] : 0.f;
              const float d = dp[nt][half * 2 + e];
              acc[nt][half * 2 + e] = pv;               // probabilities outside the sequence's block are zero
              if (valid) dsf = fmaf(d, pv - invL, dsf);
              dot = fmaf(pv, d, dot);
            }
          }
          dot = quad_sum(dot);
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
            for (int e = 0; e < 2; ++e)
              dp[nt][half * 2 + e] = sf * acc[nt][half * 2 + e] * (dp[nt][half * 2 + e] - dot);
          }
        }
      }
      store_tile_bf16<NT>(sP, LS, mt * 16, acc, lane);
      store_tile_bf16<NT>(sS, LS, mt * 16, dp, lane);
    }
    __syncwarp();
    // ---- bias-embedding and scale-factor gradients ----
    if (p.d_bias_emb != nullptr) {
      for (int r = lane; r < 2 * L - 1; r += 32) {
        float s = 0.f;
        for (int gq = 0; gq < G; ++gq) {
          for (int i = 0; i < L; ++i) {
            const int j = i + r - (L - 1);
            if (j >= 0 && j < L) s += __bfloat162float(sS[(gq * L + i) * LS + gq * L + j]);
          }
        }
        atomicAdd(s_demb + __ldg(p.bucket + r) * p.heads + head, s);
      }
    }
    if (p.d_scale_factor != nullptr) {
      dsf = warp_sum(dsf);
      if (lane == 0) atomicAdd(s_dsf + head, dsf);
    }
    // ---- phase 2a: dV = s * P^T dO + (1 - s)/L * colsum(dO)  -> staged in sV ----
    for (int mt = 0; mt * 16 < rows; ++mt) {
      float o[D / 8][4];
#pragma unroll
      for (int nt = 0; nt < D / 8; ++nt) { o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f; }
      warp_gemm<D / 8, LP / 16, true, true>(o, sP, LS, mt * 16, sG, DS, lane);
      const int ga = min(rowgp[mt * 16 + g8] >> 8, G - 1), gb = min(rowgp[mt * 16 + g8 + 8] >> 8, G - 1);
      const float k1 = (1.f - sf) * invL;
#pragma unroll
      for (int nt = 0; nt < D / 8; ++nt) {
        const int c = nt * 8 + 2 * t;
        o[nt][0] = sf * o[nt][0] + k1 * cdo[ga * D + c]; o[nt][1] = sf * o[nt][1] + k1 * cdo[ga * D + c + 1];
        o[nt][2] = sf * o[nt][2] + k1 * cdo[gb * D + c]; o[nt][3] = sf * o[nt][3] + k1 * cdo[gb * D + c + 1];
      }
      store_tile_bf16<D / 8>(sV, DS, mt * 16, o, lane);    // V is dead once every dP tile has been formed
    }
    __syncwarp();
    store_rows<D>(sV, DS, p.out + (long)head * 3 * D + 2 * D, p.ld_out, rowtok, rows, p.accumulate, lane);
    // raw k rows -> sG (dO is dead now)
    __syncwarp();
    stage_rows<D, LP>(p.qkv + (long)head * 3 * D + D, p.ld_qkv, rowtok, sG, 1.f, lane);
    __syncwarp();
    // ---- phase 2b: dK^ = dS^T Q^  -> LN backward -> sV ----
    for (int mt = 0; mt * 16 < rows; ++mt) {
      float o[D / 8][4];
#pragma unroll
      for (int nt = 0; nt < D / 8; ++nt) { o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f; }
      warp_gemm<D / 8, LP / 16, true, true>(o, sS, LS, mt * 16, sQ, DS, lane);
      ln_bwd_tile<D>(o, 1.f, sG, DS, stat + 2 * LP, mt * 16, p.kn_w, s_dkw, s_dkb, lane);
      store_tile_bf16<D / 8>(sV, DS, mt * 16, o, lane);
    }
    __syncwarp();
    store_rows<D>(sV, DS, p.out + (long)head * 3 * D + D, p.ld_out, rowtok, rows, p.accumulate, lane);
    __syncwarp();
    stage_rows<D, LP>(p.qkv + (long)head * 3 * D, p.ld_qkv, rowtok, sG, 1.f, lane);
    __syncwarp();
    // ---- phase 2c: dQ^ = dS K^ (w.r.t. the pre-scaled q^) -> LN backward -> sV ----
    for (int mt = 0; mt * 16 < rows; ++mt) {
      float o[D / 8][4];
#pragma unroll
      for (int nt = 0; nt < D / 8; ++nt) { o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f; }
      warp_gemm<D / 8, LP / 16, false, true>(o, sS, LS, mt * 16, sK, DS, lane);
      ln_bwd_tile<D>(o, qscale, sG, DS, stat, mt * 16, p.qn_w, s_dqw, s_dqb, lane);
      store_tile_bf16<D / 8>(sV, DS, mt * 16, o, lane);
    }
    __syncwarp();
    store_rows<D>(sV, DS, p.out + (long)head * 3 * D, p.ld_out, rowtok, rows, p.accumulate, lane);
    __syncwarp();
  }
  __syncthreads();
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    atomicAdd(p.d_qn_w + i, s_dqw[i]); atomicAdd(p.d_qn_b + i, s_dqb[i]);
    atomicAdd(p.d_kn_w + i, s_dkw[i]); atomicAdd(p.d_kn_b + i, s_dkb[i]);
  }
  if (p.d_bias_emb != nullptr)
    for (int i = threadIdx.x; i < 32 * p.heads; i += blockDim.x) atomicAdd(p.d_bias_emb + i, s_demb[i]);
  if (p.d_scale_factor != nullptr)
    for (int i = threadIdx.x; i < p.heads; i += blockDim.x) atomicAdd(p.d_scale_factor + i, s_dsf[i]);
}

// ---------------------------------------------------------------------------------------------
// host dispatch
// ---------------------------------------------------------------------------------------------
template <int D, int LP, bool BWD>
static int launch_attn(const AttnParams& p, cudaStream_t st) {
  using SM = AttnSmem<D, LP>;
  constexpr int per_warp = BWD ? SM::kBwdBytes : SM::kFwdBytes;
  constexpr int budget = 220 * 1024;
  constexpr int w_fit = budget / per_warp;
  constexpr int WARPS = w_fit >= 4 ? 4 : (w_fit >= 2 ? 2 : 1);     // small blocks: several co-reside per SM
  static_assert(w_fit >= 1, "attention tile does not fit in shared memory");
  const size_t smem = (size_t)WARPS * per_warp + (BWD ? (size_t)(4 * D + 33 * p.heads) * sizeof(float) : 0);
  void (*kern)(AttnParams);
  if constexpr (BWD) kern = attn_bwd_kernel<D, LP, WARPS>;
  else kern = attn_fwd_kernel<D, LP, WARPS>;
  static bool attr_done = false;
  if (!attr_done) {
    if (int e = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024),
                           "cudaFuncSetAttribute(attention)"))
      return e;
    attr_done = true;
  }
  BF_REQUIRE(smem <= 227 * 1024, "attention: shared memory %zu too large (heads=%d)", smem, p.heads);
  const long n_work = ((p.n_seq + p.G - 1) / p.G) * p.heads;
  const int per_sm = (int)((227 * 1024) / (smem + 1024)) > 0 ? (int)((227 * 1024) / (smem + 1024)) : 1;
  long blocks = (n_work + WARPS - 1) / WARPS;
  const long cap = (long)num_sms() * (per_sm > 8 ? 8 : per_sm);
  if (blocks > cap) blocks = cap;
  launch_k(kern, dim3((unsigned)blocks), dim3(WARPS * 32), (size_t)(smem), st, p);
  count_launch();
  return check_cuda(cudaGetLastError(), BWD ? "attn_bwd_kernel launch" : "attn_fwd_kernel launch");
}

template <bool BWD>
static int dispatch_attn(int D, int LP, const AttnParams& p, cudaStream_t st) {
#define BF_CASE(D_, LP_) if (D == D_ && LP == LP_) return launch_attn<D_, LP_, BWD>(p, st);
  BF_CASE(32, 32) BF_CASE(32, 64) BF_CASE(32, 128)
  BF_CASE(48, 32) BF_CASE(48, 64) BF_CASE(48, 128)
  BF_CASE(64, 32) BF_CASE(64, 64) BF_CASE(64, 128)
  BF_CASE(96, 32) BF_CASE(96, 64) BF_CASE(96, 128)
  BF_CASE(128, 32) BF_CASE(128, 64) BF_CASE(128, 128)
#undef BF_CASE
  set_error("bf_attention: unsupported head_dim %d (supported: 32, 48, 64, 96, 128)", D);
  return BF_ERR_INVALID;
}

}  // namespace bf

using namespace bf;

static int attn_common(const bf_attn_args* a, AttnParams& p, int& LP, bool bwd) {
  BF_REQUIRE(a != nullptr, "bf_attention: null args");
  BF_REQUIRE(a->qkv && a->out, "bf_attention: null tensor");
  BF_REQUIRE(a->heads > 0 && a->heads <= 64 && a->head_dim > 0, "bf_attention: heads=%d head_dim=%d", a->heads, a->head_dim);
  BF_REQUIRE(a->L >= 1 && a->L <= 128,
             "bf_attention: sequence length %d not in [1, 128] (one warp holds a whole sequence: axes longer than 128 "
             "tokens = 2048 pixels at patch 16 are not supported)", a->L);
  BF_REQUIRE(a->n_seq > 0 && a->inner > 0, "bf_attention: n_seq=%ld inner=%ld", (long)a->n_seq, (long)a->inner);
  BF_REQUIRE(a->qn_w && a->qn_b && a->kn_w && a->kn_b && a->bias_emb && a->bucket, "bf_attention: null parameter");
  BF_REQUIRE(a->dtype == BF_BF16 || a->dtype == BF_F32, "bf_attention: dtype %d (bf16, or fp32 for the validation backend)", a->dtype);
  BF_REQUIRE(a->dtype == BF_BF16 || !a->prenorm, "bf_attention: the pre-normalised fast path is bf16 only");
  BF_REQUIRE(a->ld_qkv % 8 == 0 && a->ld_out % 8 == 0, "bf_attention: leading dimensions must be multiples of 8");
  BF_REQUIRE((reinterpret_cast<uintptr_t>(a->qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->out) & 15) == 0,
             "bf_attention: tensors must be 16-byte aligned");
  if (bwd) {
    BF_REQUIRE(a->dout && a->ld_dout % 8 == 0, "bf_attention_bwd: dout");
    BF_REQUIRE(a->d_qn_w && a->d_qn_b && a->d_kn_w && a->d_kn_b, "bf_attention_bwd: LayerNorm gradient buffers");
  }
  // short sequences are packed G per tile (block-diagonal attention): L <= 16 -> 32-row tiles holding 32/L of them
  LP = a->L <= 32 ? 32 : (a->L <= 64 ? 64 : 128);
  p = AttnParams{};
  p.G = LP / a->L;
  if (p.G > 8) p.G = 8;
  if (p.G < 1) p.G = 1;
  p.qkv = static_cast<const bf16*>(a->qkv); p.ld_qkv = a->ld_qkv;
  p.out = static_cast<bf16*>(a->out); p.ld_out = a->ld_out;
  p.dout = static_cast<const bf16*>(a->dout); p.ld_dout = a->ld_dout;
  p.heads = a->heads; p.L = a->L; p.n_seq = a->n_seq; p.inner = a->inner;
  p.outer_stride = a->outer_stride; p.inner_stride = a->inner_stride; p.tok_stride = a->tok_stride;
  p.qn_w = a->qn_w; p.qn_b = a->qn_b; p.kn_w = a->kn_w; p.kn_b = a->kn_b;
  p.bias_emb = a->bias_emb; p.bucket = a->bucket; p.scale_factor = a->scale_factor;
  p.out_scale = a->out_scale; p.accumulate = a->accumulate;
  p.d_qn_w = a->d_qn_w; p.d_qn_b = a->d_qn_b; p.d_kn_w = a->d_kn_w; p.d_kn_b = a->d_kn_b;
  p.d_bias_emb = a->d_bias_emb; p.d_scale_factor = a->d_scale_factor;
  return BF_OK;
}

namespace bf {
int launch_attn_fast(const bf_attn_args* a, bool bwd, cudaStream_t st);
int launch_attn_fast64(const bf_attn_args* a, bool bwd, cudaStream_t st);
int launch_attn_f32(const bf_attn_args* a, bool bwd, cudaStream_t st);
}

extern "C" int bf_attention_fwd(const bf_attn_args* a, void* stream) {
  AttnParams p; int LP;
  if (int st = attn_common(a, p, LP, false)) return st;
  if (a->dtype == BF_F32) return launch_attn_f32(a, false, static_cast<cudaStream_t>(stream));
  if (a->prenorm)
    return a->L > 32 ? launch_attn_fast64(a, false, static_cast<cudaStream_t>(stream))
                     : launch_attn_fast(a, false, static_cast<cudaStream_t>(stream));
  return dispatch_attn<false>(a->head_dim, LP, p, static_cast<cudaStream_t>(stream));
}

extern "C" int bf_attention_bwd(const bf_attn_args* a, void* stream) {
  AttnParams p; int LP;
  if (int st = attn_common(a, p, LP, true)) return st;
  if (a->dtype == BF_F32) return launch_attn_f32(a, true, static_cast<cudaStream_t>(stream));
  if (a->prenorm)
    return a->L > 32 ? launch_attn_fast64(a, true, static_cast<cudaStream_t>(stream))
                     : launch_attn_fast(a, true, static_cast<cudaStream_t>(stream));
  return dispatch_attn<true>(a->head_dim, LP, p, static_cast<cudaStream_t>(stream));
}
