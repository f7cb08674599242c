// Fast path of the fused 1-D attention (head_dim 64, sequence length <= 32) over PRE-NORMALISED q / k.
//
// The QKV projection (bf_gemm, BF_EPI_QKV_LN) already stored xhat_q, xhat_k (LayerNorm without its affine part)
// and v, plus rstd of the raw rows.  What is left of upstream layers/attention.py:80-101 / :212-238 / :258-277
// per (sequence, head) is tiny and HBM bound, so the kernels are built for low instruction count and many
// resident warps:
//   * one warp owns one (tile of <= 32 tokens, head); its q|k|v rows (384 B each, contiguous in the token-major
//     QKV matrix whatever the axis) arrive by one 1-D bulk TMA copy per row (each lane issues its own row) on a
//     per-warp mbarrier -- no staging through registers, no per-element copy instructions;
//   * the LayerNorm affine (and d^-1/2) is applied to the ldmatrix fragments with packed bf16 FMAs;
//   * attn = 1/L + (softmax - 1/L)*s  is folded into the probabilities (P' = s*P + (1-s)/L), so attn @ v is a
//     single MMA and the mean-of-v term disappears; P stays in registers between the two MMAs;
//   * backward keeps dP for the whole tile in registers, parks P' / dS in the (then dead) v columns of the
//     tile, and accumulates LayerNorm weight gradients in registers across all items of the persistent loop.
// Short sequences (temporal attention, L = 5) are packed G = 32 / L per tile with a block-diagonal mask.
#include "common.cuh"

namespace bf {

using bf16 = __nv_bfloat16;

constexpr int FD = 64;                       // head dim
constexpr int FLP = 32;                      // rows per tile
constexpr int kRS = 3 * FD * 2 + 16;         // bytes per q|k|v row in shared memory (+16: conflict-free ldmatrix)
constexpr int kRSe = kRS / 2;                // same in elements
constexpr int kRG = FD * 2 + 16;             // bytes per dO row
constexpr int kRGe = kRG / 2;
constexpr int kFwdWarps = 16;                // one CTA per SM: per-warp tiles fill the shared memory
constexpr int kBwdWarps = 12;

struct FastParams {
  const bf16* qkv; long ld_qkv;
  bf16* out; long ld_out;
  const bf16* dout; long ld_dout;
  const float* rstd;                 // (tokens, heads, 2)
  int heads, L, G;
  long n_seq, inner, outer_stride, inner_stride, tok_stride;
  const float* qn_w; const float* qn_b; const float* kn_w; const float* kn_b;
  const float* bias_emb; const int* bucket; const float* scale_factor;
  float out_scale; int accumulate;
  float* d_qn_w; float* d_qn_b; float* d_kn_w; float* d_kn_b; float* d_bias_emb; float* d_scale_factor;
  float* d_qkv_bias;                 // [heads * 3 * 64] or null: += column sums of the dqkv written by this launch
};

// per-warp shared memory
struct FwdWarp {
  static constexpr int kTile = 0;                          // [32][kRS]
  static constexpr int kBar = FLP * kRS;                   // mbarrier (8 B, 16 aligned)
  static constexpr int kRowTok = kBar + 16;                // long[32]
  static constexpr int kRowGp = kRowTok + FLP * 8;         // int[32]
  static constexpr int kBrel = kRowGp + FLP * 4;           // float[64]
  static constexpr int kBytes = ((kBrel + 64 * 4) + 127) / 128 * 128;
};
struct BwdWarp {
  static constexpr int kTile = 0;                          // [32][kRS]
  static constexpr int kDo = FLP * kRS;                    // [32][kRG]
  static constexpr int kBar = kDo + FLP * kRG;
  static constexpr int kRowTok = kBar + 16;
  static constexpr int kRowGp = kRowTok + FLP * 8;
  static constexpr int kBrel = kRowGp + FLP * 4;           // float[64]
  static constexpr int kRstd = kBrel + 64 * 4;             // float[32][2]
  static constexpr int kBytes = ((kRstd + FLP * 8) + 127) / 128 * 128;
};
// CTA-level tables: affine parts of the two LayerNorms as packed bf16 (pairs for the K-contiguous fragments,
// splats for the transposed fragments) and fp32 weights for the backward
constexpr int kTabPairs = 0;        // uint32[4][32]: aq, bq, ak, bk  (pair i = columns 2i, 2i+1)
constexpr int kTabSplat = 512;      // uint32[4][64]: aq, bq, ak, bk  (both halves = column i)
constexpr int kTabW = 512 + 1024;   // float[2][64]: wq * d^-1/2 ... see kernels
constexpr int kTabBytes = 512 + 1024 + 512;

__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldsm4t(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t hfma2_bf16(uint32_t x, uint32_t a, uint32_t b) {
  uint32_t r;
  asm("fma.rn.bf16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ uint32_t pack_bf2(float lo, float hi) { return pack2<bf16>(lo, hi); }

// 1-D bulk TMA copy global -> shared with mbarrier completion (bytes: multiple of 16, both sides 16 B aligned)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ float qsum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}
__device__ __forceinline__ float qmax(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
  return v;
}

// A fragment (16 x 16 at rows m0, K columns k0) of a row-major [row][k] tile; `base` points at column 0 of row 0
__device__ __forceinline__ void frag_a(uint32_t (&a)[4], const uint8_t* base, int row_bytes, int m0, int k0, int lane) {
  ldsm4(a, base + (m0 + (lane & 15)) * row_bytes + (k0 + (lane >> 4) * 8) * 2);
}
// A fragment of the TRANSPOSE of a [k][m] tile (A[m][k] = T[k][m])
__device__ __forceinline__ void frag_a_t(uint32_t (&a)[4], const uint8_t* base, int row_bytes, int m0, int k0, int lane) {
  ldsm4t(a, base + (k0 + (lane & 7) + (lane >> 4) * 8) * row_bytes + (m0 + ((lane >> 3) & 1) * 8) * 2);
}
// B fragments of two adjacent n tiles (n0, n0 + 8) at k0 from a [n][k] tile: b[0..1] tile 0, b[2..3] tile 1
__device__ __forceinline__ void frag_b(uint32_t (&b)[4], const uint8_t* base, int row_bytes, int n0, int k0, int lane) {
  ldsm4(b, base + (n0 + (lane & 7) + (lane >> 4) * 8) * row_bytes + (k0 + ((lane >> 3) & 1) * 8) * 2);
}
// ... from a [k][n] tile
__device__ __forceinline__ void frag_b_t(uint32_t (&b)[4], const uint8_t* base, int row_bytes, int n0, int k0, int lane) {
  ldsm4t(b, base + (k0 + (lane & 7) + ((lane >> 3) & 1) * 8) * row_bytes + (n0 + (lane >> 4) * 8) * 2);
}

__device__ __forceinline__ void fill_tables(const FastParams& p, uint8_t* tab, int tid, int nthreads) {
  uint32_t* pairs = reinterpret_cast<uint32_t*>(tab + kTabPairs);
  uint32_t* splat = reinterpret_cast<uint32_t*>(tab + kTabSplat);
  float* wf = reinterpret_cast<float*>(tab + kTabW);
  const float qs = rsqrtf((float)FD);
  for (int i = tid; i < 32; i += nthreads) {
    pairs[i] = pack_bf2(p.qn_w[2 * i] * qs, p.qn_w[2 * i + 1] * qs);
    pairs[32 + i] = pack_bf2(p.qn_b[2 * i] * qs, p.qn_b[2 * i + 1] * qs);
    pairs[64 + i] = pack_bf2(p.kn_w[2 * i], p.kn_w[2 * i + 1]);
    pairs[96 + i] = pack_bf2(p.kn_b[2 * i], p.kn_b[2 * i + 1]);
  }
  for (int i = tid; i < 64; i += nthreads) {
    splat[i] = pack_bf2(p.qn_w[i] * qs, p.qn_w[i] * qs);
    splat[64 + i] = pack_bf2(p.qn_b[i] * qs, p.qn_b[i] * qs);
    splat[128 + i] = pack_bf2(p.kn_w[i], p.kn_w[i]);
    splat[192 + i] = pack_bf2(p.kn_b[i], p.kn_b[i]);
    wf[i] = p.qn_w[i];
    wf[64 + i] = p.kn_w[i];
  }
}

// Row tables + bias vector of one work item, then the bulk loads.  Returns after the data has landed.
template <bool BWD>
__device__ __forceinline__ void load_item(const FastParams& p, long tile, int head, uint8_t* my, uint64_t* bar,
                                          uint32_t& phase, long* rowtok, int* rowgp, float* brel, float* rstd_s, int lane) {
  const int L = p.L, G = p.G;
  {
    const int g = lane / L, i = lane - g * L;
    const long sq = tile * G + g;
    const bool ok = g < G && sq < p.n_seq;
    long tok = -1;
    if (ok) {
      const unsigned long o = (unsigned long)sq / (unsigned long)p.inner;
      const long in = sq - (long)o * p.inner;
      tok = (long)o * p.outer_stride + in * p.inner_stride + (long)i * p.tok_stride;
    }
    rowtok[lane] = tok;
    rowgp[lane] = ok ? ((g << 8) | i) : (255 << 8);
    for (int r = lane; r < 2 * L - 1; r += 32) brel[r] = __ldg(p.bias_emb + __ldg(p.bucket + r) * p.heads + head);
    const long seqs = min((long)G, p.n_seq - tile * G);
    const uint32_t rows = (uint32_t)(seqs * L);
    if (lane == 0) mbar_arrive_expect_tx(bar, rows * (uint32_t)((BWD ? 4 : 3) * FD * 2));
    __syncwarp();
    uint8_t* dst = my + lane * kRS;
    if (tok >= 0) {
      bulk_g2s(dst, p.qkv + tok * p.ld_qkv + (long)head * 3 * FD, 3 * FD * 2, bar);
      if (BWD) {
        bulk_g2s(my + BwdWarp::kDo + lane * kRG, p.dout + tok * p.ld_dout + (long)head * FD, FD * 2, bar);
        const float2 rr = __ldg(reinterpret_cast<const float2*>(p.rstd + (tok * p.heads + head) * 2));
        rstd_s[2 * lane] = rr.x; rstd_s[2 * lane + 1] = rr.y;
      }
    } else {
      const uint4 z = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
      for (int c = 0; c < 3 * FD * 2 / 16; ++c) *reinterpret_cast<uint4*>(dst + 16 * c) = z;
      if (BWD) {
#pragma unroll
        for (int c = 0; c < FD * 2 / 16; ++c) *reinterpret_cast<uint4*>(my + BwdWarp::kDo + lane * kRG + 16 * c) = z;
        rstd_s[2 * lane] = 0.f; rstd_s[2 * lane + 1] = 0.f;
      }
    }
  }
  mbar_wait(bar, phase);
  phase ^= 1u;
  __syncwarp();
}

// scores of one 16-row tile (already = q'k'^T) -> + bias, mask -> softmax probabilities in place
template <bool PACKED>
__device__ __forceinline__ void softmax16(float (&acc)[4][4], const float* brel, const int* rowgp, int L, int m0, int lane) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int i = m0 + g + half * 8;
    const int gi = rowgp[i];
    float mx = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = nt * 8 + 2 * t + e;
        float s = acc[nt][half * 2 + e];
        if (PACKED) {
          const int gj = rowgp[j];
          if ((gi >> 8) == 255) s = 0.f;                                  // unused query row: harmless uniform row
          else if ((gj >> 8) == (gi >> 8)) s += brel[(gj & 255) - (gi & 255) + L - 1];
          else s = -INFINITY;
        } else {
          s += brel[j - i + FLP - 1];
        }
        acc[nt][half * 2 + e] = s;
        mx = fmaxf(mx, s);
      }
    }
    mx = qmax(mx);
    float sum = 0.f;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float pe = __expf(acc[nt][half * 2 + e] - mx);
        acc[nt][half * 2 + e] = pe;
        sum += pe;
      }
    }
    const float inv = __fdividef(1.f, qsum(sum));
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) { acc[nt][half * 2] *= inv; acc[nt][half * 2 + 1] *= inv; }
  }
}

// S(16 x 32) = Q'[m0:m0+16] K'^T with the LayerNorm affine applied to the fragments
__device__ __forceinline__ void scores16(float (&acc)[4][4], const uint8_t* tile, const uint32_t* pairs, int m0, int lane) {
  const int t = lane & 3;
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) { acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f; }
#pragma unroll
  for (int ks = 0; ks < FD / 16; ++ks) {
    const uint32_t aq0 = pairs[ks * 8 + t], aq1 = pairs[ks * 8 + 4 + t];
    const uint32_t bq0 = pairs[32 + ks * 8 + t], bq1 = pairs[32 + ks * 8 + 4 + t];
    const uint32_t ak0 = pairs[64 + ks * 8 + t], ak1 = pairs[64 + ks * 8 + 4 + t];
    const uint32_t bk0 = pairs[96 + ks * 8 + t], bk1 = pairs[96 + ks * 8 + 4 + t];
    uint32_t a[4];
    frag_a(a, tile, kRS, m0, ks * 16, lane);
    a[0] = hfma2_bf16(a[0], aq0, bq0); a[1] = hfma2_bf16(a[1], aq0, bq0);
    a[2] = hfma2_bf16(a[2], aq1, bq1); a[3] = hfma2_bf16(a[3], aq1, bq1);
#pragma unroll
    for (int np = 0; np < 2; ++np) {
      uint32_t b[4];
      frag_b(b, tile + FD * 2, kRS, np * 16, ks * 16, lane);
      b[0] = hfma2_bf16(b[0], ak0, bk0); b[1] = hfma2_bf16(b[1], ak1, bk1);
      b[2] = hfma2_bf16(b[2], ak0, bk0); b[3] = hfma2_bf16(b[3], ak1, bk1);
      mma16816(acc[2 * np], a, b[0], b[1]);
      mma16816(acc[2 * np + 1], a, b[2], b[3]);
    }
  }
}

// copy rows (64 bf16 each) staged at `stage` (row pitch `pitch` bytes) to dst[tok * ld + col0 ...]
__device__ __forceinline__ void store_rows64(const uint8_t* stage, int pitch, bf16* dst, long ld, const long* rowtok,
                                             int accumulate, int lane) {
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int idx = it * 32 + lane;
    const int r = idx >> 3, ch = idx & 7;
    const long tok = rowtok[r];
    if (tok < 0) continue;
    uint4 v = *reinterpret_cast<const uint4*>(stage + r * pitch + ch * 16);
    bf16* gp = dst + tok * ld + ch * 8;
    if (accumulate) {
      const uint4 o = *reinterpret_cast<const uint4*>(gp);
      uint32_t* vv = reinterpret_cast<uint32_t*>(&v);
      const uint32_t* oo = reinterpret_cast<const uint32_t*>(&o);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 a = unpack2<bf16>(vv[k]), b = unpack2<bf16>(oo[k]);
        vv[k] = pack2<bf16>(a.x + b.x, a.y + b.y);
      }
    }
    *reinterpret_cast<uint4*>(gp) = v;
  }
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
template <bool PACKED>
__global__ void __launch_bounds__(kFwdWarps * 32, 1)
attn_fast_fwd_kernel(FastParams p) {
  constexpr int kFastWarps = kFwdWarps;
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  fill_tables(p, smem, threadIdx.x, blockDim.x);
  uint8_t* my = smem + kTabBytes + warp * FwdWarp::kBytes;
  uint64_t* bar = reinterpret_cast<uint64_t*>(my + FwdWarp::kBar);
  long* rowtok = reinterpret_cast<long*>(my + FwdWarp::kRowTok);
  int* rowgp = reinterpret_cast<int*>(my + FwdWarp::kRowGp);
  float* brel = reinterpret_cast<float*>(my + FwdWarp::kBrel);
  if (lane == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  __syncthreads();
  const uint32_t* pairs = reinterpret_cast<const uint32_t*>(smem + kTabPairs);
  uint32_t phase = 0;
  const int L = p.L, G = p.G;
  const float invL = 1.f / (float)L;
  const long n_tiles = (p.n_seq + G - 1) / G;
  const long n_work = n_tiles * p.heads;
  const int g8 = lane >> 2, t = lane & 3;

  for (long wi = (long)blockIdx.x * kFastWarps + warp; wi < n_work; wi += (long)gridDim.x * kFastWarps) {
    const long tile = wi / p.heads;
    const int head = (int)(wi - tile * p.heads);
    load_item<false>(p, tile, head, my, bar, phase, rowtok, rowgp, brel, nullptr, lane);
    const float sf = p.scale_factor != nullptr ? __ldg(p.scale_factor + head) : 1.f;
    const float lowc = (1.f - sf) * invL;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      float acc[4][4];
      scores16(acc, my, pairs, mt * 16, lane);
      softmax16<PACKED>(acc, brel, rowgp, L, mt * 16, lane);
      // P' = s*P + (1-s)/L inside the query's own sequence, 0 elsewhere; straight into A fragments
      uint32_t pa[2][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float pv = fmaf(acc[nt][e], sf, lowc);
          if (PACKED) {
            const int i = mt * 16 + g8 + (e >> 1) * 8, j = nt * 8 + 2 * t + (e & 1);
            const int gi = rowgp[i] >> 8;
            if (gi == 255 || (rowgp[j] >> 8) != gi) pv = 0.f;
          }
          v[e] = pv;
        }
        pa[nt >> 1][(nt & 1) * 2] = pack_bf2(v[0], v[1]);
        pa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf2(v[2], v[3]);
      }
      float o[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) { o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f; }
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
#pragma unroll
        for (int np = 0; np < 4; ++np) {
          uint32_t b[4];
          frag_b_t(b, my + 2 * FD * 2, kRS, np * 16, kk * 16, lane);
          mma16816(o[2 * np], pa[kk], b[0], b[1]);
          mma16816(o[2 * np + 1], pa[kk], b[2], b[3]);
        }
      }
      // the q rows of this m tile are dead: stage the output there
      __syncwarp();
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        *reinterpret_cast<uint32_t*>(my + (mt * 16 + g8) * kRS + (nt * 8 + 2 * t) * 2) =
            pack_bf2(o[nt][0] * p.out_scale, o[nt][1] * p.out_scale);
        *reinterpret_cast<uint32_t*>(my + (mt * 16 + g8 + 8) * kRS + (nt * 8 + 2 * t) * 2) =
            pack_bf2(o[nt][2] * p.out_scale, o[nt][3] * p.out_scale);
      }
    }
    __syncwarp();
    store_rows64(my, kRS, p.out + (long)head * FD, p.ld_out, rowtok, p.accumulate, lane);
    fence_proxy_async();          // generic writes to the tile precede the next item's async-proxy writes
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------
// LayerNorm backward on a 16-row tile of gradients w.r.t. y = xhat*w + b held in C fragments.
// xhat rows are read from `xh` (row pitch kRS); rstd per row from rstd_s[row*2 + which].
// Writes d(raw) into acc, accumulates per-lane partial dw (and db when DB) over rows.
template <bool DB>
__device__ __forceinline__ void ln_bwd16(float (&acc)[8][4], float pre, const uint8_t* xh, const float* rstd_s, int which,
                                         int m0, const float* w, float (&dw)[8][2], float (&db)[8][2], int lane) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int r = m0 + g + half * 8;
    const float rstd = rstd_s[2 * r + which];
    float nrm[8][2];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float2 xv = unpack2<bf16>(*reinterpret_cast<const uint32_t*>(xh + r * kRS + (nt * 8 + 2 * t) * 2));
      const float2 wv = *reinterpret_cast<const float2*>(w + nt * 8 + 2 * t);
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float n = e ? xv.y : xv.x;
        const float dy = acc[nt][half * 2 + e] * pre;
        nrm[nt][e] = n;
        dw[nt][e] = fmaf(dy, n, dw[nt][e]);
        if (DB) db[nt][e] += dy;
        const float dn = dy * (e ? wv.y : wv.x);
        acc[nt][half * 2 + e] = dn;
        s1 += dn;
        s2 = fmaf(dn, n, s2);
      }
    }
    s1 = qsum(s1) * (1.f / FD);
    s2 = qsum(s2) * (1.f / FD);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) acc[nt][half * 2 + e] = rstd * (acc[nt][half * 2 + e] - s1 - nrm[nt][e] * s2);
    }
  }
}

// column sums of a 32-row tile held as two 16-row C-fragment sets -> shared accumulators (8 lanes per column collide)
__device__ __forceinline__ void colsum_frag(float* s_dst, const float (&o0)[8][4], const float (&o1)[8][4], int lane) {
  const int t = lane & 3;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      float v = (o0[nt][e] + o0[nt][2 + e]) + (o1[nt][e] + o1[nt][2 + e]);
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      if ((lane >> 2) == 0) atomicAdd(s_dst + nt * 8 + 2 * t + e, v);
    }
  }
}

__device__ __forceinline__ void colsum_frag1(float* s_dst, const float (&o)[8][4], int lane) {
  const int t = lane & 3;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      float v = o[nt][e] + o[nt][2 + e];
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      if ((lane >> 2) == 0) atomicAdd(s_dst + nt * 8 + 2 * t + e, v);
    }
  }
}

__device__ __forceinline__ void stage16(uint8_t* stage, int pitch, int m0, const float (&o)[8][4], int lane) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    *reinterpret_cast<uint32_t*>(stage + (m0 + g) * pitch + (nt * 8 + 2 * t) * 2) = pack_bf2(o[nt][0], o[nt][1]);
    *reinterpret_cast<uint32_t*>(stage + (m0 + g + 8) * pitch + (nt * 8 + 2 * t) * 2) = pack_bf2(o[nt][2], o[nt][3]);
  }
}

template <bool PACKED>
__global__ void __launch_bounds__(kBwdWarps * 32, 1)
attn_fast_bwd_kernel(FastParams p) {
  constexpr int kFastWarps = kBwdWarps;
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  fill_tables(p, smem, threadIdx.x, blockDim.x);
  // CTA-level gradient accumulators after the per-warp regions
  float* s_acc = reinterpret_cast<float*>(smem + kTabBytes + kFastWarps * BwdWarp::kBytes);
  float* s_dqw = s_acc;                 // [64]
  float* s_dqb = s_dqw + FD;
  float* s_dkw = s_dqb + FD;
  float* s_demb = s_dkw + FD;           // [32 * heads]
  float* s_dsf = s_demb + 32 * p.heads; // [heads]
  float* s_dbias = s_dsf + p.heads;     // [heads * 3 * 64] column sums of dq | dk | dv (input_head bias gradient)
  const int n_acc = 3 * FD + 33 * p.heads + (p.d_qkv_bias != nullptr ? 3 * FD * p.heads : 0);
  for (int i = threadIdx.x; i < n_acc; i += blockDim.x) s_acc[i] = 0.f;
  uint8_t* my = smem + kTabBytes + warp * BwdWarp::kBytes;
  uint8_t* sdo = my + BwdWarp::kDo;
  uint64_t* bar = reinterpret_cast<uint64_t*>(my + BwdWarp::kBar);
  long* rowtok = reinterpret_cast<long*>(my + BwdWarp::kRowTok);
  int* rowgp = reinterpret_cast<int*>(my + BwdWarp::kRowGp);
  float* brel = reinterpret_cast<float*>(my + BwdWarp::kBrel);
  float* rstd_s = reinterpret_cast<float*>(my + BwdWarp::kRstd);
  if (lane == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  __syncthreads();
  const uint32_t* pairs = reinterpret_cast<const uint32_t*>(smem + kTabPairs);
  const uint32_t* splat = reinterpret_cast<const uint32_t*>(smem + kTabSplat);
  const float* wq = reinterpret_cast<const float*>(smem + kTabW);
  const float* wk = wq + FD;
  uint8_t* sP = my + 2 * FD * 2;        // P' parked in the first half of the v columns, dS in the second
  uint8_t* sS = sP + FLP * 2;
  uint32_t phase = 0;
  const int L = p.L, G = p.G;
  const float invL = 1.f / (float)L;
  const long n_tiles = (p.n_seq + G - 1) / G;
  const long n_work = n_tiles * p.heads;
  const int g8 = lane >> 2, t = lane & 3;
  const float qscale = rsqrtf((float)FD);
  float dwq[8][2], dbq[8][2], dwk[8][2], dbk_unused[8][2];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) { dwq[nt][0] = dwq[nt][1] = dbq[nt][0] = dbq[nt][1] = dwk[nt][0] = dwk[nt][1] = 0.f; }

  for (long wi = (long)blockIdx.x * kFastWarps + warp; wi < n_work; wi += (long)gridDim.x * kFastWarps) {
    const long tile = wi / p.heads;
    const int head = (int)(wi - tile * p.heads);
    load_item<true>(p, tile, head, my, bar, phase, rowtok, rowgp, brel, rstd_s, lane);
    const float sf = p.scale_factor != nullptr ? __ldg(p.scale_factor + head) : 1.f;
    const float lowc = (1.f - sf) * invL;
    // ---- dP'(raw) = dO V^T for the whole tile (v is dead afterwards) ----
    float dp[2][4][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) { dp[mt][nt][0] = dp[mt][nt][1] = dp[mt][nt][2] = dp[mt][nt][3] = 0.f; }
    }
#pragma unroll
    for (int ks = 0; ks < FD / 16; ++ks) {
      uint32_t a0[4], a1[4];
      frag_a(a0, sdo, kRG, 0, ks * 16, lane);
      frag_a(a1, sdo, kRG, 16, ks * 16, lane);
#pragma unroll
      for (int np = 0; np < 2; ++np) {
        uint32_t b[4];
        frag_b(b, my + 2 * FD * 2, kRS, np * 16, ks * 16, lane);
        mma16816(dp[0][2 * np], a0, b[0], b[1]);
        mma16816(dp[0][2 * np + 1], a0, b[2], b[3]);
        mma16816(dp[1][2 * np], a1, b[0], b[1]);
        mma16816(dp[1][2 * np + 1], a1, b[2], b[3]);
      }
    }
    __syncwarp();                       // every lane is done reading v: its columns now receive P' and dS
    float dsf = 0.f;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      float acc[4][4];
      scores16(acc, my, pairs, mt * 16, lane);
      softmax16<PACKED>(acc, brel, rowgp, L, mt * 16, lane);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int i = mt * 16 + g8 + half * 8;
        const int gi = rowgp[i] >> 8;
        float dot = 0.f;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            bool valid = true;
            if (PACKED) valid = gi != 255 && (rowgp[nt * 8 + 2 * t + e] >> 8) == gi;
            const float pv = valid ? acc[nt][half * 2 + e] : 0.f;
            const float d = dp[mt][nt][half * 2 + e] * p.out_scale;
            acc[nt][half * 2 + e] = pv;
            dp[mt][nt][half * 2 + e] = d;
            if (valid) dsf = fmaf(d, pv - invL, dsf);
            dot = fmaf(pv, d, dot);
          }
        }
        dot = qsum(dot);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          float ds[2], pp[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const float pv = acc[nt][half * 2 + e];
            ds[e] = sf * pv * (dp[mt][nt][half * 2 + e] - dot);
            bool valid = true;
            if (PACKED) valid = gi != 255 && (rowgp[nt * 8 + 2 * t + e] >> 8) == gi;
            pp[e] = valid ? fmaf(pv, sf, lowc) : 0.f;
          }
          *reinterpret_cast<uint32_t*>(sP + i * kRS + (nt * 8 + 2 * t) * 2) = pack_bf2(pp[0], pp[1]);
          *reinterpret_cast<uint32_t*>(sS + i * kRS + (nt * 8 + 2 * t) * 2) = pack_bf2(ds[0], ds[1]);
        }
      }
    }
    __syncwarp();
    // ---- bias-embedding and scale-factor gradients ----
    if (p.d_bias_emb != nullptr) {
      for (int r = lane; r < 2 * L - 1; r += 32) {
        float s = 0.f;
        for (int gq = 0; gq < G; ++gq) {
          for (int i = 0; i < L; ++i) {
            const int j = i + r - (L - 1);
            if (j >= 0 && j < L) s += __bfloat162float(*reinterpret_cast<const bf16*>(sS + (gq * L + i) * kRS + (gq * L + j) * 2));
          }
        }
        atomicAdd(s_demb + __ldg(p.bucket + r) * p.heads + head, s);
      }
    }
    if (p.d_scale_factor != nullptr) {
      dsf = warp_sum(dsf);
      if (lane == 0) atomicAdd(s_dsf + head, dsf);
    }
    // ---- dV = P'^T dO' ----
    {
      float o[2][8][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) { o[mt][nt][0] = o[mt][nt][1] = o[mt][nt][2] = o[mt][nt][3] = 0.f; }
      }
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        uint32_t a0[4], a1[4];
        frag_a_t(a0, sP, kRS, 0, kk * 16, lane);
        frag_a_t(a1, sP, kRS, 16, kk * 16, lane);
#pragma unroll
        for (int np = 0; np < 4; ++np) {
          uint32_t b[4];
          frag_b_t(b, sdo, kRG, np * 16, kk * 16, lane);
          mma16816(o[0][2 * np], a0, b[0], b[1]);
          mma16816(o[0][2 * np + 1], a0, b[2], b[3]);
          mma16816(o[1][2 * np], a1, b[0], b[1]);
          mma16816(o[1][2 * np + 1], a1, b[2], b[3]);
        }
      }
      __syncwarp();                     // dO is dead: its tile becomes the output staging area
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
          for (int e = 0; e < 4; ++e) o[mt][nt][e] *= p.out_scale;
        }
        stage16(sdo, kRG, mt * 16, o[mt], lane);
      }
      if (p.d_qkv_bias != nullptr) colsum_frag(s_dbias + head * 3 * FD + 2 * FD, o[0], o[1], lane);
      __syncwarp();
      store_rows64(sdo, kRG, p.out + (long)head * 3 * FD + 2 * FD, p.ld_out, rowtok, p.accumulate, lane);
      __syncwarp();
    }
    // ---- dK' = dS^T Q'  -> LayerNorm backward -> d(raw k) ----
    {
      float o[2][8][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) { o[mt][nt][0] = o[mt][nt][1] = o[mt][nt][2] = o[mt][nt][3] = 0.f; }
      }
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        uint32_t a0[4], a1[4];
        frag_a_t(a0, sS, kRS, 0, kk * 16, lane);
        frag_a_t(a1, sS, kRS, 16, kk * 16, lane);
#pragma unroll
        for (int np = 0; np < 4; ++np) {
          uint32_t b[4];
          frag_b_t(b, my, kRS, np * 16, kk * 16, lane);      // Q as [k = i][n = d]: both halves of a register share d
          const uint32_t al = splat[np * 16 + g8], ah = splat[np * 16 + 8 + g8];
          const uint32_t bl = splat[64 + np * 16 + g8], bh = splat[64 + np * 16 + 8 + g8];
          b[0] = hfma2_bf16(b[0], al, bl); b[1] = hfma2_bf16(b[1], al, bl);
          b[2] = hfma2_bf16(b[2], ah, bh); b[3] = hfma2_bf16(b[3], ah, bh);
          mma16816(o[0][2 * np], a0, b[0], b[1]);
          mma16816(o[0][2 * np + 1], a0, b[2], b[3]);
          mma16816(o[1][2 * np], a1, b[0], b[1]);
          mma16816(o[1][2 * np + 1], a1, b[2], b[3]);
        }
      }
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        ln_bwd16<false>(o[mt], 1.f, my + FD * 2, rstd_s, 1, mt * 16, wk, dwk, dbk_unused, lane);
        stage16(sdo, kRG, mt * 16, o[mt], lane);
      }
      if (p.d_qkv_bias != nullptr) colsum_frag(s_dbias + head * 3 * FD + FD, o[0], o[1], lane);
      __syncwarp();
      store_rows64(sdo, kRG, p.out + (long)head * 3 * FD + FD, p.ld_out, rowtok, p.accumulate, lane);
      __syncwarp();
    }
    // ---- dQ' = dS K'  -> LayerNorm backward -> d(raw q) ----
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      float o[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) { o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f; }
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        uint32_t a[4];
        frag_a(a, sS, kRS, mt * 16, kk * 16, lane);
#pragma unroll
        for (int np = 0; np < 4; ++np) {
          uint32_t b[4];
          frag_b_t(b, my + FD * 2, kRS, np * 16, kk * 16, lane);
          const uint32_t al = splat[128 + np * 16 + g8], ah = splat[128 + np * 16 + 8 + g8];
          const uint32_t bl = splat[192 + np * 16 + g8], bh = splat[192 + np * 16 + 8 + g8];
          b[0] = hfma2_bf16(b[0], al, bl); b[1] = hfma2_bf16(b[1], al, bl);
          b[2] = hfma2_bf16(b[2], ah, bh); b[3] = hfma2_bf16(b[3], ah, bh);
          mma16816(o[2 * np], a, b[0], b[1]);
          mma16816(o[2 * np + 1], a, b[2], b[3]);
        }
      }
      ln_bwd16<true>(o, qscale, my, rstd_s, 0, mt * 16, wq, dwq, dbq, lane);
      stage16(sdo, kRG, mt * 16, o, lane);
      if (p.d_qkv_bias != nullptr) colsum_frag1(s_dbias + head * 3 * FD, o, lane);
    }
    __syncwarp();
    store_rows64(sdo, kRG, p.out + (long)head * 3 * FD, p.ld_out, rowtok, p.accumulate, lane);
    fence_proxy_async();
    __syncwarp();
  }
  // ---- LayerNorm parameter gradients: lanes with equal t hold partial sums of the same columns ----
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      float a = dwq[nt][e], b = dbq[nt][e], c = dwk[nt][e];
#pragma unroll
      for (int o = 4; o < 32; o <<= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
        c += __shfl_xor_sync(0xffffffffu, c, o);
      }
      if (g8 == 0) {
        atomicAdd(s_dqw + nt * 8 + 2 * t + e, a);
        atomicAdd(s_dqb + nt * 8 + 2 * t + e, b);
        atomicAdd(s_dkw + nt * 8 + 2 * t + e, c);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < FD; i += blockDim.x) {
    atomicAdd(p.d_qn_w + i, s_dqw[i]);
    atomicAdd(p.d_qn_b + i, s_dqb[i]);
    atomicAdd(p.d_kn_w + i, s_dkw[i]);
  }
  if (p.d_bias_emb != nullptr)
    for (int i = threadIdx.x; i < 32 * p.heads; i += blockDim.x) atomicAdd(p.d_bias_emb + i, s_demb[i]);
  if (p.d_scale_factor != nullptr)
    for (int i = threadIdx.x; i < p.heads; i += blockDim.x) atomicAdd(p.d_scale_factor + i, s_dsf[i]);
  if (p.d_qkv_bias != nullptr)
    for (int i = threadIdx.x; i < 3 * FD * p.heads; i += blockDim.x) atomicAdd(p.d_qkv_bias + i, s_dbias[i]);
}

// ---------------------------------------------------------------------------------------------
// host launch
// ---------------------------------------------------------------------------------------------
int launch_attn_fast(const bf_attn_args* a, bool bwd, cudaStream_t st) {
  BF_REQUIRE(a->head_dim == FD && a->L >= 1 && a->L <= FLP,
             "bf_attention (prenorm): the pre-normalised path handles head_dim 64 and L <= 32 (got d=%d L=%d)",
             a->head_dim, a->L);
  BF_REQUIRE(!bwd || a->rstd != nullptr, "bf_attention_bwd (prenorm): rstd required");
  FastParams p{};
  p.qkv = static_cast<const bf16*>(a->qkv); p.ld_qkv = a->ld_qkv;
  p.out = static_cast<bf16*>(a->out); p.ld_out = a->ld_out;
  p.dout = static_cast<const bf16*>(a->dout); p.ld_dout = a->ld_dout;
  p.rstd = a->rstd;
  p.heads = a->heads; p.L = a->L;
  p.G = FLP / a->L;
  p.n_seq = a->n_seq; p.inner = a->inner;
  p.outer_stride = a->outer_stride; p.inner_stride = a->inner_stride; p.tok_stride = a->tok_stride;
  p.qn_w = a->qn_w; p.qn_b = a->qn_b; p.kn_w = a->kn_w; p.kn_b = a->kn_b;
  p.bias_emb = a->bias_emb; p.bucket = a->bucket; p.scale_factor = a->scale_factor;
  p.out_scale = a->out_scale; p.accumulate = a->accumulate;
  p.d_qn_w = a->d_qn_w; p.d_qn_b = a->d_qn_b; p.d_kn_w = a->d_kn_w; p.d_kn_b = a->d_kn_b;
  p.d_bias_emb = a->d_bias_emb; p.d_scale_factor = a->d_scale_factor;
  p.d_qkv_bias = bwd ? a->d_qkv_bias : nullptr;
  const bool packed = !(p.G == 1 && a->L == FLP);
  const int kFastWarps = bwd ? kBwdWarps : kFwdWarps;
  const size_t smem = kTabBytes + (size_t)kFastWarps * (bwd ? BwdWarp::kBytes : FwdWarp::kBytes) +
                      (bwd ? (size_t)(3 * FD + 33 * p.heads + 3 * FD * p.heads) * sizeof(float) : 0);
  BF_REQUIRE(smem <= 227 * 1024, "bf_attention (prenorm): shared memory %zu too large", smem);
  void (*kern)(FastParams);
  if (bwd) kern = packed ? attn_fast_bwd_kernel<true> : attn_fast_bwd_kernel<false>;
  else kern = packed ? attn_fast_fwd_kernel<true> : attn_fast_fwd_kernel<false>;
  static bool attr_done[4] = {false, false, false, false};
  const int ki = (bwd ? 2 : 0) + (packed ? 1 : 0);
  if (!attr_done[ki]) {
    if (int e = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024),
                           "cudaFuncSetAttribute(attention fast)"))
      return e;
    attr_done[ki] = true;
  }
  const long n_work = ((p.n_seq + p.G - 1) / p.G) * p.heads;
  long blocks = (n_work + kFastWarps - 1) / kFastWarps;
  const long cap = (long)num_sms();
  if (blocks > cap) blocks = cap;
  kern<<<(unsigned)blocks, kFastWarps * 32, smem, st>>>(p);
  count_launch();
  return check_cuda(cudaGetLastError(), bwd ? "attn_fast_bwd_kernel launch" : "attn_fast_fwd_kernel launch");
}

}  // namespace bf
