// Fast path of the fused 1-D attention (head_dim 64, sequence length <= 32) over PRE-NORMALISED q / k.
//
// The QKV projection (bf_gemm, BF_EPI_QKV_LN) already stored xhat_q, xhat_k (LayerNorm without its affine part)
// and v, plus rstd of the raw rows.  What is left of upstream layers/attention.py:80-101 / :212-238 / :258-277
// per (sequence, head) is tiny and HBM bound, so the kernels are built around the TMA unit and a low instruction
// count:
//   * one warp owns one (tile of <= 32 tokens, head).  The token-major QKV matrix is described to TMA as a 4-D
//     tensor (column, position in sequence, sequence, outer index) whatever the axis (time / image row / image
//     column), so the q, k and v tiles of an item arrive as THREE tensor copies (128B-swizzled 32 x 64 tiles)
//     on a per-warp mbarrier, and results leave as ONE tensor store per tile (or a bf16 TMA reduce-add when a
//     second axis accumulates into the same tensor) -- no per-row copies, no read-modify-write loops;
//   * the LayerNorm affine (and d^-1/2) is applied to the ldmatrix fragments with packed bf16 FMAs;
//   * attn = 1/L + (softmax - 1/L)*s  is folded into the probabilities (P' = s*P + (1-s)/L), so attn @ v is a
//     single MMA and the mean-of-v term disappears; P stays in registers between the two MMAs;
//   * backward keeps dP for the whole tile in registers, parks P' / dS in the (then dead) v tile, and accumulates
//     LayerNorm weight gradients in registers across all items of the persistent loop.
// Short sequences (temporal attention, L = 5) are packed G = 32 / L per tile with a block-diagonal mask.
#include "attention_fast.cuh"

namespace bf {



// Row tables + bias vector of one work item, then the tensor loads.  Returns after the data has landed.
template <bool BWD>
__device__ __forceinline__ Item load_item(const FastParams& p, const CUtensorMap* map_qkv, const CUtensorMap* map_do,
                                          long wi, uint8_t* my, uint64_t* bar, uint32_t& phase, int* rowgp,
                                          float* rstd_s, int lane) {
  const int L = p.L, G = p.G;
  Item it;
  const long tile = wi / p.heads;
  it.head = (int)(wi - tile * p.heads);
  it.s_out = (int)(tile / p.tiles_per_outer);
  it.s_in0 = (int)(tile - (long)it.s_out * p.tiles_per_outer) * G;
  if (lane == 0) {
    tma_store_wait_read<0>();          // the previous item's stores have finished reading the tiles
    mbar_arrive_expect_tx(bar, (uint32_t)((BWD ? 4 : 3) * G * L * FD * 2));
    const int col = it.head * 3 * FD;
    tma_load_4d(my, map_qkv, bar, col, 0, it.s_in0, it.s_out);
    tma_load_4d(my + kTile, map_qkv, bar, col + FD, 0, it.s_in0, it.s_out);
    tma_load_4d(my + 2 * kTile, map_qkv, bar, col + 2 * FD, 0, it.s_in0, it.s_out);
    if (BWD) tma_load_4d(my + 3 * kTile, map_do, bar, it.head * FD, 0, it.s_in0, it.s_out);
  }
  const int g = lane / L, i = lane - g * L;
  const bool ok = g < G && it.s_in0 + g < p.inner;
  float2 rr = make_float2(0.f, 0.f);
  if (BWD && ok) {     // issued now, consumed after the tiles have landed
    const long tok = (long)it.s_out * p.outer_stride + (long)(it.s_in0 + g) * p.inner_stride + (long)i * p.tok_stride;
    rr = __ldg(reinterpret_cast<const float2*>(p.rstd + (tok * p.heads + it.head) * 2));
  }
  rowgp[lane] = ok ? ((g << 8) | i) : (255 << 8);
  mbar_wait(bar, phase);
  phase ^= 1u;
  if (BWD) { rstd_s[2 * lane] = rr.x; rstd_s[2 * lane + 1] = rr.y; }
  __syncwarp();
  return it;
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
__device__ __forceinline__ void scores16(float (&acc)[4][4], const uint8_t* sQ, const uint8_t* sK, const uint32_t* pairs,
                                         int m0, int lane) {
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
    frag_a(a, sQ, m0, ks * 16, lane);
    a[0] = hfma2_bf16(a[0], aq0, bq0); a[1] = hfma2_bf16(a[1], aq0, bq0);
    a[2] = hfma2_bf16(a[2], aq1, bq1); a[3] = hfma2_bf16(a[3], aq1, bq1);
#pragma unroll
    for (int np = 0; np < 2; ++np) {
      uint32_t b[4];
      frag_b(b, sK, np * 16, ks * 16, lane);
      b[0] = hfma2_bf16(b[0], ak0, bk0); b[1] = hfma2_bf16(b[1], ak1, bk1);
      b[2] = hfma2_bf16(b[2], ak0, bk0); b[3] = hfma2_bf16(b[3], ak1, bk1);
      mma16816(acc[2 * np], a, b[0], b[1]);
      mma16816(acc[2 * np + 1], a, b[2], b[3]);
    }
  }
}

// 16 x 64 C fragments -> bf16 rows m0.. of a swizzled tile
__device__ __forceinline__ void stage16(uint8_t* tile, int m0, const float (&o)[8][4], int lane) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    *reinterpret_cast<uint32_t*>(tile + swz(m0 + g, nt * 8 + 2 * t)) = pack_bf2(o[nt][0], o[nt][1]);
    *reinterpret_cast<uint32_t*>(tile + swz(m0 + g + 8, nt * 8 + 2 * t)) = pack_bf2(o[nt][2], o[nt][3]);
  }
}

// tile (G*L rows x 64 columns) -> global through the 4-D map at column `col`; add = bf16 TMA reduction
__device__ __forceinline__ void store_tile(const CUtensorMap* map, const uint8_t* tile, int col, const Item& it,
                                           int accumulate, int lane) {
  fence_proxy_async();
  __syncwarp();
  if (lane == 0) {
    if (accumulate) tma_reduce_add_4d(map, tile, col, 0, it.s_in0, it.s_out);
    else tma_store_4d(map, tile, col, 0, it.s_in0, it.s_out);
    tma_store_commit();
  }
}

__device__ __forceinline__ void zero_tiles(uint8_t* my, int bytes, int lane) {
  const uint4 z = make_uint4(0u, 0u, 0u, 0u);
  for (int i = lane * 16; i < bytes; i += 32 * 16) *reinterpret_cast<uint4*>(my + i) = z;
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
template <bool PACKED>
__global__ void __launch_bounds__(kFwdWarps * 32, 1)
attn_fast_fwd_kernel(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_out,
                     const FastParams p) {
  pdl_prologue_done();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* tab = smem + kFwdWarps * FwdWarp::kBytes;
  fill_tables(p, tab, threadIdx.x, blockDim.x);
  uint8_t* my = smem + warp * FwdWarp::kBytes;
  uint8_t* sQ = my + FwdWarp::kQ;
  uint8_t* sK = my + FwdWarp::kK;
  uint8_t* sV = my + FwdWarp::kV;
  uint64_t* bar = reinterpret_cast<uint64_t*>(my + FwdWarp::kBar);
  int* rowgp = reinterpret_cast<int*>(my + FwdWarp::kRowGp);
  const float* brel_all = reinterpret_cast<const float*>(tab + kTabBrel);
  zero_tiles(my, 3 * kTile, lane);          // rows beyond G*L are never written by TMA and must stay finite
  if (lane == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
    tma_prefetch_desc(&map_qkv);
    tma_prefetch_desc(&map_out);
  }
  fence_proxy_async();
  __syncthreads();
  const uint32_t* pairs = reinterpret_cast<const uint32_t*>(tab + kTabPairs);
  uint32_t phase = 0;
  const int L = p.L;
  const float invL = 1.f / (float)L;
  const long n_work = p.n_tiles * p.heads;
  const int g8 = lane >> 2, t = lane & 3;

  for (long wi = (long)blockIdx.x * kFwdWarps + warp; wi < n_work; wi += (long)gridDim.x * kFwdWarps) {
    const Item it = load_item<false>(p, &map_qkv, nullptr, wi, my, bar, phase, rowgp, nullptr, lane);
    const float* brel = brel_all + it.head * 64;
    const float sf = p.scale_factor != nullptr ? __ldg(p.scale_factor + it.head) : 1.f;
    const float lowc = (1.f - sf) * invL;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      float acc[4][4];
      scores16(acc, sQ, sK, pairs, mt * 16, lane);
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
          frag_b_t(b, sV, np * 16, kk * 16, lane);
          mma16816(o[2 * np], pa[kk], b[0], b[1]);
          mma16816(o[2 * np + 1], pa[kk], b[2], b[3]);
        }
      }
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) o[nt][e] *= p.out_scale;
      }
      // the q rows of this m tile are dead: stage the output there
      __syncwarp();
      stage16(sQ, mt * 16, o, lane);
    }
    store_tile(&map_out, sQ, it.head * FD, it, p.accumulate, lane);
  }
  if (lane == 0) tma_store_wait_all();
}

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------
// LayerNorm backward on a 16-row tile of gradients w.r.t. y = xhat*w + b held in C fragments.
// xhat rows are read from the tile `xh`; rstd per row from rstd_s[row*2 + which].
// Writes d(raw) into acc, accumulates per-lane partial dw (and db when DB) over rows.
template <bool DB, bool SCALE>
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
      const float2 xv = unpack2<bf16>(*reinterpret_cast<const uint32_t*>(xh + swz(r, nt * 8 + 2 * t)));
      const float2 wv = *reinterpret_cast<const float2*>(w + nt * 8 + 2 * t);
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float n = e ? xv.y : xv.x;
        const float dy = SCALE ? acc[nt][half * 2 + e] * pre : acc[nt][half * 2 + e];
        nrm[nt][e] = n;
        dw[nt][e] = fmaf(dy, n, dw[nt][e]);
        if (DB) db[nt][e] += dy;
        const float dn = dy * (e ? wv.y : wv.x);
        acc[nt][half * 2 + e] = dn;
        s1 += dn;
        s2 = fmaf(dn, n, s2);
      }
    }
    const float c0 = -rstd * qsum(s1) * (1.f / FD);
    const float c1 = -rstd * qsum(s2) * (1.f / FD);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) acc[nt][half * 2 + e] = fmaf(nrm[nt][e], c1, fmaf(acc[nt][half * 2 + e], rstd, c0));
    }
  }
}

// Column sums of a 16-row (or, two tiles at once, 32-row) tile held in C fragments -> the WARP's private accumulators in
// shared memory.  A warp keeps one head for the whole launch, every column of its table is owned by exactly one lane
// (t, nt, e), so the update is a plain read-modify-write: no atomics (shared fp32 atomics are compare-and-swap loops,
// and twelve warps contending on one table cost +45 us per launch in the first version of this fusion).
__device__ __forceinline__ void colsum_frag1(float* s_dst, const float (&o)[8][4], int lane) {
  const int t = lane & 3;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    float v0 = o[nt][0] + o[nt][2], v1 = o[nt][1] + o[nt][3];
#pragma unroll
    for (int sh = 4; sh < 32; sh <<= 1) {
      v0 += __shfl_xor_sync(0xffffffffu, v0, sh);
      v1 += __shfl_xor_sync(0xffffffffu, v1, sh);
    }
    if ((lane >> 2) == 0) {
      float2* d = reinterpret_cast<float2*>(s_dst + nt * 8 + 2 * t);
      float2 c = *d;
      c.x += v0; c.y += v1;
      *d = c;
    }
  }
}
__device__ __forceinline__ void colsum_frag2(float* s_dst, const float (&o)[2][8][4], int lane) {
  const int t = lane & 3;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    float v0 = (o[0][nt][0] + o[0][nt][2]) + (o[1][nt][0] + o[1][nt][2]);
    float v1 = (o[0][nt][1] + o[0][nt][3]) + (o[1][nt][1] + o[1][nt][3]);
#pragma unroll
    for (int sh = 4; sh < 32; sh <<= 1) {
      v0 += __shfl_xor_sync(0xffffffffu, v0, sh);
      v1 += __shfl_xor_sync(0xffffffffu, v1, sh);
    }
    if ((lane >> 2) == 0) {
      float2* d = reinterpret_cast<float2*>(s_dst + nt * 8 + 2 * t);
      float2 c = *d;
      c.x += v0; c.y += v1;
      *d = c;
    }
  }
}

template <bool PACKED>
__global__ void __launch_bounds__(kBwdWarps * 32, 1)
attn_fast_bwd_kernel(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_do,
                     const __grid_constant__ CUtensorMap map_dqkv, const FastParams p) {
  pdl_prologue_done();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* tab = smem + kBwdWarps * BwdWarp::kBytes;
  fill_tables(p, tab, threadIdx.x, blockDim.x);
  // CTA-level gradient accumulators after the tables
  float* s_acc = reinterpret_cast<float*>(tab + tab_bytes(p.heads));
  float* s_dqw = s_acc;                 // [64]
  float* s_dqb = s_dqw + FD;
  float* s_dkw = s_dqb + FD;
  float* s_demb = s_dkw + FD;           // [32 * heads]
  float* s_dsf = s_demb + 32 * p.heads; // [heads]
  // [warps][3 * 64] column sums of dq | dk | dv of each warp's head (input_head bias gradient), 8-byte aligned
  float* s_dbias = s_dsf + ((p.heads + 1) & ~1);
  const int n_acc = 3 * FD + 33 * p.heads + 2 + (p.d_qkv_bias != nullptr ? 3 * FD * kBwdWarps : 0);
  for (int i = threadIdx.x; i < n_acc; i += blockDim.x) s_acc[i] = 0.f;
  uint8_t* my = smem + warp * BwdWarp::kBytes;
  uint8_t* sQ = my + BwdWarp::kQ;
  uint8_t* sK = my + BwdWarp::kK;
  uint8_t* sV = my + BwdWarp::kV;       // after dP: P' in columns 0..31, dS in columns 32..63
  uint8_t* sdo = my + BwdWarp::kDo;
  uint64_t* bar = reinterpret_cast<uint64_t*>(my + BwdWarp::kBar);
  int* rowgp = reinterpret_cast<int*>(my + BwdWarp::kRowGp);
  const float* brel_all = reinterpret_cast<const float*>(tab + kTabBrel);
  float* rstd_s = reinterpret_cast<float*>(my + BwdWarp::kRstd);
  zero_tiles(my, 4 * kTile, lane);
  if (lane == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
    tma_prefetch_desc(&map_qkv);
    tma_prefetch_desc(&map_do);
    tma_prefetch_desc(&map_dqkv);
  }
  fence_proxy_async();
  __syncthreads();
  const uint32_t* pairs = reinterpret_cast<const uint32_t*>(tab + kTabPairs);
  const uint32_t* splat = reinterpret_cast<const uint32_t*>(tab + kTabSplat);
  const float* wq = reinterpret_cast<const float*>(tab + kTabW);
  const float* wk = wq + FD;
  uint32_t phase = 0;
  const int L = p.L, G = p.G;
  const float invL = 1.f / (float)L;
  const long n_work = p.n_tiles * p.heads;
  const int g8 = lane >> 2, t = lane & 3;
  const float qscale = rsqrtf((float)FD);
  float dwq[8][2], dbq[8][2], dwk[8][2], dbk_unused[8][2];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) { dwq[nt][0] = dwq[nt][1] = dbq[nt][0] = dbq[nt][1] = dwk[nt][0] = dwk[nt][1] = 0.f; }
  // !PACKED (one 32-token sequence per tile): the launch guarantees gridDim*warps % heads == 0, so a warp always
  // works on the same head and can keep the relative-position-bias and scale-factor gradients in registers:
  // a lane's dS elements fall on 14 diagonals  rel = 8*q + (2t + e - g),  q = nt - 2*mt - half in [-3, 3].
  float dacc[14];
#pragma unroll
  for (int k = 0; k < 14; ++k) dacc[k] = 0.f;
  float dsf_acc = 0.f;
  const int my_head = (int)(((long)blockIdx.x * kBwdWarps + warp) % p.heads);
  const int pk_g = lane / L;            // packed tiles: this lane's row belongs to sequence pk_g of the tile

  for (long wi = (long)blockIdx.x * kBwdWarps + warp; wi < n_work; wi += (long)gridDim.x * kBwdWarps) {
    const Item it = load_item<true>(p, &map_qkv, &map_do, wi, my, bar, phase, rowgp, rstd_s, lane);
    const int head = it.head;
    const float* brel = brel_all + head * 64;
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
      frag_a(a0, sdo, 0, ks * 16, lane);
      frag_a(a1, sdo, 16, ks * 16, lane);
#pragma unroll
      for (int np = 0; np < 2; ++np) {
        uint32_t b[4];
        frag_b(b, sV, np * 16, ks * 16, lane);
        mma16816(dp[0][2 * np], a0, b[0], b[1]);
        mma16816(dp[0][2 * np + 1], a0, b[2], b[3]);
        mma16816(dp[1][2 * np], a1, b[0], b[1]);
        mma16816(dp[1][2 * np + 1], a1, b[2], b[3]);
      }
    }
    __syncwarp();                       // every lane is done reading v: its tile now receives P' and dS
    float dsf = 0.f;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      float acc[4][4];
      scores16(acc, sQ, sK, pairs, mt * 16, lane);
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
          *reinterpret_cast<uint32_t*>(sV + swz(i, nt * 8 + 2 * t)) = pack_bf2(pp[0], pp[1]);
          *reinterpret_cast<uint32_t*>(sV + swz(i, 32 + nt * 8 + 2 * t)) = pack_bf2(ds[0], ds[1]);
          if (!PACKED) {
            dacc[(nt - 2 * mt - half + 3) * 2] += ds[0];
            dacc[(nt - 2 * mt - half + 3) * 2 + 1] += ds[1];
          }
        }
      }
    }
    __syncwarp();
    // ---- bias-embedding and scale-factor gradients ----
    if (PACKED) {
      if (p.d_bias_emb != nullptr) {
        // Lane p = (sequence gq, query i) reads the L entries dS[p][gq*L + j] of its own row; the G sequences are added
        // up with strided shuffles (lanes i, i + L, ... hold the same i), and entry (i, j) then moves to the lane that
        // owns its relative position r = j - i + L - 1 (lane r, second accumulator for r >= 32) for the whole launch:
        // ~2 G shuffles per j instead of a G x L loop per relative position on 2L - 1 of the 32 lanes (that loop was
        // 17 % of the packed kernel's instructions).
        for (int j = 0; j < L; ++j) {
          const float v = pk_g < G ? __bfloat162float(*reinterpret_cast<const bf16*>(sV + swz(lane, 32 + pk_g * L + j))) : 0.f;
          float tot = v;
          for (int sq = 1; sq < G; ++sq) tot += __shfl_sync(0xffffffffu, v, (lane + sq * L) & 31);
          const int src0 = j - lane + (L - 1), src1 = src0 - 32;
          const float w0 = __shfl_sync(0xffffffffu, tot, src0 & 31);
          const float w1 = __shfl_sync(0xffffffffu, tot, src1 & 31);
          if (src0 >= 0 && src0 < L) dacc[0] += w0;
          if (src1 >= 0 && src1 < L) dacc[1] += w1;
        }
      }
    }
    dsf_acc += dsf;
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
        frag_a_t(a0, sV, 0, kk * 16, lane);
        frag_a_t(a1, sV, 16, kk * 16, lane);
#pragma unroll
        for (int np = 0; np < 4; ++np) {
          uint32_t b[4];
          frag_b_t(b, sdo, np * 16, kk * 16, lane);
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
        stage16(sdo, mt * 16, o[mt], lane);
      }
      if (p.d_qkv_bias != nullptr) colsum_frag2(s_dbias + warp * 3 * FD + 2 * FD, o, lane);
      store_tile(&map_dqkv, sdo, head * 3 * FD + 2 * FD, it, p.accumulate, lane);
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
        frag_a_t(a0, sV, 32, kk * 16, lane);
        frag_a_t(a1, sV, 32 + 16, kk * 16, lane);
#pragma unroll
        for (int np = 0; np < 4; ++np) {
          uint32_t b[4];
          frag_b_t(b, sQ, np * 16, kk * 16, lane);          // Q as [k = i][n = d]: both halves of a register share d
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
      for (int mt = 0; mt < 2; ++mt) ln_bwd16<false, false>(o[mt], 1.f, sK, rstd_s, 1, mt * 16, wk, dwk, dbk_unused, lane);
      if (lane == 0) tma_store_wait_read<0>();     // the dV store has finished reading the staging tile
      __syncwarp();
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        stage16(sdo, mt * 16, o[mt], lane);
      }
      if (p.d_qkv_bias != nullptr) colsum_frag2(s_dbias + warp * 3 * FD + FD, o, lane);
      store_tile(&map_dqkv, sdo, head * 3 * FD + FD, it, p.accumulate, lane);
    }
    // ---- dQ' = dS K'  -> LayerNorm backward -> d(raw q), staged over the (dead) q rows it was derived from ----
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      float o[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) { o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f; }
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        uint32_t a[4];
        frag_a(a, sV, mt * 16, 32 + kk * 16, lane);
#pragma unroll
        for (int np = 0; np < 4; ++np) {
          uint32_t b[4];
          frag_b_t(b, sK, np * 16, kk * 16, lane);
          const uint32_t al = splat[128 + np * 16 + g8], ah = splat[128 + np * 16 + 8 + g8];
          const uint32_t bl = splat[192 + np * 16 + g8], bh = splat[192 + np * 16 + 8 + g8];
          b[0] = hfma2_bf16(b[0], al, bl); b[1] = hfma2_bf16(b[1], al, bl);
          b[2] = hfma2_bf16(b[2], ah, bh); b[3] = hfma2_bf16(b[3], ah, bh);
          mma16816(o[2 * np], a, b[0], b[1]);
          mma16816(o[2 * np + 1], a, b[2], b[3]);
        }
      }
      ln_bwd16<true, true>(o, qscale, sQ, rstd_s, 0, mt * 16, wq, dwq, dbq, lane);
      __syncwarp();                     // all lanes have read the xhat rows of this m tile
      stage16(sQ, mt * 16, o, lane);
      if (p.d_qkv_bias != nullptr) colsum_frag1(s_dbias + warp * 3 * FD, o, lane);
    }
    store_tile(&map_dqkv, sQ, head * 3 * FD, it, p.accumulate, lane);
  }
  if (lane == 0) tma_store_wait_all();
  __syncwarp();
  // ---- parameter gradients.  The warp's tiles are dead: their first 1 KiB + becomes the warp's PRIVATE table
  //      [0:64) d qnorm.weight, [64:128) d qnorm.bias, [128:192) d knorm.weight, [192:255) bias gradient per relative
  //      position (rel + 31), [256] d scale factor.  No shared-memory atomics (fp32 shared atomics are compare-and-swap
  //      loops: twelve warps contending on the block-level tables were 15 % of this kernel's samples); the block's
  //      tables are added up after one __syncthreads and leave as one global atomic per entry and block. ----
  float* wtab = reinterpret_cast<float*>(my);
  wtab[192 + lane] = 0.f;
  wtab[224 + lane] = 0.f;
  if (lane == 0) wtab[256] = 0.f;
  __syncwarp();
  if (p.d_bias_emb != nullptr) {
    if (PACKED) {
      if (lane < 2 * L - 1) wtab[192 + lane] = dacc[0];
      if (lane + 32 < 2 * L - 1) wtab[224 + lane] = dacc[1];
    } else {
      // rel = 8*q + 2t + e - g: for a fixed k the lanes that share a rel differ in t, so four rounds (one per t) never
      // have two lanes on the same entry
#pragma unroll
      for (int k = 0; k < 14; ++k) {
        const int rel = 8 * ((k >> 1) - 3) + 2 * t + (k & 1) - g8;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          if (t == r) wtab[192 + rel + FLP - 1] += dacc[k];
          __syncwarp();
        }
      }
    }
  }
  if (p.d_scale_factor != nullptr) {
    dsf_acc = warp_sum(dsf_acc);
    if (lane == 0) wtab[256] = dsf_acc;
  }
  // LayerNorm parameter gradients: lanes with equal t hold partial sums of the same columns
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
        wtab[nt * 8 + 2 * t + e] = a;
        wtab[FD + nt * 8 + 2 * t + e] = b;
        wtab[2 * FD + nt * 8 + 2 * t + e] = c;
      }
    }
  }
  __syncthreads();
  const float* wt0 = reinterpret_cast<const float*>(smem);
  constexpr int kWStride = BwdWarp::kBytes / 4;          // floats between two warps' tables
  for (int i = threadIdx.x; i < 3 * FD; i += blockDim.x) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < kBwdWarps; ++w) v += wt0[w * kWStride + i];
    float* dst = i < FD ? p.d_qn_w + i : (i < 2 * FD ? p.d_qn_b + (i - FD) : p.d_kn_w + (i - 2 * FD));
    atomicAdd(dst, v);
  }
  // the launch guarantees gridDim * warps % heads == 0: warp w of this block worked on head (block * warps + w) % heads
  if (p.d_bias_emb != nullptr) {
    const int nrel = 2 * L - 1;                 // table index = rel + L - 1
    for (int i = threadIdx.x; i < nrel * p.heads; i += blockDim.x) {
      const int h = i / nrel, r = i - h * nrel;
      float v = 0.f;
      for (int w = 0; w < kBwdWarps; ++w)
        if ((int)(((long)blockIdx.x * kBwdWarps + w) % p.heads) == h) v += wt0[w * kWStride + 192 + r];
      if (v != 0.f) atomicAdd(p.d_bias_emb + __ldg(p.bucket + r) * p.heads + h, v);
    }
  }
  if (p.d_scale_factor != nullptr) {
    for (int h = threadIdx.x; h < p.heads; h += blockDim.x) {
      float v = 0.f;
      for (int w = 0; w < kBwdWarps; ++w)
        if ((int)(((long)blockIdx.x * kBwdWarps + w) % p.heads) == h) v += wt0[w * kWStride + 256];
      atomicAdd(p.d_scale_factor + h, v);
    }
  }
  if (p.d_qkv_bias != nullptr) {
    for (int i = threadIdx.x; i < 3 * FD * p.heads; i += blockDim.x) {
      const int h = i / (3 * FD), col = i - h * 3 * FD;
      float v = 0.f;
      for (int w = 0; w < kBwdWarps; ++w)
        if ((int)(((long)blockIdx.x * kBwdWarps + w) % p.heads) == h) v += s_dbias[w * 3 * FD + col];
      if (v != 0.f) atomicAdd(p.d_qkv_bias + i, v);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// backward, two warps per work item
// ---------------------------------------------------------------------------------------------
// The one-warp backward above is bound by its own length: ~8 400 instructions per (tile, head) with 168 registers per
// thread, so only 12 warps fit an SM (register file AND 17 KB of tiles per warp) and each of them is one long dependent
// chain.  Here a PAIR of warps shares the 17 KB of one item: 24 warps per SM on the same shared memory, half the chain
// per warp.
//   phase 1 (split by query m-tile: warp h owns rows 16h..16h+15): dP = dO V^T, S = Q'K'^T, softmax, dS; P' and dS rows go
//            into the (then dead) V tile
//   phase 2 (split by output COLUMNS: warp h owns columns 32h..32h+31 of dV, dK, dQ): each warp only ever reads and
//            overwrites its own column half of the dO / Q tiles, so staging needs no extra buffers; the LayerNorm
//            backward's row sums over all 64 columns are exchanged through 512 B of shared memory.
// Eight named barriers of 64 threads per item order the hand-offs.  Warp 0 lane 0 issues every TMA load and store.
constexpr int kBwd2Pairs = 12;

struct Bwd2Pair {
  static constexpr int kPart = BwdWarp::kRstd + FLP * 8;          // float[2][32][2]: per-warp partial LayerNorm row sums
  static constexpr int kBytes = ((kPart + 2 * FLP * 8) + 1023) / 1024 * 1024;
};

__device__ __forceinline__ void pair_sync(int pair) {
  asm volatile("bar.sync %0, 64;" ::"r"(pair + 1) : "memory");
}

// 16 x 32 C fragments (4 n tiles) -> bf16 columns col0.. of rows m0.. of a swizzled tile
__device__ __forceinline__ void stage16h(uint8_t* tile, int m0, int col0, const float (&o)[4][4], int lane) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    *reinterpret_cast<uint32_t*>(tile + swz(m0 + g, col0 + nt * 8 + 2 * t)) = pack_bf2(o[nt][0], o[nt][1]);
    *reinterpret_cast<uint32_t*>(tile + swz(m0 + g + 8, col0 + nt * 8 + 2 * t)) = pack_bf2(o[nt][2], o[nt][3]);
  }
}

// LayerNorm backward on this warp's 32 columns of a 16-row tile, first half: acc = dL/dy -> dn = dy * w (in place),
// parameter-gradient partials, and this warp's share of the row sums (sum dn, sum dn * xhat) -> part[row][2]
template <bool DB, bool SCALE>
__device__ __forceinline__ void ln_bwd_h1(float (&acc)[4][4], float pre, const uint8_t* xh, int m0, int col0, const float* w,
                                          float (&dw)[4][2], float (&db)[4][2], float* part, int lane) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int r = m0 + g + half * 8;
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int c = col0 + nt * 8 + 2 * t;
      const float2 xv = unpack2<bf16>(*reinterpret_cast<const uint32_t*>(xh + swz(r, c)));
      const float2 wv = *reinterpret_cast<const float2*>(w + c);
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float n = e ? xv.y : xv.x;
        const float dy = SCALE ? acc[nt][half * 2 + e] * pre : acc[nt][half * 2 + e];
        dw[nt][e] = fmaf(dy, n, dw[nt][e]);
        if (DB) db[nt][e] += dy;
        const float dn = dy * (e ? wv.y : wv.x);
        acc[nt][half * 2 + e] = dn;
        s1 += dn;
        s2 = fmaf(dn, n, s2);
      }
    }
    s1 = qsum(s1); s2 = qsum(s2);
    if (t == 0) { part[2 * r] = s1; part[2 * r + 1] = s2; }
  }
}
// second half, after the partner's partial sums are visible: d raw = rstd * (dn - mean(dn) - xhat * mean(dn * xhat))
__device__ __forceinline__ void ln_bwd_h2(float (&acc)[4][4], const uint8_t* xh, const float* rstd_s, int which, int m0,
                                          int col0, const float* part_a, const float* part_b, int lane) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int r = m0 + g + half * 8;
    const float rstd = rstd_s[2 * r + which];
    const float c0 = -rstd * (part_a[2 * r] + part_b[2 * r]) * (1.f / FD);
    const float c1 = -rstd * (part_a[2 * r + 1] + part_b[2 * r + 1]) * (1.f / FD);
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const float2 xv = unpack2<bf16>(*reinterpret_cast<const uint32_t*>(xh + swz(r, col0 + nt * 8 + 2 * t)));
      acc[nt][half * 2] = fmaf(xv.x, c1, fmaf(acc[nt][half * 2], rstd, c0));
      acc[nt][half * 2 + 1] = fmaf(xv.y, c1, fmaf(acc[nt][half * 2 + 1], rstd, c0));
    }
  }
}

template <bool PACKED>
__global__ void __launch_bounds__(kBwd2Pairs * 64, 1)
attn_fast_bwd2_kernel(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_do,
                      const __grid_constant__ CUtensorMap map_dqkv, const FastParams p) {
  pdl_prologue_done();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = warp >> 1, h = warp & 1;
  uint8_t* tab = smem + kBwd2Pairs * Bwd2Pair::kBytes;
  fill_tables(p, tab, threadIdx.x, blockDim.x);
  float* s_acc = reinterpret_cast<float*>(tab + tab_bytes(p.heads));
  float* s_dqw = s_acc;                 // [64]
  float* s_dqb = s_dqw + FD;
  float* s_dkw = s_dqb + FD;
  float* s_demb = s_dkw + FD;           // [32 * heads]
  float* s_dsf = s_demb + 32 * p.heads; // [heads]
  const int n_acc = 3 * FD + 33 * p.heads;
  for (int i = threadIdx.x; i < n_acc; i += blockDim.x) s_acc[i] = 0.f;
  uint8_t* my = smem + pair * Bwd2Pair::kBytes;
  uint8_t* sQ = my + BwdWarp::kQ;
  uint8_t* sK = my + BwdWarp::kK;
  uint8_t* sV = my + BwdWarp::kV;       // after dP: P' in columns 0..31, dS in columns 32..63
  uint8_t* sdo = my + BwdWarp::kDo;
  uint64_t* bar = reinterpret_cast<uint64_t*>(my + BwdWarp::kBar);
  int* rowgp = reinterpret_cast<int*>(my + BwdWarp::kRowGp);
  float* rstd_s = reinterpret_cast<float*>(my + BwdWarp::kRstd);
  float* part_me = reinterpret_cast<float*>(my + Bwd2Pair::kPart) + h * 2 * FLP;
  float* part_ot = reinterpret_cast<float*>(my + Bwd2Pair::kPart) + (1 - h) * 2 * FLP;
  const float* brel_all = reinterpret_cast<const float*>(tab + kTabBrel);
  if (h == 0) zero_tiles(my, 4 * kTile, lane);
  if (h == 0 && lane == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
    tma_prefetch_desc(&map_qkv);
    tma_prefetch_desc(&map_do);
    tma_prefetch_desc(&map_dqkv);
  }
  fence_proxy_async();
  __syncthreads();
  const uint32_t* pairs = reinterpret_cast<const uint32_t*>(tab + kTabPairs);
  const uint32_t* splat = reinterpret_cast<const uint32_t*>(tab + kTabSplat);
  const float* wq = reinterpret_cast<const float*>(tab + kTabW);
  const float* wk = wq + FD;
  uint32_t phase = 0;
  const int L = p.L, G = p.G;
  const float invL = 1.f / (float)L;
  const long n_work = p.n_tiles * p.heads;
  const int g8 = lane >> 2, t = lane & 3;
  const float qscale = rsqrtf((float)FD);
  const int mt = h;                     // phase 1: this warp's query m-tile
  const int col0 = 32 * h;              // phase 2: this warp's output columns
  float dwq[4][2], dbq[4][2], dwk[4][2], dbk_unused[4][2];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) { dwq[nt][0] = dwq[nt][1] = dbq[nt][0] = dbq[nt][1] = dwk[nt][0] = dwk[nt][1] = 0.f; }
  // !PACKED: the launch guarantees gridDim * pairs % heads == 0, so a pair keeps one head; a lane's dS elements of m-tile
  // mt fall on the diagonals  rel = 8*q + (2t + e - g),  q = nt - 2*mt - half in [-3, 3]
  float dacc[14];
#pragma unroll
  for (int k = 0; k < 14; ++k) dacc[k] = 0.f;
  float dsf_acc = 0.f;
  const int my_head = (int)(((long)blockIdx.x * kBwd2Pairs + pair) % p.heads);

  for (long wi = (long)blockIdx.x * kBwd2Pairs + pair; wi < n_work; wi += (long)gridDim.x * kBwd2Pairs) {
    // ---- item geometry (both warps), loads + row tables (warp 0) ----
    Item it;
    {
      const long tile = wi / p.heads;
      it.head = (int)(wi - tile * p.heads);
      it.s_out = (int)(tile / p.tiles_per_outer);
      it.s_in0 = (int)(tile - (long)it.s_out * p.tiles_per_outer) * G;
    }
    if (h == 0) {
      if (lane == 0) {
        tma_store_wait_read<0>();          // the previous item's three stores have finished reading the tiles
        mbar_arrive_expect_tx(bar, (uint32_t)(4 * G * L * FD * 2));
        const int col = it.head * 3 * FD;
        tma_load_4d(my, &map_qkv, bar, col, 0, it.s_in0, it.s_out);
        tma_load_4d(my + kTile, &map_qkv, bar, col + FD, 0, it.s_in0, it.s_out);
        tma_load_4d(my + 2 * kTile, &map_qkv, bar, col + 2 * FD, 0, it.s_in0, it.s_out);
        tma_load_4d(my + 3 * kTile, &map_do, bar, it.head * FD, 0, it.s_in0, it.s_out);
      }
      const int g = lane / L, i = lane - g * L;
      const bool ok = g < G && it.s_in0 + g < p.inner;
      float2 rr = make_float2(0.f, 0.f);
      if (ok) {
        const long tok = (long)it.s_out * p.outer_stride + (long)(it.s_in0 + g) * p.inner_stride + (long)i * p.tok_stride;
        rr = __ldg(reinterpret_cast<const float2*>(p.rstd + (tok * p.heads + it.head) * 2));
      }
      rowgp[lane] = ok ? ((g << 8) | i) : (255 << 8);
      rstd_s[2 * lane] = rr.x; rstd_s[2 * lane + 1] = rr.y;
    }
    mbar_wait(bar, phase);
    phase ^= 1u;
    pair_sync(pair);                                         // B1: row tables visible to warp 1
    const int head = it.head;
    const float* brel = brel_all + head * 64;
    const float sf = p.scale_factor != nullptr ? __ldg(p.scale_factor + head) : 1.f;
    const float lowc = (1.f - sf) * invL;

    // ---- phase 1: rows 16*mt .. 16*mt+15 ----
    float dp[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) { dp[nt][0] = dp[nt][1] = dp[nt][2] = dp[nt][3] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < FD / 16; ++ks) {
      uint32_t a0[4];
      frag_a(a0, sdo, mt * 16, ks * 16, lane);
#pragma unroll
      for (int np = 0; np < 2; ++np) {
        uint32_t b[4];
        frag_b(b, sV, np * 16, ks * 16, lane);
        mma16816(dp[2 * np], a0, b[0], b[1]);
        mma16816(dp[2 * np + 1], a0, b[2], b[3]);
      }
    }
    float acc[4][4];
    scores16(acc, sQ, sK, pairs, mt * 16, lane);
    softmax16<PACKED>(acc, brel, rowgp, L, mt * 16, lane);
    pair_sync(pair);                                         // B2: both warps are done reading v
    float dsf = 0.f;
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
          const float d = dp[nt][half * 2 + e] * p.out_scale;
          acc[nt][half * 2 + e] = pv;
          dp[nt][half * 2 + e] = d;
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
          ds[e] = sf * pv * (dp[nt][half * 2 + e] - dot);
          bool valid = true;
          if (PACKED) valid = gi != 255 && (rowgp[nt * 8 + 2 * t + e] >> 8) == gi;
          pp[e] = valid ? fmaf(pv, sf, lowc) : 0.f;
        }
        *reinterpret_cast<uint32_t*>(sV + swz(i, nt * 8 + 2 * t)) = pack_bf2(pp[0], pp[1]);
        *reinterpret_cast<uint32_t*>(sV + swz(i, 32 + nt * 8 + 2 * t)) = pack_bf2(ds[0], ds[1]);
        if (!PACKED) {
          dacc[(nt - 2 * mt - half + 3) * 2] += ds[0];
          dacc[(nt - 2 * mt - half + 3) * 2 + 1] += ds[1];
        }
      }
    }
    pair_sync(pair);                                         // B3: all of P' and dS is in the v tile
    if (PACKED) {
      if (p.d_bias_emb != nullptr) {
        const int r = lane + 32 * h;
        if (r < 2 * L - 1) {
          float s = 0.f;
          for (int gq = 0; gq < G; ++gq) {
            for (int i = 0; i < L; ++i) {
              const int j = i + r - (L - 1);
              if (j >= 0 && j < L) s += __bfloat162float(*reinterpret_cast<const bf16*>(sV + swz(gq * L + i, 32 + gq * L + j)));
            }
          }
          atomicAdd(s_demb + __ldg(p.bucket + r) * p.heads + head, s);
        }
      }
      if (p.d_scale_factor != nullptr) {
        dsf = warp_sum(dsf);
        if (lane == 0) atomicAdd(s_dsf + head, dsf);
      }
    } else {
      dsf_acc += dsf;
    }

    // ---- phase 2: output columns col0 .. col0+31 ----
    float o[2][4][4];
    // dV = P'^T dO'
#pragma unroll
    for (int m = 0; m < 2; ++m) {
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) { o[m][nt][0] = o[m][nt][1] = o[m][nt][2] = o[m][nt][3] = 0.f; }
    }
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
      uint32_t a0[4], a1[4];
      frag_a_t(a0, sV, 0, kk * 16, lane);
      frag_a_t(a1, sV, 16, kk * 16, lane);
#pragma unroll
      for (int np = 0; np < 2; ++np) {
        uint32_t b[4];
        frag_b_t(b, sdo, col0 + np * 16, kk * 16, lane);
        mma16816(o[0][2 * np], a0, b[0], b[1]);
        mma16816(o[0][2 * np + 1], a0, b[2], b[3]);
        mma16816(o[1][2 * np], a1, b[0], b[1]);
        mma16816(o[1][2 * np + 1], a1, b[2], b[3]);
      }
    }
    __syncwarp();                       // this warp's columns of dO are dead: they become its staging columns
#pragma unroll
    for (int m = 0; m < 2; ++m) {
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) o[m][nt][e] *= p.out_scale;
      }
      stage16h(sdo, m * 16, col0, o[m], lane);
    }
    fence_proxy_async();
    pair_sync(pair);                                         // B4: dV staged by both warps
    if (h == 0 && lane == 0) {
      if (p.accumulate) tma_reduce_add_4d(&map_dqkv, sdo, head * 3 * FD + 2 * FD, 0, it.s_in0, it.s_out);
      else tma_store_4d(&map_dqkv, sdo, head * 3 * FD + 2 * FD, 0, it.s_in0, it.s_out);
      tma_store_commit();
    }
    // dK' = dS^T Q'  -> LayerNorm backward -> d(raw k)
#pragma unroll
    for (int m = 0; m < 2; ++m) {
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) { o[m][nt][0] = o[m][nt][1] = o[m][nt][2] = o[m][nt][3] = 0.f; }
    }
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
      uint32_t a0[4], a1[4];
      frag_a_t(a0, sV, 32, kk * 16, lane);
      frag_a_t(a1, sV, 32 + 16, kk * 16, lane);
#pragma unroll
      for (int np = 0; np < 2; ++np) {
        uint32_t b[4];
        const int n0 = col0 + np * 16;
        frag_b_t(b, sQ, n0, kk * 16, lane);                  // Q as [k = i][n = d]: both halves of a register share d
        const uint32_t al = splat[n0 + g8], ah = splat[n0 + 8 + g8];
        const uint32_t bl = splat[64 + n0 + g8], bh = splat[64 + n0 + 8 + g8];
        b[0] = hfma2_bf16(b[0], al, bl); b[1] = hfma2_bf16(b[1], al, bl);
        b[2] = hfma2_bf16(b[2], ah, bh); b[3] = hfma2_bf16(b[3], ah, bh);
        mma16816(o[0][2 * np], a0, b[0], b[1]);
        mma16816(o[0][2 * np + 1], a0, b[2], b[3]);
        mma16816(o[1][2 * np], a1, b[0], b[1]);
        mma16816(o[1][2 * np + 1], a1, b[2], b[3]);
      }
    }
#pragma unroll
    for (int m = 0; m < 2; ++m) ln_bwd_h1<false, false>(o[m], 1.f, sK, m * 16, col0, wk, dwk, dbk_unused, part_me, lane);
    if (h == 0 && lane == 0) tma_store_wait_read<0>();       // the dV store has finished reading the staging tile
    pair_sync(pair);                                         // B5: partial row sums exchanged, staging tile free
#pragma unroll
    for (int m = 0; m < 2; ++m) {
      ln_bwd_h2(o[m], sK, rstd_s, 1, m * 16, col0, part_me, part_ot, lane);
      stage16h(sdo, m * 16, col0, o[m], lane);
    }
    fence_proxy_async();
    pair_sync(pair);                                         // B6: dK staged (and the partial sums consumed)
    if (h == 0 && lane == 0) {
      if (p.accumulate) tma_reduce_add_4d(&map_dqkv, sdo, head * 3 * FD + FD, 0, it.s_in0, it.s_out);
      else tma_store_4d(&map_dqkv, sdo, head * 3 * FD + FD, 0, it.s_in0, it.s_out);
      tma_store_commit();
    }
    // dQ' = dS K'  -> LayerNorm backward -> d(raw q), staged over the q columns this warp owns
#pragma unroll
    for (int m = 0; m < 2; ++m) {
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) { o[m][nt][0] = o[m][nt][1] = o[m][nt][2] = o[m][nt][3] = 0.f; }
    }
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
      uint32_t a0[4], a1[4];
      frag_a(a0, sV, 0, 32 + kk * 16, lane);
      frag_a(a1, sV, 16, 32 + kk * 16, lane);
#pragma unroll
      for (int np = 0; np < 2; ++np) {
        uint32_t b[4];
        const int n0 = col0 + np * 16;
        frag_b_t(b, sK, n0, kk * 16, lane);
        const uint32_t al = splat[128 + n0 + g8], ah = splat[128 + n0 + 8 + g8];
        const uint32_t bl = splat[192 + n0 + g8], bh = splat[192 + n0 + 8 + g8];
        b[0] = hfma2_bf16(b[0], al, bl); b[1] = hfma2_bf16(b[1], al, bl);
        b[2] = hfma2_bf16(b[2], ah, bh); b[3] = hfma2_bf16(b[3], ah, bh);
        mma16816(o[0][2 * np], a0, b[0], b[1]);
        mma16816(o[0][2 * np + 1], a0, b[2], b[3]);
        mma16816(o[1][2 * np], a1, b[0], b[1]);
        mma16816(o[1][2 * np + 1], a1, b[2], b[3]);
      }
    }
#pragma unroll
    for (int m = 0; m < 2; ++m) ln_bwd_h1<true, true>(o[m], qscale, sQ, m * 16, col0, wq, dwq, dbq, part_me, lane);
    pair_sync(pair);                                         // B7: partial row sums exchanged
#pragma unroll
    for (int m = 0; m < 2; ++m) {
      ln_bwd_h2(o[m], sQ, rstd_s, 0, m * 16, col0, part_me, part_ot, lane);
      __syncwarp();                     // all lanes have read the xhat values of these rows / columns
      stage16h(sQ, m * 16, col0, o[m], lane);
    }
    fence_proxy_async();
    pair_sync(pair);                                         // B8: dQ staged; row tables / partials free for the next item
    if (h == 0 && lane == 0) {
      if (p.accumulate) tma_reduce_add_4d(&map_dqkv, sQ, head * 3 * FD, 0, it.s_in0, it.s_out);
      else tma_store_4d(&map_dqkv, sQ, head * 3 * FD, 0, it.s_in0, it.s_out);
      tma_store_commit();
    }
  }
  if (h == 0 && lane == 0) tma_store_wait_all();
  if (!PACKED) {
    if (p.d_bias_emb != nullptr) {
#pragma unroll
      for (int k = 0; k < 14; ++k) {
        const int rel = 8 * ((k >> 1) - 3) + 2 * t + (k & 1) - g8;
        if (rel > -FLP && rel < FLP) atomicAdd(s_demb + __ldg(p.bucket + rel + FLP - 1) * p.heads + my_head, dacc[k]);
      }
    }
    if (p.d_scale_factor != nullptr) {
      dsf_acc = warp_sum(dsf_acc);
      if (lane == 0) atomicAdd(s_dsf + my_head, dsf_acc);
    }
  }
  // LayerNorm parameter gradients: lanes with equal t hold partial sums of the same columns
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      float a = dwq[nt][e], b = dbq[nt][e], c = dwk[nt][e];
#pragma unroll
      for (int o2 = 4; o2 < 32; o2 <<= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o2);
        b += __shfl_xor_sync(0xffffffffu, b, o2);
        c += __shfl_xor_sync(0xffffffffu, c, o2);
      }
      if (g8 == 0) {
        atomicAdd(s_dqw + col0 + nt * 8 + 2 * t + e, a);
        atomicAdd(s_dqb + col0 + nt * 8 + 2 * t + e, b);
        atomicAdd(s_dkw + col0 + nt * 8 + 2 * t + e, c);
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
}

// ---------------------------------------------------------------------------------------------
// host launch
// ---------------------------------------------------------------------------------------------
// 4-D view (column, position, sequence, outer) of a token-major (tokens, width) bf16 matrix; box = 64 x L x G x 1
static int make_seq_map(CUtensorMap* map, const void* base, long ld, int width, const bf_attn_args* a, int G) {
  const long n_outer = a->n_seq / a->inner;
  uint64_t dims[4] = {(uint64_t)width, (uint64_t)a->L, (uint64_t)a->inner, (uint64_t)n_outer};
  uint64_t str[3] = {(uint64_t)a->tok_stride * ld * 2, (uint64_t)a->inner_stride * ld * 2, (uint64_t)a->outer_stride * ld * 2};
  uint32_t box[4] = {FD, (uint32_t)a->L, (uint32_t)G, 1};
  return make_map(map, BF_BF16, base, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

int launch_attn_fast(const bf_attn_args* a, bool bwd, cudaStream_t st) {
  BF_REQUIRE(a->head_dim == FD && a->L >= 1 && a->L <= FLP,
             "bf_attention (prenorm): the pre-normalised path handles head_dim 64 and L <= 32 (got d=%d L=%d)",
             a->head_dim, a->L);
  BF_REQUIRE(!bwd || a->rstd != nullptr, "bf_attention_bwd (prenorm): rstd required");
  BF_REQUIRE(a->n_seq % a->inner == 0, "bf_attention (prenorm): n_seq=%ld must be a multiple of inner=%ld",
             (long)a->n_seq, (long)a->inner);
  BF_REQUIRE(a->inner < (1l << 31) && a->n_seq / a->inner < (1l << 31), "bf_attention (prenorm): geometry too large");
  FastParams p{};
  p.rstd = a->rstd;
  p.heads = a->heads; p.L = a->L;
  p.G = FLP / a->L;
  p.inner = (int)a->inner;
  p.tiles_per_outer = (int)((a->inner + p.G - 1) / p.G);
  p.n_tiles = (a->n_seq / a->inner) * p.tiles_per_outer;
  p.outer_stride = a->outer_stride; p.inner_stride = a->inner_stride; p.tok_stride = a->tok_stride;
  p.qn_w = a->qn_w; p.qn_b = a->qn_b; p.kn_w = a->kn_w; p.kn_b = a->kn_b;
  p.bias_emb = a->bias_emb; p.bucket = a->bucket; p.scale_factor = a->scale_factor;
  p.out_scale = a->out_scale; p.accumulate = a->accumulate;
  p.d_qn_w = a->d_qn_w; p.d_qn_b = a->d_qn_b; p.d_kn_w = a->d_kn_w; p.d_kn_b = a->d_kn_b;
  p.d_bias_emb = a->d_bias_emb; p.d_scale_factor = a->d_scale_factor;
  p.d_qkv_bias = bwd ? a->d_qkv_bias : nullptr;
  const bool packed = !(p.G == 1 && a->L == FLP);
  const int E3 = 3 * FD * a->heads;
  CUtensorMap m_qkv, m_b, m_c;
  if (int e = make_seq_map(&m_qkv, a->qkv, a->ld_qkv, E3, a, p.G)) return e;
  if (bwd) {
    if (int e = make_seq_map(&m_b, a->dout, a->ld_dout, E3 / 3, a, p.G)) return e;
    if (int e = make_seq_map(&m_c, a->out, a->ld_out, E3, a, p.G)) return e;
  } else {
    if (int e = make_seq_map(&m_b, a->out, a->ld_out, E3 / 3, a, p.G)) return e;
  }
  // backward: BF_ATTN_BWD2=1 selects the two-warps-per-item kernel (attn_fast_bwd2_kernel).  Measured on B200 (round 2,
  // same-box A/B of the replayed step): 29.36 vs 29.24 ms per step, 105 vs 104 us per L = 32 launch -- doubling the
  // resident warps and halving each warp's dependent chain buys nothing, i.e. the kernel is bound by its instruction
  // count, not by latency; the one-warp kernel stays the default.
  static int bwd2_on = -1;
  if (bwd2_on < 0) { const char* e = getenv("BF_ATTN_BWD2"); bwd2_on = (e != nullptr && e[0] == '1') ? 1 : 0; }
  if (bwd && bwd2_on && p.d_qkv_bias == nullptr) {
    const size_t smem2 = 1024 + (size_t)kBwd2Pairs * Bwd2Pair::kBytes + tab_bytes(p.heads) +
                         (size_t)(3 * FD + 33 * p.heads) * sizeof(float);
    BF_REQUIRE(smem2 <= 227 * 1024, "bf_attention (prenorm): shared memory %zu too large (heads=%d)", smem2, p.heads);
    static bool attr2[2] = {false, false};
    if (!attr2[packed]) {
      cudaError_t e = packed ? cudaFuncSetAttribute(attn_fast_bwd2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)
                             : cudaFuncSetAttribute(attn_fast_bwd2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      if (int st_ = check_cuda(e, "cudaFuncSetAttribute(attention fast bwd2)")) return st_;
      attr2[packed] = true;
    }
    const long n_work2 = p.n_tiles * p.heads;
    long blocks2 = (n_work2 + kBwd2Pairs - 1) / kBwd2Pairs;
    if (blocks2 > num_sms()) blocks2 = num_sms();
    if (!packed) {
      while (blocks2 > 1 && (blocks2 * kBwd2Pairs) % p.heads != 0) --blocks2;
      BF_REQUIRE((blocks2 * kBwd2Pairs) % p.heads == 0, "bf_attention_bwd (prenorm): heads=%d does not divide %d warp pairs",
                 p.heads, kBwd2Pairs);
    }
    if (packed) launch_k(attn_fast_bwd2_kernel<true>, dim3((unsigned)blocks2), dim3(kBwd2Pairs * 64), smem2, st, m_qkv, m_b, m_c, p);
    else launch_k(attn_fast_bwd2_kernel<false>, dim3((unsigned)blocks2), dim3(kBwd2Pairs * 64), smem2, st, m_qkv, m_b, m_c, p);
    count_launch();
    return check_cuda(cudaGetLastError(), "attn_fast_bwd2_kernel launch");
  }
  const int warps = bwd ? kBwdWarps : kFwdWarps;
  const size_t smem = 1024 + (size_t)warps * (bwd ? BwdWarp::kBytes : FwdWarp::kBytes) + tab_bytes(p.heads) +
                      (bwd ? (size_t)(3 * FD + 33 * p.heads + 2 + 3 * FD * kBwdWarps) * sizeof(float) : 0);
  BF_REQUIRE(smem <= 227 * 1024, "bf_attention (prenorm): shared memory %zu too large (heads=%d)", smem, p.heads);
  static bool attr_done[4] = {false, false, false, false};
  const int ki = (bwd ? 2 : 0) + (packed ? 1 : 0);
  if (!attr_done[ki]) {
    cudaError_t e;
    if (bwd) e = packed ? cudaFuncSetAttribute(attn_fast_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)
                        : cudaFuncSetAttribute(attn_fast_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    else e = packed ? cudaFuncSetAttribute(attn_fast_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)
                    : cudaFuncSetAttribute(attn_fast_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (int st_ = check_cuda(e, "cudaFuncSetAttribute(attention fast)")) return st_;
    attr_done[ki] = true;
  }
  const long n_work = p.n_tiles * p.heads;
  long blocks = (n_work + warps - 1) / warps;
  if (blocks > num_sms()) blocks = num_sms();
  if (bwd) {
    // a warp must keep one head for the whole launch (work item wi -> head wi % heads, stride gridDim * warps)
    while (blocks > 1 && (blocks * warps) % p.heads != 0) --blocks;
    BF_REQUIRE((blocks * warps) % p.heads == 0, "bf_attention_bwd (prenorm): heads=%d does not divide %d warps", p.heads, warps);
  }
  if (bwd) {
    if (packed) launch_k(attn_fast_bwd_kernel<true>, dim3((unsigned)blocks), dim3(warps * 32), (size_t)(smem), st, m_qkv, m_b, m_c, p);
    else launch_k(attn_fast_bwd_kernel<false>, dim3((unsigned)blocks), dim3(warps * 32), (size_t)(smem), st, m_qkv, m_b, m_c, p);
  } else {
    if (packed) launch_k(attn_fast_fwd_kernel<true>, dim3((unsigned)blocks), dim3(warps * 32), (size_t)(smem), st, m_qkv, m_b, p);
    else launch_k(attn_fast_fwd_kernel<false>, dim3((unsigned)blocks), dim3(warps * 32), (size_t)(smem), st, m_qkv, m_b, p);
  }
  count_launch();
  return check_cuda(cudaGetLastError(), bwd ? "attn_fast_bwd_kernel launch" : "attn_fast_fwd_kernel launch");
}

}  // namespace bf
