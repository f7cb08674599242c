// Fast path of the fused 1-D attention for sequences of 33..64 tokens (head_dim 64) over PRE-NORMALISED q / k: the
// axes of a 1024-pixel domain at patch 16 (BASELINE configs[4]: film_avit_big at 1024 x 1024, h = w = 64), where the
// generic kernel (attention.cu) took 59 % of the training step (1.5 ms per backward launch, 0.37 ms forward).
//
// Same construction as attention_fast.cu (one 128B-swizzled tile per operand loaded by ONE 4-D TMA tensor copy whatever
// the axis, LayerNorm affine applied to the ldmatrix fragments with packed bf16 FMAs, P' = s*P + (1-s)/L folds the
// high-frequency scaling, results leave by TMA store / bf16 reduce-add), with 64-row tiles:
//   forward : one warp per (sequence, head), four 16-row query tiles against all 64 keys; 8 warps per CTA (25 KB each)
//   backward: FOUR warps per (sequence, head) sharing one 44 KB set of tiles -- phase 1 split by query m-tile (warp w owns
//             rows 16w..16w+15: dP = dO V^T, S, softmax, dS; P' replaces V, dS goes to a fifth tile), phase 2 split by
//             output columns (warp w owns columns 16w..16w+15 of dV, dK, dQ, so every warp reads and overwrites only
//             its own columns of the dO / Q tiles); the LayerNorm backward's row sums over the 64 columns are exchanged
//             through shared memory.  Named barriers of 128 threads order the hand-offs; 4 items (16 warps) per CTA.
// Sequences shorter than 64 (L = 33..63) run the same kernels with the padded keys masked (PACKED = true).
#include "attention_fast.cuh"

namespace bf {

constexpr int LT = 64;                       // rows per tile
constexpr int kTile64 = LT * FD * 2;         // 8192 B
constexpr int kBrelStride = 128;             // bias table entries per head (rel + L - 1 in [0, 126])
constexpr int kFwd64Warps = 8;
constexpr int kBwd64Items = 4;               // work items per CTA, 4 warps each
__host__ __device__ constexpr int tab_bytes64(int heads) { return kTabFixed + heads * kBrelStride * 4; }

struct Fwd64Warp {
  static constexpr int kQ = 0, kK = kTile64, kV = 2 * kTile64;
  static constexpr int kBar = 3 * kTile64;
  static constexpr int kRowGp = kBar + 16;                 // int[64]
  static constexpr int kBytes = ((kRowGp + LT * 4) + 1023) / 1024 * 1024;
};
struct Bwd64Item {
  static constexpr int kQ = 0, kK = kTile64, kV = 2 * kTile64, kDo = 3 * kTile64, kDs = 4 * kTile64;
  static constexpr int kBar = 5 * kTile64;
  static constexpr int kRowGp = kBar + 16;                 // int[64]
  static constexpr int kRstd = kRowGp + LT * 4;            // float[64][2]
  static constexpr int kPart = kRstd + LT * 8;             // float[4][64][2]: per-warp partial LayerNorm row sums
  static constexpr int kBytes = ((kPart + 4 * LT * 8) + 1023) / 1024 * 1024;
};

__device__ __forceinline__ void item_sync(int item) {
  asm volatile("bar.sync %0, 128;" ::"r"(item + 1) : "memory");
}

// S(16 x 64) = Q'[m0:m0+16] K'^T with the LayerNorm affine applied to the fragments
__device__ __forceinline__ void scores16x64(float (&acc)[8][4], const uint8_t* sQ, const uint8_t* sK, const uint32_t* pairs,
                                            int m0, int lane) {
  const int t = lane & 3;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) { acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f; }
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
    for (int np = 0; np < 4; ++np) {
      uint32_t b[4];
      frag_b(b, sK, np * 16, ks * 16, lane);
      b[0] = hfma2_bf16(b[0], ak0, bk0); b[1] = hfma2_bf16(b[1], ak1, bk1);
      b[2] = hfma2_bf16(b[2], ak0, bk0); b[3] = hfma2_bf16(b[3], ak1, bk1);
      mma16816(acc[2 * np], a, b[0], b[1]);
      mma16816(acc[2 * np + 1], a, b[2], b[3]);
    }
  }
}

// scores of one 16-row tile -> + bias, mask of the padded keys / rows -> softmax probabilities in place
template <bool PACKED>
__device__ __forceinline__ void softmax16x64(float (&acc)[8][4], const float* brel, int L, int m0, int lane) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int i = m0 + g + half * 8;
    float mx = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = nt * 8 + 2 * t + e;
        float s = acc[nt][half * 2 + e];
        if (PACKED) {
          if (i >= L) s = 0.f;                       // unused query row: harmless uniform row
          else if (j < L) s += brel[j - i + L - 1];
          else s = -INFINITY;                        // padded key
        } else {
          s += brel[j - i + LT - 1];
        }
        acc[nt][half * 2 + e] = s;
        mx = fmaxf(mx, s);
      }
    }
    mx = qmax(mx);
    float sum = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float pe = __expf(acc[nt][half * 2 + e] - mx);
        acc[nt][half * 2 + e] = pe;
        sum += pe;
      }
    }
    const float inv = __fdividef(1.f, qsum(sum));
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) { acc[nt][half * 2] *= inv; acc[nt][half * 2 + 1] *= inv; }
  }
}

__device__ __forceinline__ void stage16x64(uint8_t* tile, int m0, const float (&o)[8][4], int lane) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    *reinterpret_cast<uint32_t*>(tile + swz(m0 + g, nt * 8 + 2 * t)) = pack_bf2(o[nt][0], o[nt][1]);
    *reinterpret_cast<uint32_t*>(tile + swz(m0 + g + 8, nt * 8 + 2 * t)) = pack_bf2(o[nt][2], o[nt][3]);
  }
}
// 16 rows x 16 columns (2 n tiles) at column col0
__device__ __forceinline__ void stage16x16(uint8_t* tile, int m0, int col0, const float (&o)[2][4], int lane) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int nt = 0; nt < 2; ++nt) {
    *reinterpret_cast<uint32_t*>(tile + swz(m0 + g, col0 + nt * 8 + 2 * t)) = pack_bf2(o[nt][0], o[nt][1]);
    *reinterpret_cast<uint32_t*>(tile + swz(m0 + g + 8, col0 + nt * 8 + 2 * t)) = pack_bf2(o[nt][2], o[nt][3]);
  }
}

__device__ __forceinline__ void zero_bytes(uint8_t* my, int bytes, int tid, int nthreads) {
  const uint4 z = make_uint4(0u, 0u, 0u, 0u);
  for (int i = tid * 16; i < bytes; i += nthreads * 16) *reinterpret_cast<uint4*>(my + i) = z;
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
template <bool PACKED>
__global__ void __launch_bounds__(kFwd64Warps * 32, 1)
attn_fast64_fwd_kernel(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_out,
                       const FastParams p) {
  pdl_prologue_done();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* tab = smem + kFwd64Warps * Fwd64Warp::kBytes;
  fill_tables(p, tab, threadIdx.x, blockDim.x, kBrelStride);
  uint8_t* my = smem + warp * Fwd64Warp::kBytes;
  uint8_t* sQ = my + Fwd64Warp::kQ;
  uint8_t* sK = my + Fwd64Warp::kK;
  uint8_t* sV = my + Fwd64Warp::kV;
  uint64_t* bar = reinterpret_cast<uint64_t*>(my + Fwd64Warp::kBar);
  const float* brel_all = reinterpret_cast<const float*>(tab + kTabBrel);
  zero_bytes(my, 3 * kTile64, lane, 32);    // rows beyond L are never written by TMA and must stay finite
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

  for (long wi = (long)blockIdx.x * kFwd64Warps + warp; wi < n_work; wi += (long)gridDim.x * kFwd64Warps) {
    Item it;
    const long tile = wi / p.heads;
    it.head = (int)(wi - tile * p.heads);
    it.s_out = (int)(tile / p.tiles_per_outer);
    it.s_in0 = (int)(tile - (long)it.s_out * p.tiles_per_outer);
    if (lane == 0) {
      tma_store_wait_read<0>();          // the previous item's store has finished reading the q tile
      mbar_arrive_expect_tx(bar, (uint32_t)(3 * L * FD * 2));
      const int col = it.head * 3 * FD;
      tma_load_4d(sQ, &map_qkv, bar, col, 0, it.s_in0, it.s_out);
      tma_load_4d(sK, &map_qkv, bar, col + FD, 0, it.s_in0, it.s_out);
      tma_load_4d(sV, &map_qkv, bar, col + 2 * FD, 0, it.s_in0, it.s_out);
    }
    mbar_wait(bar, phase);
    phase ^= 1u;
    __syncwarp();
    const float* brel = brel_all + it.head * kBrelStride;
    const float sf = p.scale_factor != nullptr ? __ldg(p.scale_factor + it.head) : 1.f;
    const float lowc = (1.f - sf) * invL;
#pragma unroll 1
    for (int mt = 0; mt < LT / 16; ++mt) {
      float acc[8][4];
      scores16x64(acc, sQ, sK, pairs, mt * 16, lane);
      softmax16x64<PACKED>(acc, brel, L, mt * 16, lane);
      // P' = s*P + (1-s)/L inside the sequence, 0 on the padding; straight into A fragments
      uint32_t pa[4][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float pv = fmaf(acc[nt][e], sf, lowc);
          if (PACKED) {
            const int i = mt * 16 + g8 + (e >> 1) * 8, j = nt * 8 + 2 * t + (e & 1);
            if (i >= L || j >= L) pv = 0.f;
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
      for (int kk = 0; kk < LT / 16; ++kk) {
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
      __syncwarp();                     // the q rows of this m tile are dead: stage the output there
      stage16x64(sQ, mt * 16, o, lane);
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      if (p.accumulate) tma_reduce_add_4d(&map_out, sQ, it.head * FD, 0, it.s_in0, it.s_out);
      else tma_store_4d(&map_out, sQ, it.head * FD, 0, it.s_in0, it.s_out);
      tma_store_commit();
    }
  }
  if (lane == 0) tma_store_wait_all();
}

// ---------------------------------------------------------------------------------------------
// backward: four warps per (sequence, head)
// ---------------------------------------------------------------------------------------------
// LayerNorm backward on this warp's 16 columns of a 16-row tile, first half: acc = dL/dy -> dn = dy * w (in place),
// parameter-gradient partials, this warp's share of the row sums (sum dn, sum dn * xhat) -> part[row][2]
template <bool DB, bool SCALE>
__device__ __forceinline__ void ln64_h1(float (&acc)[2][4], float pre, const uint8_t* xh, int m0, int col0, const float* w,
                                        float (&dw)[2][2], float (&db)[2][2], float* part, int lane) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int r = m0 + g + half * 8;
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
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
// second half, once the four warps' partial sums are visible
__device__ __forceinline__ void ln64_h2(float (&acc)[2][4], const uint8_t* xh, const float* rstd_s, int which, int m0,
                                        int col0, const float* part_all, int lane) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int r = m0 + g + half * 8;
    const float rstd = rstd_s[2 * r + which];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) { s1 += part_all[q * 2 * LT + 2 * r]; s2 += part_all[q * 2 * LT + 2 * r + 1]; }
    const float c0 = -rstd * s1 * (1.f / FD);
    const float c1 = -rstd * s2 * (1.f / FD);
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      const float2 xv = unpack2<bf16>(*reinterpret_cast<const uint32_t*>(xh + swz(r, col0 + nt * 8 + 2 * t)));
      acc[nt][half * 2] = fmaf(xv.x, c1, fmaf(acc[nt][half * 2], rstd, c0));
      acc[nt][half * 2 + 1] = fmaf(xv.y, c1, fmaf(acc[nt][half * 2 + 1], rstd, c0));
    }
  }
}

template <bool PACKED>
__global__ void __launch_bounds__(kBwd64Items * 128, 1)
attn_fast64_bwd_kernel(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_do,
                       const __grid_constant__ CUtensorMap map_dqkv, const FastParams p) {
  pdl_prologue_done();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = warp >> 2, w = warp & 3;        // work-item slot of the CTA, warp inside the item
  const int tid_item = (w << 5) | lane;
  uint8_t* tab = smem + kBwd64Items * Bwd64Item::kBytes;
  fill_tables(p, tab, threadIdx.x, blockDim.x, kBrelStride);
  float* s_acc = reinterpret_cast<float*>(tab + tab_bytes64(p.heads));
  float* s_dqw = s_acc;                 // [64]
  float* s_dqb = s_dqw + FD;
  float* s_dkw = s_dqb + FD;
  float* s_demb = s_dkw + FD;           // [32 * heads]
  float* s_dsf = s_demb + 32 * p.heads; // [heads]
  const int n_acc = 3 * FD + 33 * p.heads;
  for (int i = threadIdx.x; i < n_acc; i += blockDim.x) s_acc[i] = 0.f;
  uint8_t* my = smem + item * Bwd64Item::kBytes;
  uint8_t* sQ = my + Bwd64Item::kQ;
  uint8_t* sK = my + Bwd64Item::kK;
  uint8_t* sV = my + Bwd64Item::kV;     // after dP: P'
  uint8_t* sdo = my + Bwd64Item::kDo;
  uint8_t* sdS = my + Bwd64Item::kDs;
  uint64_t* bar = reinterpret_cast<uint64_t*>(my + Bwd64Item::kBar);
  float* rstd_s = reinterpret_cast<float*>(my + Bwd64Item::kRstd);
  float* part_all = reinterpret_cast<float*>(my + Bwd64Item::kPart);
  float* part_me = part_all + w * 2 * LT;
  const float* brel_all = reinterpret_cast<const float*>(tab + kTabBrel);
  zero_bytes(my, 5 * kTile64, tid_item, 128);
  if (w == 0 && lane == 0) {
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
  const int L = p.L;
  const float invL = 1.f / (float)L;
  const long n_work = p.n_tiles * p.heads;
  const int g8 = lane >> 2, t = lane & 3;
  const float qscale = rsqrtf((float)FD);
  const int mt = w;                     // phase 1: this warp's query m-tile
  const int col0 = 16 * w;              // phase 2: this warp's output columns
  float dwq[2][2], dbq[2][2], dwk[2][2], dbk_unused[2][2];
#pragma unroll
  for (int nt = 0; nt < 2; ++nt) { dwq[nt][0] = dwq[nt][1] = dbq[nt][0] = dbq[nt][1] = dwk[nt][0] = dwk[nt][1] = 0.f; }
  // !PACKED (L = 64): the launch guarantees gridDim * items % heads == 0, so an item slot keeps one head and the bias /
  // scale gradients stay in registers: a lane's dS elements of m-tile mt fall on the diagonals
  // rel = 8*q + (2t + e - g),  q = nt - 2*mt - half in [-7, 7]
  float dacc[30];
#pragma unroll
  for (int k = 0; k < 30; ++k) dacc[k] = 0.f;
  float dsf_acc = 0.f;
  const int my_head = (int)(((long)blockIdx.x * kBwd64Items + item) % p.heads);

  for (long wi = (long)blockIdx.x * kBwd64Items + item; wi < n_work; wi += (long)gridDim.x * kBwd64Items) {
    Item it;
    {
      const long tile = wi / p.heads;
      it.head = (int)(wi - tile * p.heads);
      it.s_out = (int)(tile / p.tiles_per_outer);
      it.s_in0 = (int)(tile - (long)it.s_out * p.tiles_per_outer);
    }
    if (w == 0 && lane == 0) {
      tma_store_wait_read<0>();          // the previous item's three stores have finished reading the tiles
      mbar_arrive_expect_tx(bar, (uint32_t)(4 * L * FD * 2));
      const int col = it.head * 3 * FD;
      tma_load_4d(sQ, &map_qkv, bar, col, 0, it.s_in0, it.s_out);
      tma_load_4d(sK, &map_qkv, bar, col + FD, 0, it.s_in0, it.s_out);
      tma_load_4d(sV, &map_qkv, bar, col + 2 * FD, 0, it.s_in0, it.s_out);
      tma_load_4d(sdo, &map_do, bar, it.head * FD, 0, it.s_in0, it.s_out);
    }
    if (tid_item < LT) {                 // rstd of the raw q / k rows of this sequence
      float2 rr = make_float2(0.f, 0.f);
      if (tid_item < L) {
        const long tok = (long)it.s_out * p.outer_stride + (long)it.s_in0 * p.inner_stride + (long)tid_item * p.tok_stride;
        rr = __ldg(reinterpret_cast<const float2*>(p.rstd + (tok * p.heads + it.head) * 2));
      }
      rstd_s[2 * tid_item] = rr.x; rstd_s[2 * tid_item + 1] = rr.y;
    }
    mbar_wait(bar, phase);
    phase ^= 1u;
    item_sync(item);                                         // B1: rstd table visible
    const int head = it.head;
    const float* brel = brel_all + head * kBrelStride;
    const float sf = p.scale_factor != nullptr ? __ldg(p.scale_factor + head) : 1.f;
    const float lowc = (1.f - sf) * invL;

    // ---- phase 1: query rows 16*mt .. 16*mt+15 ----
    {
      float dp[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) { dp[nt][0] = dp[nt][1] = dp[nt][2] = dp[nt][3] = 0.f; }
#pragma unroll
      for (int ks = 0; ks < FD / 16; ++ks) {
        uint32_t a0[4];
        frag_a(a0, sdo, mt * 16, ks * 16, lane);
#pragma unroll
        for (int np = 0; np < 4; ++np) {
          uint32_t b[4];
          frag_b(b, sV, np * 16, ks * 16, lane);
          mma16816(dp[2 * np], a0, b[0], b[1]);
          mma16816(dp[2 * np + 1], a0, b[2], b[3]);
        }
      }
      float acc[8][4];
      scores16x64(acc, sQ, sK, pairs, mt * 16, lane);
      softmax16x64<PACKED>(acc, brel, L, mt * 16, lane);
      item_sync(item);                                       // B2: every warp is done reading v
      float dsf = 0.f;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int i = mt * 16 + g8 + half * 8;
        float dot = 0.f;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            bool valid = true;
            if (PACKED) valid = i < L && (nt * 8 + 2 * t + e) < L;
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
        for (int nt = 0; nt < 8; ++nt) {
          float ds[2], pp[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const float pv = acc[nt][half * 2 + e];
            ds[e] = sf * pv * (dp[nt][half * 2 + e] - dot);
            bool valid = true;
            if (PACKED) valid = i < L && (nt * 8 + 2 * t + e) < L;
            pp[e] = valid ? fmaf(pv, sf, lowc) : 0.f;
          }
          *reinterpret_cast<uint32_t*>(sV + swz(i, nt * 8 + 2 * t)) = pack_bf2(pp[0], pp[1]);
          *reinterpret_cast<uint32_t*>(sdS + swz(i, nt * 8 + 2 * t)) = pack_bf2(ds[0], ds[1]);
          if (!PACKED) {
            dacc[(nt - 2 * mt - half + 7) * 2] += ds[0];
            dacc[(nt - 2 * mt - half + 7) * 2 + 1] += ds[1];
          }
        }
      }
      if (PACKED) {
        if (p.d_scale_factor != nullptr) {
          dsf = warp_sum(dsf);
          if (lane == 0) atomicAdd(s_dsf + head, dsf);
        }
      } else {
        dsf_acc += dsf;
      }
    }
    item_sync(item);                                         // B3: all of P' and dS is in shared memory
    if (PACKED && p.d_bias_emb != nullptr) {
      // bias-embedding gradient: sum of dS along the diagonals (L < 64: no fixed head per slot)
      if (tid_item < 2 * L - 1) {
        const int r = tid_item;
        float s = 0.f;
        for (int i = 0; i < L; ++i) {
          const int j = i + r - (L - 1);
          if (j >= 0 && j < L) s += __bfloat162float(*reinterpret_cast<const bf16*>(sdS + swz(i, j)));
        }
        atomicAdd(s_demb + __ldg(p.bucket + r) * p.heads + head, s);
      }
    }

    // ---- phase 2: output columns col0 .. col0+15 ----
    float o[4][2][4];
    // dV = P'^T dO'
#pragma unroll
    for (int m = 0; m < 4; ++m) {
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) { o[m][nt][0] = o[m][nt][1] = o[m][nt][2] = o[m][nt][3] = 0.f; }
    }
#pragma unroll
    for (int kk = 0; kk < LT / 16; ++kk) {
      uint32_t b[4];
      frag_b_t(b, sdo, col0, kk * 16, lane);
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        uint32_t a[4];
        frag_a_t(a, sV, m * 16, kk * 16, lane);
        mma16816(o[m][0], a, b[0], b[1]);
        mma16816(o[m][1], a, b[2], b[3]);
      }
    }
    __syncwarp();                       // this warp's columns of dO are dead: they become its staging columns
#pragma unroll
    for (int m = 0; m < 4; ++m) {
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) o[m][nt][e] *= p.out_scale;
      }
      stage16x16(sdo, m * 16, col0, o[m], lane);
    }
    fence_proxy_async();
    item_sync(item);                                         // B4: dV staged by the four warps
    if (w == 0 && lane == 0) {
      if (p.accumulate) tma_reduce_add_4d(&map_dqkv, sdo, head * 3 * FD + 2 * FD, 0, it.s_in0, it.s_out);
      else tma_store_4d(&map_dqkv, sdo, head * 3 * FD + 2 * FD, 0, it.s_in0, it.s_out);
      tma_store_commit();
    }
    // dK' = dS^T Q'  -> LayerNorm backward -> d(raw k)
#pragma unroll
    for (int m = 0; m < 4; ++m) {
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) { o[m][nt][0] = o[m][nt][1] = o[m][nt][2] = o[m][nt][3] = 0.f; }
    }
#pragma unroll
    for (int kk = 0; kk < LT / 16; ++kk) {
      uint32_t b[4];
      frag_b_t(b, sQ, col0, kk * 16, lane);                  // Q as [k = i][n = d]: both halves of a register share d
      const uint32_t al = splat[col0 + g8], ah = splat[col0 + 8 + g8];
      const uint32_t bl = splat[64 + col0 + g8], bh = splat[64 + col0 + 8 + g8];
      b[0] = hfma2_bf16(b[0], al, bl); b[1] = hfma2_bf16(b[1], al, bl);
      b[2] = hfma2_bf16(b[2], ah, bh); b[3] = hfma2_bf16(b[3], ah, bh);
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        uint32_t a[4];
        frag_a_t(a, sdS, m * 16, kk * 16, lane);
        mma16816(o[m][0], a, b[0], b[1]);
        mma16816(o[m][1], a, b[2], b[3]);
      }
    }
#pragma unroll
    for (int m = 0; m < 4; ++m) ln64_h1<false, false>(o[m], 1.f, sK, m * 16, col0, wk, dwk, dbk_unused, part_me, lane);
    if (w == 0 && lane == 0) tma_store_wait_read<0>();       // the dV store has finished reading the staging tile
    item_sync(item);                                         // B5: partial row sums exchanged, staging tile free
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      ln64_h2(o[m], sK, rstd_s, 1, m * 16, col0, part_all, lane);
      stage16x16(sdo, m * 16, col0, o[m], lane);
    }
    fence_proxy_async();
    item_sync(item);                                         // B6: dK staged (and the partial sums consumed)
    if (w == 0 && lane == 0) {
      if (p.accumulate) tma_reduce_add_4d(&map_dqkv, sdo, head * 3 * FD + FD, 0, it.s_in0, it.s_out);
      else tma_store_4d(&map_dqkv, sdo, head * 3 * FD + FD, 0, it.s_in0, it.s_out);
      tma_store_commit();
    }
    // dQ' = dS K'  -> LayerNorm backward -> d(raw q), staged over the q columns this warp owns
#pragma unroll
    for (int m = 0; m < 4; ++m) {
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) { o[m][nt][0] = o[m][nt][1] = o[m][nt][2] = o[m][nt][3] = 0.f; }
    }
#pragma unroll
    for (int kk = 0; kk < LT / 16; ++kk) {
      uint32_t b[4];
      frag_b_t(b, sK, col0, kk * 16, lane);
      const uint32_t al = splat[128 + col0 + g8], ah = splat[128 + col0 + 8 + g8];
      const uint32_t bl = splat[192 + col0 + g8], bh = splat[192 + col0 + 8 + g8];
      b[0] = hfma2_bf16(b[0], al, bl); b[1] = hfma2_bf16(b[1], al, bl);
      b[2] = hfma2_bf16(b[2], ah, bh); b[3] = hfma2_bf16(b[3], ah, bh);
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        uint32_t a[4];
        frag_a(a, sdS, m * 16, kk * 16, lane);
        mma16816(o[m][0], a, b[0], b[1]);
        mma16816(o[m][1], a, b[2], b[3]);
      }
    }
#pragma unroll
    for (int m = 0; m < 4; ++m) ln64_h1<true, true>(o[m], qscale, sQ, m * 16, col0, wq, dwq, dbq, part_me, lane);
    item_sync(item);                                         // B7: partial row sums exchanged
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      ln64_h2(o[m], sQ, rstd_s, 0, m * 16, col0, part_all, lane);
      __syncwarp();                     // all lanes have read the xhat values of these rows / columns
      stage16x16(sQ, m * 16, col0, o[m], lane);
    }
    fence_proxy_async();
    item_sync(item);                                         // B8: dQ staged; tables / partials free for the next item
    if (w == 0 && lane == 0) {
      if (p.accumulate) tma_reduce_add_4d(&map_dqkv, sQ, head * 3 * FD, 0, it.s_in0, it.s_out);
      else tma_store_4d(&map_dqkv, sQ, head * 3 * FD, 0, it.s_in0, it.s_out);
      tma_store_commit();
    }
  }
  if (w == 0 && lane == 0) tma_store_wait_all();
  if (!PACKED) {
    if (p.d_bias_emb != nullptr) {
#pragma unroll
      for (int k = 0; k < 30; ++k) {
        const int rel = 8 * ((k >> 1) - 7) + 2 * t + (k & 1) - g8;
        if (rel > -LT && rel < LT) atomicAdd(s_demb + __ldg(p.bucket + rel + LT - 1) * p.heads + my_head, dacc[k]);
      }
    }
    if (p.d_scale_factor != nullptr) {
      dsf_acc = warp_sum(dsf_acc);
      if (lane == 0) atomicAdd(s_dsf + my_head, dsf_acc);
    }
  }
  // LayerNorm parameter gradients: lanes with equal t hold partial sums of the same columns
#pragma unroll
  for (int nt = 0; nt < 2; ++nt) {
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
// 4-D view (column, position, sequence, outer) of a token-major (tokens, width) bf16 matrix; box = 64 x L x 1 x 1
static int make_seq_map64(CUtensorMap* map, const void* base, long ld, int width, const bf_attn_args* a) {
  const long n_outer = a->n_seq / a->inner;
  uint64_t dims[4] = {(uint64_t)width, (uint64_t)a->L, (uint64_t)a->inner, (uint64_t)n_outer};
  uint64_t str[3] = {(uint64_t)a->tok_stride * ld * 2, (uint64_t)a->inner_stride * ld * 2, (uint64_t)a->outer_stride * ld * 2};
  uint32_t box[4] = {FD, (uint32_t)a->L, 1, 1};
  return make_map(map, BF_BF16, base, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

int launch_attn_fast64(const bf_attn_args* a, bool bwd, cudaStream_t st) {
  BF_REQUIRE(a->head_dim == FD && a->L > 32 && a->L <= LT,
             "bf_attention (prenorm, 64-row tiles): head_dim 64 and 32 < L <= 64 (got d=%d L=%d)", a->head_dim, a->L);
  BF_REQUIRE(!bwd || a->rstd != nullptr, "bf_attention_bwd (prenorm): rstd required");
  BF_REQUIRE(a->n_seq % a->inner == 0, "bf_attention (prenorm): n_seq=%ld must be a multiple of inner=%ld",
             (long)a->n_seq, (long)a->inner);
  BF_REQUIRE(a->inner < (1l << 31) && a->n_seq / a->inner < (1l << 31), "bf_attention (prenorm): geometry too large");
  BF_REQUIRE(!bwd || a->d_qkv_bias == nullptr, "bf_attention_bwd (prenorm, 64-row tiles): d_qkv_bias is not fused here");
  FastParams p{};
  p.rstd = a->rstd;
  p.heads = a->heads; p.L = a->L;
  p.G = 1;
  p.inner = (int)a->inner;
  p.tiles_per_outer = (int)a->inner;
  p.n_tiles = a->n_seq;
  p.outer_stride = a->outer_stride; p.inner_stride = a->inner_stride; p.tok_stride = a->tok_stride;
  p.qn_w = a->qn_w; p.qn_b = a->qn_b; p.kn_w = a->kn_w; p.kn_b = a->kn_b;
  p.bias_emb = a->bias_emb; p.bucket = a->bucket; p.scale_factor = a->scale_factor;
  p.out_scale = a->out_scale; p.accumulate = a->accumulate;
  p.d_qn_w = a->d_qn_w; p.d_qn_b = a->d_qn_b; p.d_kn_w = a->d_kn_w; p.d_kn_b = a->d_kn_b;
  p.d_bias_emb = a->d_bias_emb; p.d_scale_factor = a->d_scale_factor;
  p.d_qkv_bias = nullptr;
  bool packed = a->L != LT;
  const int E3 = 3 * FD * a->heads;
  CUtensorMap m_qkv, m_b, m_c;
  if (int e = make_seq_map64(&m_qkv, a->qkv, a->ld_qkv, E3, a)) return e;
  if (bwd) {
    if (int e = make_seq_map64(&m_b, a->dout, a->ld_dout, E3 / 3, a)) return e;
    if (int e = make_seq_map64(&m_c, a->out, a->ld_out, E3, a)) return e;
  } else {
    if (int e = make_seq_map64(&m_b, a->out, a->ld_out, E3 / 3, a)) return e;
  }
  const long n_work = p.n_tiles * p.heads;
  static bool attr_done[4] = {false, false, false, false};
  const int ki = (bwd ? 2 : 0) + (packed ? 1 : 0);
  if (!attr_done[ki]) {
    cudaError_t e;
    if (bwd) e = packed ? cudaFuncSetAttribute(attn_fast64_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)
                        : cudaFuncSetAttribute(attn_fast64_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    else e = packed ? cudaFuncSetAttribute(attn_fast64_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)
                    : cudaFuncSetAttribute(attn_fast64_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (int st_ = check_cuda(e, "cudaFuncSetAttribute(attention fast64)")) return st_;
    attr_done[ki] = true;
  }
  if (!bwd) {
    const size_t smem = 1024 + (size_t)kFwd64Warps * Fwd64Warp::kBytes + tab_bytes64(p.heads);
    BF_REQUIRE(smem <= 227 * 1024, "bf_attention (prenorm, 64-row tiles): shared memory %zu too large (heads=%d)", smem, p.heads);
    long blocks = (n_work + kFwd64Warps - 1) / kFwd64Warps;
    if (blocks > num_sms()) blocks = num_sms();
    if (packed) launch_k(attn_fast64_fwd_kernel<true>, dim3((unsigned)blocks), dim3(kFwd64Warps * 32), smem, st, m_qkv, m_b, p);
    else launch_k(attn_fast64_fwd_kernel<false>, dim3((unsigned)blocks), dim3(kFwd64Warps * 32), smem, st, m_qkv, m_b, p);
    count_launch();
    return check_cuda(cudaGetLastError(), "attn_fast64_fwd_kernel launch");
  }
  const size_t smem = 1024 + (size_t)kBwd64Items * Bwd64Item::kBytes + tab_bytes64(p.heads) +
                      (size_t)(3 * FD + 33 * p.heads) * sizeof(float);
  BF_REQUIRE(smem <= 227 * 1024, "bf_attention_bwd (prenorm, 64-row tiles): shared memory %zu too large (heads=%d)", smem, p.heads);
  long blocks = (n_work + kBwd64Items - 1) / kBwd64Items;
  if (blocks > num_sms()) blocks = num_sms();
  if (!packed) {
    // an item slot must keep one head for the whole launch (work item wi -> head wi % heads, stride gridDim * items);
    // when no grid size allows that, the masked variant (gradients of the bias table through shared memory) runs
    long b2 = blocks;
    while (b2 > 1 && (b2 * kBwd64Items) % p.heads != 0) --b2;
    if ((b2 * kBwd64Items) % p.heads == 0 && 2 * b2 >= blocks) blocks = b2; else packed = true;
  }
  {
    const int ki2 = 2 + (packed ? 1 : 0);
    if (!attr_done[ki2]) {
      cudaError_t e = packed ? cudaFuncSetAttribute(attn_fast64_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)
                             : cudaFuncSetAttribute(attn_fast64_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      if (int st_ = check_cuda(e, "cudaFuncSetAttribute(attention fast64)")) return st_;
      attr_done[ki2] = true;
    }
  }
  if (packed) launch_k(attn_fast64_bwd_kernel<true>, dim3((unsigned)blocks), dim3(kBwd64Items * 128), smem, st, m_qkv, m_b, m_c, p);
  else launch_k(attn_fast64_bwd_kernel<false>, dim3((unsigned)blocks), dim3(kBwd64Items * 128), smem, st, m_qkv, m_b, m_c, p);
  count_launch();
  return check_cuda(cudaGetLastError(), "attn_fast64_bwd_kernel launch");
}

}  // namespace bf
