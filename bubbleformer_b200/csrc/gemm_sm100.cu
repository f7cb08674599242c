// tcgen05 / TMEM / TMA GEMM for sm_100a with fused epilogues -- the contraction engine of the
// FiLMAViT hot path (see include/bubbleformer_b200.h, bf_gemm).
//
// Structure (one CTA per SM, persistent over output tiles, 128 x BN tile, BK = 64):
//   warp 0      : TMA producer   (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier tx)
//   warp 1      : MMA issuer     (tcgen05.mma cta_group::1 kind::f16, fp32 accumulators in TMEM,
//                                 tcgen05.commit releases smem slots / publishes accumulators)
//   warps 2..9  : epilogue       (tcgen05.ld 32x32b -> registers -> fused epilogue -> global)
// TMEM holds two accumulator stages so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// Operand layouts: K-major (row-major (rows, K)) or MN-major (row-major (K, rows)) for both A and B,
// so forward, dgrad (B = W read K x N) and wgrad (A = dY^T, B = X, contraction over tokens, split-K
// with fp32 atomics) all run through this kernel without any transposed copy in HBM.
// A can also be an implicit 2x2/stride-2 patch gather expressed purely as a 4-D TMA tensor map.
#include <cstdarg>
#include <cstdio>
#include <mutex>

#include "common.cuh"

namespace bf {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kEpiWarps = 8;
constexpr int kThreads = 64 + 32 * kEpiWarps;

struct GemmParams {
  int M, N, K;
  int num_m_blocks, num_n_blocks, split_k;
  int k_iters;        // k iterations (of BK) per output tile (per split)
  int k_seg_iters;    // S2D: iterations per ky segment; otherwise == total iterations
  int s2d;            // 1: A coords are 4-D (k, xo, ky, row), B coords 3-D (k, ky, n)
  int s2d_box_w;      // xo extent of the A box (divides 128)
  int s2d_wo;         // output width
  uint32_t idesc;
  int is_f16;
  int epilogue;
  int rows_per_group;
  int d2s_h, d2s_w, d2s_cout;
  const float* bias;
  const float* col_scale;
  const float* col_shift;
  const float* col_gamma;
  const float* row_scale;
  const float* in32;
  const void* aux16;
  void* out16;
  void* out16b;
  float* out32;
  long ldo, ld32;
};

template <int BN>
struct SmemLayout {
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BN <= 128) ? 5 : 4;
  static constexpr int kTileBytes = kStages * kStageBytes;
  static constexpr int kScratchBytes = kEpiWarps * 32 * 33 * 4;  // per-warp 32x33 fp32 transpose scratch
  static constexpr int kBarBytes = (2 * kStages + 4) * 8 + 16;
  static constexpr int kTotal = kTileBytes + kScratchBytes + kBarBytes + 1024;  // + slack for 1 KiB alignment
  static constexpr int kTmemCols = (2 * BN <= 128) ? 128 : (2 * BN <= 256 ? 256 : 512);
};

// ---------------------------------------------------------------------------------------------
// Epilogue on one 32-row x 32-column chunk owned by one warp.
//
// tcgen05.ld hands every thread one ROW of the chunk (32 consecutive fp32 columns).  Touching global
// memory in that shape makes each warp instruction hit 32 different rows, so all global traffic goes
// through a per-warp 32x33 fp32 shared-memory transpose instead: with 8 lanes per row (fp32) or 4 lanes
// per row (16-bit) every instruction reads / writes whole 128-byte / 64-byte row segments.
// The +1 padding makes both the row-wise and the transposed accesses bank-conflict free.
// ---------------------------------------------------------------------------------------------
struct ChunkCtx {
  float* scratch;      // this warp's [32][33] floats
  int lane;
  int rows_valid;      // rows of the chunk inside the matrix (0..32), warp uniform
  int ncols;           // columns of the chunk inside the matrix (1..32), warp uniform
};

__device__ __forceinline__ void regs_to_scratch(const ChunkCtx& c, const float (&v)[32]) {
#pragma unroll
  for (int j = 0; j < 32; ++j) c.scratch[c.lane * 33 + j] = v[j];
  __syncwarp();
}
__device__ __forceinline__ void scratch_to_regs(const ChunkCtx& c, float (&v)[32]) {
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = c.scratch[c.lane * 33 + j];
  __syncwarp();
}

// global (rows at base + r*ld) fp32 -> thread-row registers
__device__ __forceinline__ void coop_load32(const ChunkCtx& c, const float* base, long ld, float (&v)[32]) {
  const bool vec = (c.ncols == 32) && (ld % 4 == 0);
  if (vec) {
    const int cc = (c.lane & 7) * 4;
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      const int r = p * 4 + (c.lane >> 3);
      float4 u = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < c.rows_valid) u = *reinterpret_cast<const float4*>(base + (long)r * ld + cc);
      float* s = c.scratch + r * 33 + cc;
      s[0] = u.x; s[1] = u.y; s[2] = u.z; s[3] = u.w;
    }
  } else {
    for (int r = 0; r < c.rows_valid; ++r)
      c.scratch[r * 33 + c.lane] = (c.lane < c.ncols) ? base[(long)r * ld + c.lane] : 0.f;
  }
  scratch_to_regs(c, v);
}

__device__ __forceinline__ void coop_store32(const ChunkCtx& c, float* base, long ld, const float (&v)[32]) {
  regs_to_scratch(c, v);
  const bool vec = (c.ncols == 32) && (ld % 4 == 0);
  if (vec) {
    const int cc = (c.lane & 7) * 4;
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      const int r = p * 4 + (c.lane >> 3);
      const float* s = c.scratch + r * 33 + cc;
      if (r < c.rows_valid) *reinterpret_cast<float4*>(base + (long)r * ld + cc) = make_float4(s[0], s[1], s[2], s[3]);
    }
  } else {
    for (int r = 0; r < c.rows_valid; ++r)
      if (c.lane < c.ncols) base[(long)r * ld + c.lane] = c.scratch[r * 33 + c.lane];
  }
  __syncwarp();
}

__device__ __forceinline__ void coop_atomic32(const ChunkCtx& c, float* base, long ld, const float (&v)[32]) {
  regs_to_scratch(c, v);
  for (int r = 0; r < c.rows_valid; ++r)
    if (c.lane < c.ncols) atomicAdd(base + (long)r * ld + c.lane, c.scratch[r * 33 + c.lane]);
  __syncwarp();
}

// 16-bit rows; row r of the chunk lives at rowptr(r)
template <typename T16, typename RowPtr>
__device__ __forceinline__ void coop_store16(const ChunkCtx& c, RowPtr rowptr, bool aligned, const float (&v)[32]) {
  regs_to_scratch(c, v);
  if (aligned && c.ncols == 32) {
    const int cc = (c.lane & 3) * 8;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const int r = p * 8 + (c.lane >> 2);
      const float* s = c.scratch + r * 33 + cc;
      if (r < c.rows_valid) {
        uint4 u;
        u.x = pack2<T16>(s[0], s[1]); u.y = pack2<T16>(s[2], s[3]);
        u.z = pack2<T16>(s[4], s[5]); u.w = pack2<T16>(s[6], s[7]);
        *reinterpret_cast<uint4*>(rowptr(r) + cc) = u;
      }
    }
  } else {
    for (int r = 0; r < c.rows_valid; ++r)
      if (c.lane < c.ncols) rowptr(r)[c.lane] = from_f32<T16>(c.scratch[r * 33 + c.lane]);
  }
  __syncwarp();
}

template <typename T16>
__device__ __forceinline__ void coop_load16(const ChunkCtx& c, const T16* base, long ld, float (&v)[32]) {
  if ((ld % 8 == 0) && c.ncols == 32) {
    const int cc = (c.lane & 3) * 8;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const int r = p * 8 + (c.lane >> 2);
      uint4 u = make_uint4(0u, 0u, 0u, 0u);
      if (r < c.rows_valid) u = *reinterpret_cast<const uint4*>(base + (long)r * ld + cc);
      const float2 a = unpack2<T16>(u.x), b = unpack2<T16>(u.y), d = unpack2<T16>(u.z), e = unpack2<T16>(u.w);
      float* s = c.scratch + r * 33 + cc;
      s[0] = a.x; s[1] = a.y; s[2] = b.x; s[3] = b.y; s[4] = d.x; s[5] = d.y; s[6] = e.x; s[7] = e.y;
    }
  } else {
    for (int r = 0; r < c.rows_valid; ++r)
      c.scratch[r * 33 + c.lane] = (c.lane < c.ncols) ? to_f32<T16>(base[(long)r * ld + c.lane]) : 0.f;
  }
  scratch_to_regs(c, v);
}

// m_base: first global row of the chunk (warp uniform), n: first global column.  Thread `lane` holds row m_base+lane.
template <typename T16>
__device__ __forceinline__ void epilogue_chunk(const GemmParams& p, const ChunkCtx& c, int m_base, int n, float (&acc)[32]) {
  const bool a16 = (p.ldo % 8 == 0);
  if (p.bias != nullptr) {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < c.ncols) acc[j] += __ldg(p.bias + n + j);
  }
  auto row16 = [&](void* ptr) {
    T16* b = reinterpret_cast<T16*>(ptr) + (long)m_base * p.ldo + n;
    const long ld = p.ldo;
    return [b, ld](int r) { return b + (long)r * ld; };
  };
  switch (p.epilogue) {
    case BF_EPI_STORE16: {
      coop_store16<T16>(c, row16(p.out16), a16, acc);
      break;
    }
    case BF_EPI_STORE32: {
      coop_store32(c, p.out32 + (long)m_base * p.ld32 + n, p.ld32, acc);
      break;
    }
    case BF_EPI_GELU: {
      if (p.out16b != nullptr) coop_store16<T16>(c, row16(p.out16b), a16, acc);
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[j] = gelu_erf(acc[j]);
      coop_store16<T16>(c, row16(p.out16), a16, acc);
      break;
    }
    case BF_EPI_RESID: {
      if (p.out16b != nullptr) coop_store16<T16>(c, row16(p.out16b), a16, acc);
      const int m = m_base + c.lane;
      const float rs = (p.row_scale != nullptr && c.lane < c.rows_valid) ? __ldg(p.row_scale + m / p.rows_per_group) : 1.f;
      float xin[32];
      coop_load32(c, p.in32 + (long)m_base * p.ld32 + n, p.ld32, xin);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (j < c.ncols) {
          float v = acc[j];
          if (p.col_scale != nullptr) v = fmaf(v, __ldg(p.col_scale + n + j), __ldg(p.col_shift + n + j));
          acc[j] = fmaf(rs * __ldg(p.col_gamma + n + j), v, xin[j]);
        }
      }
      coop_store32(c, p.out32 + (long)m_base * p.ld32 + n, p.ld32, acc);
      if (p.out16 != nullptr) coop_store16<T16>(c, row16(p.out16), a16, acc);
      break;
    }
    case BF_EPI_DGELU: {
      float pre[32];
      coop_load16<T16>(c, reinterpret_cast<const T16*>(p.aux16) + (long)m_base * p.ldo + n, p.ldo, pre);
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[j] *= gelu_erf_grad(pre[j]);
      coop_store16<T16>(c, row16(p.out16), a16, acc);
      break;
    }
    case BF_EPI_ACC32: {
      float g[32];
      coop_load32(c, p.in32 + (long)m_base * p.ld32 + n, p.ld32, g);
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[j] += g[j];
      coop_store32(c, p.out32 + (long)m_base * p.ld32 + n, p.ld32, acc);
      break;
    }
    case BF_EPI_ATOMIC32: {
      coop_atomic32(c, p.out32 + (long)m_base * p.ld32 + n, p.ld32, acc);
      break;
    }
    case BF_EPI_D2S: {
      // m = (img, y, x), n = (ky, kx, co) -> out[((img*2h + 2y+ky)*2w + 2x+kx)*cout + co]
      const int w = p.d2s_w, h = p.d2s_h, co = p.d2s_cout;
      T16* out = reinterpret_cast<T16*>(p.out16);
      if (co % 32 == 0) {
        const int q = n / co, c0 = n - q * co;
        auto rowptr = [=](int r) {
          const int m = m_base + r;
          const int x = m % w, y = (m / w) % h, img = m / (w * h);
          const long pix = ((long)(img * 2 * h + 2 * y + (q >> 1)) * (2 * w) + 2 * x + (q & 1));
          return out + pix * co + c0;
        };
        coop_store16<T16>(c, rowptr, true, acc);
      } else {
        const int m = m_base + c.lane;
        if (c.lane < c.rows_valid) {
          const int x = m % w, y = (m / w) % h, img = m / (w * h);
          for (int j = 0; j < c.ncols; ++j) {
            const int q = (n + j) / co, c0 = (n + j) - q * co;
            const long pix = ((long)(img * 2 * h + 2 * y + (q >> 1)) * (2 * w) + 2 * x + (q & 1));
            out[pix * co + c0] = from_f32<T16>(acc[j]);
          }
        }
      }
      break;
    }
    default: break;
  }
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                    const GemmParams p) {
  using L = SmemLayout<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  float* scratch_all = reinterpret_cast<float*>(smem + L::kTileBytes);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kTileBytes + L::kScratchBytes);
  uint64_t* empty_bar = full_bar + L::kStages;
  uint64_t* tmem_full = empty_bar + L::kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    for (int s = 0; s < L::kStages; ++s) {
      mbar_init(full_bar + s, 1);
      mbar_init(empty_bar + s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tmem_full + s, 1);
      mbar_init(tmem_empty + s, kEpiWarps);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, L::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int tiles_mn = p.num_m_blocks * p.num_n_blocks;
  const int num_tiles = tiles_mn * p.split_k;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int ks = tile / tiles_mn;
        const int mn = tile - ks * tiles_mn;
        const int m_blk = mn / p.num_n_blocks;
        const int n_blk = mn - m_blk * p.num_n_blocks;
        const int m0 = m_blk * BM, n0 = n_blk * BN;
        for (int it = 0; it < p.k_iters; ++it) {
          mbar_wait(empty_bar + stage, phase ^ 1u);
          uint8_t* sa = smem + stage * L::kStageBytes;
          uint8_t* sb = sa + L::kABytes;
          mbar_arrive_expect_tx(full_bar + stage, L::kStageBytes);
          const int kit = ks * p.k_iters + it;      // global k iteration
          if (p.s2d) {
            const int ky = kit / p.k_seg_iters;
            const int kc = (kit - ky * p.k_seg_iters) * BK;
            // rows of the tile are output pixels (row-of-images, xo); box = (BK, box_w, 1, 128/box_w)
            const int xo0 = m0 % p.s2d_wo;
            const int r0 = m0 / p.s2d_wo;
            tma_load_4d(sa, &map_a, full_bar + stage, kc, xo0, ky, r0);
            tma_load_3d(sb, &map_b, full_bar + stage, kc, ky, n0);
          } else {
            const int k0 = kit * BK;
            if (A_MN) {
              tma_load_2d(sa, &map_a, full_bar + stage, m0, k0);
              tma_load_2d(sa + L::kABytes / 2, &map_a, full_bar + stage, m0 + 64, k0);
            } else {
              tma_load_2d(sa, &map_a, full_bar + stage, k0, m0);
            }
            if (B_MN) {
#pragma unroll
              for (int j = 0; j < BN / 64; ++j)
                tma_load_2d(sb + j * (64 * BK * 2), &map_b, full_bar + stage, n0 + 64 * j, k0);
            } else {
              tma_load_2d(sb, &map_b, full_bar + stage, k0, n0);
            }
          }
          if (++stage == L::kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(tmem_empty + as, aphase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * BN);
        for (int it = 0; it < p.k_iters; ++it) {
          mbar_wait(full_bar + stage, phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * L::kStageBytes);
          const uint32_t sb = sa + L::kABytes;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // K-major: advance 16 elements = 32 B inside the 128 B swizzle row.
            // MN-major: advance 16 K rows = two 8-row groups of 1024 B.
            const uint64_t da = A_MN ? make_smem_desc(sa + k * 2048, 64 * BK * 2, 1024)
                                     : make_smem_desc(sa + k * 32, 16, 1024);
            const uint64_t db = B_MN ? make_smem_desc(sb + k * 2048, 64 * BK * 2, 1024)
                                     : make_smem_desc(sb + k * 32, 16, 1024);
            umma_f16(d_tmem, da, db, p.idesc, (it | k) != 0 ? 1u : 0u);
          }
          umma_commit(empty_bar + stage);
          if (++stage == L::kStages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(tmem_full + as);
        if (++as == 2) { as = 0; aphase ^= 1u; }
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int ew = warp - 2;
    const int quad = warp & 3;              // TMEM lane quadrant this warp may access
    const int half = ew >> 2;               // which half of the tile's columns
    ChunkCtx cc;
    cc.scratch = scratch_all + ew * (32 * 33);
    cc.lane = lane;
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int ks = tile / tiles_mn;
      const int mn = tile - ks * tiles_mn;
      const int m_blk = mn / p.num_n_blocks;
      const int n_blk = mn - m_blk * p.num_n_blocks;
      const int m_base = m_blk * BM + quad * 32;
      const int n0 = n_blk * BN;
      cc.rows_valid = min(32, max(0, p.M - m_base));
      mbar_wait(tmem_full + as, aphase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(as * BN);
#pragma unroll 1
      for (int c = half * (BN / 2); c < (half + 1) * (BN / 2); c += 32) {
        float acc[32];
        tmem_ld_32x32(t_row + static_cast<uint32_t>(c), acc);
        tmem_ld_wait();
        cc.ncols = min(32, p.N - (n0 + c));
        if (cc.rows_valid > 0 && cc.ncols > 0) {
          if (p.is_f16) epilogue_chunk<__half>(p, cc, m_base, n0 + c, acc);
          else          epilogue_chunk<__nv_bfloat16>(p, cc, m_base, n0 + c, acc);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty + as);
      if (++as == 2) { as = 0; aphase ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, L::kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  });
  return fn;
}

// rank-d tensor map over 16-bit elements, 128B swizzle, zero fill out of bounds
static int make_map(CUtensorMap* map, int dtype, const void* base, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes /* rank-1 entries */, const uint32_t* box) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return BF_ERR_CUDA;
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  CUresult r = fn(map, dtype == BF_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                  static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdim, gstr, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d): rank %d dims [%llu,%llu,%llu,%llu] box [%u,%u,%u,%u]",
              (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
              (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0), box[0],
              rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
    return BF_ERR_CUDA;
  }
  return BF_OK;
}

template <int BN, bool A_MN, bool B_MN>
static int launch(const CUtensorMap& ma, const CUtensorMap& mb, const GemmParams& p, cudaStream_t st) {
  using L = SmemLayout<BN>;
  auto kern = gemm_tcgen05_kernel<BN, A_MN, B_MN>;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [&] {
    attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal);
  });
  if (attr_err != cudaSuccess) return check_cuda(attr_err, "cudaFuncSetAttribute(gemm)");
  const int tiles = p.num_m_blocks * p.num_n_blocks * p.split_k;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  kern<<<grid, kThreads, L::kTotal, st>>>(ma, mb, p);
  count_launch();
  return check_cuda(cudaGetLastError(), "gemm_tcgen05_kernel launch");
}

static int pick_bn(const bf_gemm_args& a) {
  if (a.bn == 64 || a.bn == 128 || a.bn == 192 || a.bn == 256) return a.bn;   // caller override (tuning)
  const int N = a.N;
  if (N <= 64) return 64;
  if (N <= 128) return 128;
  if (a.epilogue == BF_EPI_ATOMIC32) return N % 128 == 0 ? 128 : (N % 192 == 0 ? 192 : 128);
  if (N % 256 == 0 && N >= 512) return 256;
  if (N % 192 == 0) return 192;
  if (N % 128 == 0) return 128;
  // ragged N: least padded columns, larger tile on ties
  int best = 256, waste = (256 - N % 256) % 256;
  const int cand[2] = {192, 128};
  for (int c : cand) {
    const int w = (c - N % c) % c;
    if (w < waste) { waste = w; best = c; }
  }
  return best;
}

}  // namespace bf

using namespace bf;

extern "C" int bf_gemm(const bf_gemm_args* a, void* stream) {
  BF_REQUIRE(a != nullptr, "bf_gemm: null args");
  BF_REQUIRE(a->M > 0 && a->N > 0 && a->K > 0, "bf_gemm: empty problem M=%d N=%d K=%d", a->M, a->N, a->K);
  BF_REQUIRE(a->dtype == BF_BF16 || a->dtype == BF_F16, "bf_gemm: dtype %d", a->dtype);
  BF_REQUIRE(a->A && a->B, "bf_gemm: null operand");
  BF_REQUIRE((reinterpret_cast<uintptr_t>(a->A) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->B) & 15) == 0,
             "bf_gemm: operands must be 16-byte aligned");
  const bool a_mn = a->a_mode == BF_A_KM;
  const bool b_mn = a->b_mode == BF_B_KN;
  const bool s2d = a->a_mode == BF_A_S2D;
  BF_REQUIRE(!(a_mn && !b_mn), "bf_gemm: A_KM requires B_KN (wgrad form)");
  BF_REQUIRE(a->split_k >= 1, "bf_gemm: split_k %d", a->split_k);
  BF_REQUIRE(a->split_k == 1 || a->epilogue == BF_EPI_ATOMIC32, "bf_gemm: split_k > 1 needs BF_EPI_ATOMIC32");

  GemmParams p{};
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.split_k = a->split_k;
  p.epilogue = a->epilogue;
  p.is_f16 = a->dtype == BF_F16;
  p.rows_per_group = a->rows_per_group > 0 ? a->rows_per_group : 1;
  p.d2s_h = a->d2s_h; p.d2s_w = a->d2s_w; p.d2s_cout = a->d2s_cout;
  p.bias = a->bias; p.col_scale = a->col_scale; p.col_shift = a->col_shift; p.col_gamma = a->col_gamma;
  p.row_scale = a->row_scale; p.in32 = a->in32; p.aux16 = a->aux16;
  p.out16 = a->out16; p.out16b = a->out16b; p.out32 = a->out32;
  p.ldo = a->ldo; p.ld32 = a->ld32;

  // epilogue operand checks
  switch (a->epilogue) {
    case BF_EPI_STORE16: case BF_EPI_GELU: BF_REQUIRE(a->out16, "bf_gemm: out16 required"); break;
    case BF_EPI_STORE32: case BF_EPI_ATOMIC32: BF_REQUIRE(a->out32, "bf_gemm: out32 required"); break;
    case BF_EPI_RESID:
      BF_REQUIRE(a->out32 && a->in32 && a->col_gamma, "bf_gemm: RESID needs in32/out32/col_gamma");
      BF_REQUIRE((a->col_scale == nullptr) == (a->col_shift == nullptr), "bf_gemm: col_scale/col_shift pair");
      break;
    case BF_EPI_DGELU: BF_REQUIRE(a->out16 && a->aux16, "bf_gemm: DGELU needs out16/aux16"); break;
    case BF_EPI_ACC32: BF_REQUIRE(a->out32 && a->in32, "bf_gemm: ACC32 needs in32/out32"); break;
    case BF_EPI_D2S:
      BF_REQUIRE(a->out16 && a->d2s_h > 0 && a->d2s_w > 0 && a->d2s_cout > 0, "bf_gemm: D2S geometry");
      BF_REQUIRE(a->N == 4 * a->d2s_cout, "bf_gemm: D2S needs N == 4*cout");
      BF_REQUIRE(a->M % (a->d2s_h * a->d2s_w) == 0, "bf_gemm: D2S M not a multiple of h*w");
      break;
    default: BF_REQUIRE(false, "bf_gemm: unknown epilogue %d", a->epilogue);
  }
  if (a->out16 || a->out16b || a->aux16) BF_REQUIRE(a->ldo >= a->N || a->epilogue == BF_EPI_D2S, "bf_gemm: ldo < N");
  if (a->out32 || a->in32) BF_REQUIRE(a->ld32 >= a->N, "bf_gemm: ld32 < N");

  const int bn = pick_bn(*a);
  BF_REQUIRE(!b_mn || bn % 64 == 0, "bf_gemm: internal bn");
  p.num_m_blocks = (a->M + BM - 1) / BM;
  p.num_n_blocks = (a->N + bn - 1) / bn;
  p.idesc = make_idesc_f16(a->dtype == BF_F16 ? 0 : 1, a_mn ? 1 : 0, b_mn ? 1 : 0, BM, bn);

  CUtensorMap ma, mb;
  int st;
  if (s2d) {
    const int C = a->s2d_cin, Hin = a->s2d_hin, Win = a->s2d_win, I = a->s2d_images;
    BF_REQUIRE(C > 0 && Hin > 0 && Win > 0 && I > 0 && Hin % 2 == 0 && Win % 2 == 0, "bf_gemm: S2D geometry");
    BF_REQUIRE(a->K == 4 * C, "bf_gemm: S2D needs K == 4*cin");
    BF_REQUIRE(!b_mn, "bf_gemm: S2D needs B_NK");
    BF_REQUIRE(C % 4 == 0, "bf_gemm: S2D needs cin %% 4 == 0 (16-byte TMA strides)");
    const int Ho = Hin / 2, Wo = Win / 2;
    BF_REQUIRE(a->M == I * Ho * Wo, "bf_gemm: S2D M != images*ho*wo");
    int box_w = Wo >= BM ? BM : Wo;
    BF_REQUIRE((Wo >= BM) ? (Wo % BM == 0) : (BM % Wo == 0),
               "bf_gemm: S2D fast path needs wo to divide or be a multiple of 128 (wo=%d)", Wo);
    const int seg = 2 * C;                       // elements per ky segment: (kx, ci)
    p.s2d = 1; p.s2d_box_w = box_w; p.s2d_wo = Wo;
    p.k_seg_iters = (seg + BK - 1) / BK;
    p.k_iters = 2 * p.k_seg_iters;
    // A: (k = 2C, xo = Wo, ky = 2, row = I*Ho); element strides: 1, 2C, Win*C, 2*Win*C
    uint64_t dims[4] = {(uint64_t)seg, (uint64_t)Wo, 2, (uint64_t)I * Ho};
    uint64_t str[3] = {(uint64_t)seg * 2, (uint64_t)Win * C * 2, (uint64_t)2 * Win * C * 2};
    uint32_t box[4] = {BK, (uint32_t)box_w, 1, (uint32_t)(BM / box_w)};
    if ((st = make_map(&ma, a->dtype, a->A, 4, dims, str, box))) return st;
    // B: weight (N, ky, 2C) -> (k = 2C, ky = 2, n = N)
    uint64_t bd[3] = {(uint64_t)seg, 2, (uint64_t)a->N};
    uint64_t bs[2] = {(uint64_t)seg * 2, (uint64_t)a->ldb * 2};
    uint32_t bb[3] = {BK, 1, (uint32_t)bn};
    BF_REQUIRE(a->ldb >= 4 * C && a->ldb % 8 == 0, "bf_gemm: S2D ldb");
    if ((st = make_map(&mb, a->dtype, a->B, 3, bd, bs, bb))) return st;
  } else {
    BF_REQUIRE(a->lda % 8 == 0 && a->ldb % 8 == 0, "bf_gemm: leading dimensions must be multiples of 8 elements");
    const int k_total = (a->K + BK - 1) / BK;
    BF_REQUIRE(k_total % a->split_k == 0, "bf_gemm: ceil(K/64)=%d not divisible by split_k=%d", k_total, a->split_k);
    p.k_iters = k_total / a->split_k;
    p.k_seg_iters = k_total;
    if (a_mn) {   // (K, M) row-major: inner dim M
      BF_REQUIRE(a->lda >= a->M, "bf_gemm: lda < M");
      uint64_t dims[2] = {(uint64_t)a->M, (uint64_t)a->K};
      uint64_t str[1] = {(uint64_t)a->lda * 2};
      uint32_t box[2] = {64, BK};
      if ((st = make_map(&ma, a->dtype, a->A, 2, dims, str, box))) return st;
    } else {
      BF_REQUIRE(a->lda >= a->K, "bf_gemm: lda < K");
      uint64_t dims[2] = {(uint64_t)a->K, (uint64_t)a->M};
      uint64_t str[1] = {(uint64_t)a->lda * 2};
      uint32_t box[2] = {BK, BM};
      if ((st = make_map(&ma, a->dtype, a->A, 2, dims, str, box))) return st;
    }
    if (b_mn) {   // (K, N) row-major: inner dim N
      BF_REQUIRE(a->ldb >= a->N, "bf_gemm: ldb < N");
      uint64_t dims[2] = {(uint64_t)a->N, (uint64_t)a->K};
      uint64_t str[1] = {(uint64_t)a->ldb * 2};
      uint32_t box[2] = {64, BK};
      if ((st = make_map(&mb, a->dtype, a->B, 2, dims, str, box))) return st;
    } else {
      BF_REQUIRE(a->ldb >= a->K, "bf_gemm: ldb < K");
      uint64_t dims[2] = {(uint64_t)a->K, (uint64_t)a->N};
      uint64_t str[1] = {(uint64_t)a->ldb * 2};
      uint32_t box[2] = {BK, (uint32_t)bn};
      if ((st = make_map(&mb, a->dtype, a->B, 2, dims, str, box))) return st;
    }
  }

  cudaStream_t s = static_cast<cudaStream_t>(stream);
#define BF_DISPATCH(BN_)                                                          \
  if (bn == BN_) {                                                                \
    if (a_mn) return launch<BN_, true, true>(ma, mb, p, s);                       \
    if (b_mn) return launch<BN_, false, true>(ma, mb, p, s);                      \
    return launch<BN_, false, false>(ma, mb, p, s);                               \
  }
  BF_DISPATCH(64)
  BF_DISPATCH(128)
  BF_DISPATCH(192)
  BF_DISPATCH(256)
#undef BF_DISPATCH
  set_error("bf_gemm: no kernel for BN=%d", bn);
  return BF_ERR_INVALID;
}
