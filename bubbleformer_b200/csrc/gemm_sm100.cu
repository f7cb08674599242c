// tcgen05 / TMEM / TMA GEMM for sm_100a with fused epilogues -- the contraction engine of the
// FiLMAViT hot path (see include/bubbleformer_b200.h, bf_gemm).
//
// Structure (one CTA per SM, persistent over output tiles, 128 x BN tile, BK = 64):
//   warp 0      : TMA producer of the A / B operand ring (cp.async.bulk.tensor, 128B swizzle, mbarrier tx)
//   warp 1      : MMA issuer (tcgen05.mma cta_group::1 kind::f16, fp32 accumulators in TMEM;
//                 tcgen05.commit releases smem slots / publishes accumulators)
//   warp 2      : TMA producer of the epilogue *input* tile (fp32 residual stream / saved pre-activation),
//                 prefetched while the tile's MMAs run
//   warp 3      : idle (keeps the epilogue warps aligned to the TMEM lane quadrants)
//   warps 4..   : epilogue (8 warps for tiles up to 128 columns, 12 for 192, 16 for 256: one warp per TMEM lane
//                 quadrant and 64-column group): tcgen05.ld -> registers -> fused math -> swizzled smem -> TMA store
//                 (cp.async.bulk.tensor shared -> global, or cp.reduce.async.bulk.tensor .add for split-K)
// TMEM holds two accumulator stages so the epilogue of tile i overlaps the MMAs of tile i+1.  No epilogue
// warp ever touches global memory with ld/st: inputs arrive by TMA into smem, outputs leave by TMA from
// per-warp 32 x 32 slabs, so every DRAM transaction is a full, coalesced line and nothing waits on it.
//
// Operand layouts: K-major (row-major (rows, K)) or MN-major (row-major (K, rows)) for both A and B,
// so forward, dgrad (B = W read K x N) and wgrad (A = dY^T, B = X, contraction over tokens, split-K)
// all run through this kernel without any transposed copy in HBM.
// A can also be an implicit 2x2/stride-2 patch gather expressed purely as a 4-D TMA tensor map.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "common.cuh"

namespace bf {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kFirstEpiWarp = 4;
constexpr int kMaxStages = 8;
// Epilogue warps: the tile period of the short-K GEMMs is the time ONE warp needs for its share of a tile's epilogue
// (a latency chain of tcgen05.ld, math, shared-memory staging and the TMA store), so wide tiles get one warp per
// (TMEM lane quadrant, 64-column group): 12 warps for 192 columns, 16 for 256; 8 for tiles up to 128 columns.
__host__ __device__ constexpr int epi_warps(int bn) { return bn <= 128 ? 8 : 4 * (bn / 64); }
__host__ __device__ constexpr int gemm_threads(int bn) { return 32 * (kFirstEpiWarp + epi_warps(bn)); }
// per-warp output staging: 8 KiB (double buffered) with 8 epilogue warps, 4 KiB (single) with more
__host__ __device__ constexpr int slab_bytes_per_warp(int bn) { return bn <= 128 ? 8192 : 4096; }
constexpr int kSmemBudget = 227 * 1024;

struct GemmParams {
  int M, N, K;
  int num_m_blocks, num_n_blocks, split_k;
  int k_iters;        // k iterations (of BK) per output tile (per split)
  int k_seg_iters;    // S2D: iterations per ky segment; otherwise == total iterations
  int s2d;            // 1: A coords are 4-D (k, xo, ky, row), B coords 3-D (k, ky, n)
  int s2d_box_w;      // xo extent of the A box (divides 128)
  int s2d_wo;         // output width
  int b_s2d;          // 1: B (MN-major, K = output pixels) is the implicit patch gather: 4-D coords (n in segment, xo, ky, row)
  int b_seg;          // B_S2D: elements per ky segment (2 * cin)
  uint32_t idesc;
  int is_f16;
  int gelu_exact;     // GELU / DGELU epilogues: 1 = exact erf (bf_set_gelu_mode), 0 = tanh form
  int gelu_h2;        // GELU_D, tanh form: evaluate in packed half precision (BF_GELU_H2=0 keeps fp32: A/B measurements)
  int epilogue;
  int rows_per_group;
  int d2s_h, d2s_w, d2s_cout;
  // shared-memory plan (bytes from the 1 KiB aligned base)
  int stages;
  int in_kind;        // 0: no epilogue input tile, 1: 16-bit (aux16), 2: fp32 (in32)
  int in_bufs;        // 1 or 2
  int in_bytes;       // bytes of one input tile buffer
  int in_off, slab_off, bar_off;
  int cg;             // CTAs per tile: 1, or 2 = CTA pairs (cta_group::2)
  int b_resident;     // 1: the CTA's whole B block (k_iters boxes) stays in shared memory for all of its tiles
  int a_off;          // start of the operand ring (after the resident B block)
  int ln_heads;       // QKV_LN: heads = N / 192
  int vec_off;        // per-column vectors staged in shared memory: [bias | col_scale | col_shift | col_gamma][vec_cols]
  int vec_cols;       // columns staged per vector (a multiple of 32): BN when B-resident, else N rounded up
  const float* bias;
  const float* col_scale;
  const float* col_shift;
  const float* col_gamma;
  const float* row_scale;
  void* out16;        // D2S only (direct stores)
  int has_out16, has_out16b;
  float* stats_out;   // RESID: stats_out[(m / rows_per_group)][n][2] += (sum, sum^2) of out32 (may be null)
  float* ln_rstd;     // QKV_LN: (M, heads, 2) rstd of the raw q / k rows
  float* colsum_out;  // DGELU: [N] += column sums of the stored output (accumulated per CTA in shared memory)
  long ldo;
};

// ---------------------------------------------------------------------------------------------
// swizzled shared-memory rows.  `base` is the (1 KiB aligned) start of a TMA box whose rows are
// 128 B (fp32, SWIZZLE_128B) or 64 B (16-bit, SWIZZLE_64B) wide; `row` is the row index inside the box.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void st_row_f32(uint8_t* base, int row, const float (&v)[32]) {
  uint8_t* r = base + row * 128;
  const int x = row & 7;
#pragma unroll
  for (int c = 0; c < 8; ++c)
    *reinterpret_cast<float4*>(r + ((c ^ x) << 4)) = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
}
__device__ __forceinline__ void ld_row_f32(const uint8_t* base, int row, float (&v)[32]) {
  const uint8_t* r = base + row * 128;
  const int x = row & 7;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const float4 u = *reinterpret_cast<const float4*>(r + ((c ^ x) << 4));
    v[4 * c] = u.x; v[4 * c + 1] = u.y; v[4 * c + 2] = u.z; v[4 * c + 3] = u.w;
  }
}
template <typename T16>
__device__ __forceinline__ void st_row_16(uint8_t* base, int row, const float (&v)[32]) {
  uint8_t* r = base + row * 64;
  const int x = (row >> 1) & 3;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint4 u;
    u.x = pack2<T16>(v[8 * c], v[8 * c + 1]); u.y = pack2<T16>(v[8 * c + 2], v[8 * c + 3]);
    u.z = pack2<T16>(v[8 * c + 4], v[8 * c + 5]); u.w = pack2<T16>(v[8 * c + 6], v[8 * c + 7]);
    *reinterpret_cast<uint4*>(r + ((c ^ x) << 4)) = u;
  }
}
// same, from 16 already packed pairs
__device__ __forceinline__ void st_row_16_packed(uint8_t* base, int row, const uint32_t (&u)[16]) {
  uint8_t* r = base + row * 64;
  const int x = (row >> 1) & 3;
#pragma unroll
  for (int c = 0; c < 4; ++c)
    *reinterpret_cast<uint4*>(r + ((c ^ x) << 4)) = make_uint4(u[4 * c], u[4 * c + 1], u[4 * c + 2], u[4 * c + 3]);
}
template <typename T16>
__device__ __forceinline__ void ld_row_16(const uint8_t* base, int row, float (&v)[32]) {
  const uint8_t* r = base + row * 64;
  const int x = (row >> 1) & 3;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const uint4 u = *reinterpret_cast<const uint4*>(r + ((c ^ x) << 4));
    const float2 a = unpack2<T16>(u.x), b = unpack2<T16>(u.y), d = unpack2<T16>(u.z), e = unpack2<T16>(u.w);
    v[8 * c] = a.x; v[8 * c + 1] = a.y; v[8 * c + 2] = b.x; v[8 * c + 3] = b.y;
    v[8 * c + 4] = d.x; v[8 * c + 5] = d.y; v[8 * c + 6] = e.x; v[8 * c + 7] = e.y;
  }
}

// per-column fp32 vector (bias, scales) staged in shared memory at kernel start: 32 consecutive entries from a
// 128-byte aligned, warp-uniform address (broadcast reads; no global-memory latency inside the tile loop)
__device__ __forceinline__ void load_cols(const float* svec, float (&o)[32]) {
  const float4* p4 = reinterpret_cast<const float4*>(svec);
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 u = p4[q];
    o[4 * q] = u.x; o[4 * q + 1] = u.y; o[4 * q + 2] = u.z; o[4 * q + 3] = u.w;
  }
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
// CG = 2: a CTA PAIR (2-CTA cluster, tcgen05 cta_group::2) owns a 256 x BN tile: each CTA stages its own 128 rows of A
// and HALF of the B tile (BN / 2 rows), the leader CTA issues 256-row MMAs that read both shared memories, and each CTA
// runs the epilogue of its own 128 accumulator rows.  Per CTA and k step that is 16 KB + BN/2 * 128 B of operands instead
// of 16 KB + BN * 128 B: the L2 -> SM operand traffic of a 128 x 192 tile (77 FLOP per L2 byte, ~955 TFLOP/s at the
// measured ~12.4 TB/s L2 ceiling -- where the single-CTA kernel saturates) drops by 30 % (33 % for BN = 256).
template <int BN, bool A_MN, bool B_MN, int CG>
__global__ void __launch_bounds__(gemm_threads(BN), 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                    const __grid_constant__ CUtensorMap map_in, const __grid_constant__ CUtensorMap map_o16,
                    const __grid_constant__ CUtensorMap map_o16b, const __grid_constant__ CUtensorMap map_o32,
                    const GemmParams p) {
  static_assert(CG == 1 || CG == 2, "cta_group");
  static_assert(CG == 1 || !A_MN, "the pair kernel takes K-major A");
  static_assert(CG == 1 || !B_MN || (BN / CG) % 64 == 0, "MN-major B halves must be whole 64-column blocks");
  constexpr int kABytes = BM * BK * 2;
  constexpr int kBRows = BN / CG;                  // B rows staged by this CTA
  constexpr int kBBytes = kBRows * BK * 2;
  constexpr int kStageBytes = kABytes + kBBytes;
  constexpr int kTmemCols = (2 * BN <= 128) ? 128 : (2 * BN <= 256 ? 256 : 512);
  constexpr int kChunks = BN / 32;                 // 32-column chunks per tile
  constexpr int kEpiWarps = epi_warps(BN);
  constexpr int kThreads = gemm_threads(BN);
  constexpr int kChunksPerWarp = kChunks / (kEpiWarps / 4);   // kEpiWarps / 4 warps share one TMEM lane quadrant
  constexpr int kSlabBytes = slab_bytes_per_warp(BN);
  constexpr bool kDouble = kSlabBytes >= 8192;     // room for two sets of staging buffers per warp

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.bar_off);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tmem_full = empty_bar + kMaxStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* in_full = tmem_empty + 2;
  uint64_t* in_empty = in_full + 2;
  uint64_t* b_full = in_empty + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(b_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int cta_rank = CG == 2 ? (int)cluster_ctarank() : 0;
  const int first_tile = (int)blockIdx.x / CG;     // the CTAs of a pair walk the same tile sequence
  const int tile_stride = (int)gridDim.x / CG;

  float* s_colsum = reinterpret_cast<float*>(smem + p.slab_off);       // DGELU + colsum_out only ([num_n_blocks * BN])
  if (p.colsum_out != nullptr)
    for (int i = threadIdx.x; i < p.num_n_blocks * BN; i += kThreads) s_colsum[i] = 0.f;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar + s, 1);
      mbar_init(empty_bar + s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tmem_full + s, 1);
      mbar_init(tmem_empty + s, CG * kEpiWarps);
      mbar_init(in_full + s, 1);
      mbar_init(in_empty + s, kEpiWarps);
    }
    mbar_init(b_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    if constexpr (CG == 2) tmem_alloc_pair(tmem_ptr, kTmemCols); else tmem_alloc(tmem_ptr, kTmemCols);
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync_all();       // the peer's barriers are initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_prologue_done();
  // per-column vectors -> shared memory (zero padded); a B-resident CTA only ever needs its own BN columns
  const int vec_base = p.b_resident ? ((int)blockIdx.x % p.num_n_blocks) * BN : 0;
  float* s_vec = reinterpret_cast<float*>(smem + p.vec_off);
  {
    const float* const vsrc[4] = {p.bias, p.col_scale, p.col_shift, p.col_gamma};
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      if (vsrc[v] != nullptr) {
        for (int i = threadIdx.x; i < p.vec_cols; i += kThreads)
          s_vec[v * p.vec_cols + i] = (vec_base + i < p.N) ? __ldg(vsrc[v] + vec_base + i) : 0.f;
      }
    }
  }
  __syncthreads();
  const float* s_bias = s_vec - vec_base;
  const float* s_cscale = s_bias + p.vec_cols;
  const float* s_cshift = s_cscale + p.vec_cols;
  const float* s_cgamma = s_cshift + p.vec_cols;

  const int tiles_mn = p.num_m_blocks * p.num_n_blocks;
  const int num_tiles = tiles_mn * p.split_k;

  if (warp == 0) {
    // ===================== TMA producer: operands =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const bool res = p.b_resident != 0;
      if (CG == 1 && res && (int)blockIdx.x < num_tiles) {
        // B-resident schedule: the grid is a multiple of num_n_blocks, so this CTA's n block never changes; its
        // (BN x K) operand block is loaded once and every tile only streams its A rows through the ring
        const int n0 = ((int)blockIdx.x % p.num_n_blocks) * BN;
        mbar_arrive_expect_tx(b_full, p.k_iters * kBBytes);
        for (int it = 0; it < p.k_iters; ++it) {
          uint8_t* sb = smem + it * kBBytes;
          if (B_MN) {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) tma_load_2d(sb + j * (64 * BK * 2), &map_b, b_full, n0 + 64 * j, it * BK);
          } else {
            tma_load_2d(sb, &map_b, b_full, it * BK, n0);
          }
        }
      }
      for (int tile = first_tile; tile < num_tiles; tile += tile_stride) {
        const int ks = tile / tiles_mn;
        const int mn = tile - ks * tiles_mn;
        const int m_blk = mn / p.num_n_blocks;
        const int n_blk = mn - m_blk * p.num_n_blocks;
        const int m0 = (m_blk * CG + cta_rank) * BM, n0 = n_blk * BN;
        for (int it = 0; it < p.k_iters; ++it) {
          mbar_wait(empty_bar + stage, phase ^ 1u);
          uint8_t* sa = res ? smem + p.a_off + stage * kABytes : smem + stage * kStageBytes;
          uint8_t* sb = sa + kABytes;
          const int kit = ks * p.k_iters + it;      // global k iteration
          if constexpr (CG == 2) {
            // both CTAs' loads complete on the LEADER's barrier, which expects the bytes of the whole pair
            const uint32_t lead_bar = mapa_shared(smem_u32(full_bar + stage), 0);
            if (cta_rank == 0) mbar_arrive_expect_tx(full_bar + stage, 2 * kStageBytes);
            const int k0 = kit * BK;
            const int nb0 = n0 + cta_rank * kBRows;
            tma_load_2d_pair(sa, &map_a, lead_bar, k0, m0);
            if (B_MN) {
#pragma unroll
              for (int j = 0; j < kBRows / 64; ++j)
                tma_load_2d_pair(sb + j * (64 * BK * 2), &map_b, lead_bar, nb0 + 64 * j, k0);
            } else {
              tma_load_2d_pair(sb, &map_b, lead_bar, k0, nb0);
            }
            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
            continue;
          }
          mbar_arrive_expect_tx(full_bar + stage, res ? kABytes : kStageBytes);
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
              tma_load_2d(sa + kABytes / 2, &map_a, full_bar + stage, m0 + 64, k0);
            } else {
              tma_load_2d(sa, &map_a, full_bar + stage, k0, m0);
            }
            if (res) {
              // B is already resident
            } else if (B_MN) {
              if (p.b_s2d) {
                // k0 .. k0+63 are 64 consecutive output pixels of one image row; each 64-column block lies in one ky segment
                const int xo0 = k0 % p.s2d_wo, r0 = k0 / p.s2d_wo;
#pragma unroll
                for (int j = 0; j < BN / 64; ++j) {
                  const int n = n0 + 64 * j;
                  const int ky = n / p.b_seg;
                  tma_load_4d(sb + j * (64 * BK * 2), &map_b, full_bar + stage, n - ky * p.b_seg, xo0, ky, r0);
                }
              } else {
#pragma unroll
                for (int j = 0; j < BN / 64; ++j)
                  tma_load_2d(sb + j * (64 * BK * 2), &map_b, full_bar + stage, n0 + 64 * j, k0);
              }
            } else {
              tma_load_2d(sb, &map_b, full_bar + stage, k0, n0);
            }
          }
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0 && cta_rank == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      const bool res = p.b_resident != 0;
      if (res && (int)blockIdx.x < num_tiles) mbar_wait(b_full, 0);
      for (int tile = first_tile; tile < num_tiles; tile += tile_stride) {
        mbar_wait(tmem_empty + as, aphase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * BN);
        for (int it = 0; it < p.k_iters; ++it) {
          mbar_wait(full_bar + stage, phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(res ? smem + p.a_off + stage * kABytes : smem + stage * kStageBytes);
          const uint32_t sb = res ? smem_u32(smem + it * kBBytes) : sa + kABytes;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // K-major: advance 16 elements = 32 B inside the 128 B swizzle row.
            // MN-major: advance 16 K rows = two 8-row groups of 1024 B.
            const uint64_t da = A_MN ? make_smem_desc(sa + k * 2048, 64 * BK * 2, 1024)
                                     : make_smem_desc(sa + k * 32, 16, 1024);
            const uint64_t db = B_MN ? make_smem_desc(sb + k * 2048, 64 * BK * 2, 1024)
                                     : make_smem_desc(sb + k * 32, 16, 1024);
            if constexpr (CG == 2) umma_f16_pair(d_tmem, da, db, p.idesc, (it | k) != 0 ? 1u : 0u);
            else umma_f16(d_tmem, da, db, p.idesc, (it | k) != 0 ? 1u : 0u);
          }
          if constexpr (CG == 2) umma_commit_pair(empty_bar + stage, 3); else umma_commit(empty_bar + stage);
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
        if constexpr (CG == 2) umma_commit_pair(tmem_full + as, 3); else umma_commit(tmem_full + as);
        if (++as == 2) { as = 0; aphase ^= 1u; }
      }
    }
  } else if (warp == 2) {
    // ===================== TMA producer: epilogue input tile =====================
    if (lane == 0 && p.in_kind != 0) {
      tma_prefetch_desc(&map_in);
      int ib = 0;
      uint32_t iphase = 0;
      const int box_bytes = BM * (p.in_kind == 2 ? 128 : 64);
      for (int tile = first_tile; tile < num_tiles; tile += tile_stride) {
        const int mn = tile % tiles_mn;
        const int m_blk = (mn / p.num_n_blocks) * CG + cta_rank;
        const int n_blk = mn % p.num_n_blocks;
        mbar_wait(in_empty + ib, iphase ^ 1u);
        uint8_t* dst = smem + p.in_off + ib * p.in_bytes;
        mbar_arrive_expect_tx(in_full + ib, kChunks * box_bytes);
#pragma unroll 1
        for (int c = 0; c < kChunks; ++c)
          tma_load_2d(dst + c * box_bytes, &map_in, in_full + ib, n_blk * BN + 32 * c, m_blk * BM);
        if (++ib == p.in_bufs) { ib = 0; iphase ^= 1u; }
      }
    }
  } else if (warp >= kFirstEpiWarp) {
    // ===================== epilogue warps =====================
    const int ew = warp - kFirstEpiWarp;
    const int quad = warp & 3;              // TMEM lane quadrant this warp may access
    const int half = ew >> 2;               // which group of the tile's 32-column chunks (0 .. kEpiWarps/4 - 1)
    const int row = quad * 32 + lane;       // row inside the tile
    uint8_t* slab = smem + p.slab_off + ew * kSlabBytes;
    if (lane == 0) {
      if (p.has_out16) tma_prefetch_desc(&map_o16);
      if (p.has_out16b) tma_prefetch_desc(&map_o16b);
      if (p.epilogue != BF_EPI_STORE16 && p.epilogue != BF_EPI_GELU && p.epilogue != BF_EPI_DGELU &&
          p.epilogue != BF_EPI_GELU_D && p.epilogue != BF_EPI_DMUL &&
          p.epilogue != BF_EPI_D2S && p.epilogue != BF_EPI_QKV_LN)
        tma_prefetch_desc(&map_o32);
    }
    int as = 0, ib = 0, sb = 0;             // accumulator stage, input buffer, slab double-buffer index
    uint32_t aphase = 0, iphase = 0;
    // hand an accumulator stage back to the (leader's) MMA warp
    const uint32_t lead_tmem_empty = CG == 2 ? mapa_shared(smem_u32(tmem_empty), 0) : 0u;
    auto release_tmem = [&](int stage_) {
      if constexpr (CG == 2) mbar_arrive_cluster(lead_tmem_empty + 8u * (uint32_t)stage_);
      else mbar_arrive(tmem_empty + stage_);
    };
    for (int tile = first_tile; tile < num_tiles; tile += tile_stride) {
      const int mn = tile % tiles_mn;
      const int m_blk = (mn / p.num_n_blocks) * CG + cta_rank;
      const int n_blk = mn % p.num_n_blocks;
      const int m0 = m_blk * BM, n0 = n_blk * BN;
      const int m = m0 + row;
      mbar_wait(tmem_full + as, aphase);
      tc_fence_after();
      uint8_t* in_tile = nullptr;
      if (p.in_kind != 0) {
        mbar_wait(in_full + ib, iphase);
        in_tile = smem + p.in_off + ib * p.in_bytes;
      }
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(as * BN);
      float rs = 1.f;
      if (p.epilogue == BF_EPI_RESID && p.row_scale != nullptr && m < p.M) rs = __ldg(p.row_scale + m / p.rows_per_group);
      if constexpr (BN == 192) {
        if (p.epilogue == BF_EPI_QKV_LN) {
          // One tile = one head: columns [q 0:64 | k 64:128 | v 128:192], one epilogue warp per (lane quadrant, part):
          // `half` = 0 normalises q, 1 normalises k, 2 copies v.  LayerNorm over the 64 columns of a row is thread
          // local (tcgen05.ld hands each thread its row).  Stored: xhat = (x - mean) * rstd without the affine part,
          // plus rstd for the backward.
          const int mrow = m0 + quad * 32;
          const bool live = mrow < p.M;                       // warp uniform
          const int ncol = n0 + 64 * half;
          float x0[32], x1[32];
          tmem_ld_32x32(t_row + static_cast<uint32_t>(64 * half), x0);
          tmem_ld_32x32(t_row + static_cast<uint32_t>(64 * half + 32), x1);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) release_tmem(as);
          if (live) {
            if (p.bias != nullptr) {
              float b[32];
              load_cols(s_bias + ncol, b);
#pragma unroll
              for (int j = 0; j < 32; ++j) x0[j] += b[j];
              load_cols(s_bias + ncol + 32, b);
#pragma unroll
              for (int j = 0; j < 32; ++j) x1[j] += b[j];
            }
            if (half < 2) {
              float sm_ = 0.f;
#pragma unroll
              for (int j = 0; j < 32; ++j) sm_ += x0[j] + x1[j];
              const float mean = sm_ * (1.f / 64.f);
              float q = 0.f;
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                x0[j] -= mean; x1[j] -= mean;
                q = fmaf(x0[j], x0[j], q); q = fmaf(x1[j], x1[j], q);
              }
              const float rstd = rsqrtf(q * (1.f / 64.f) + 1e-5f);
#pragma unroll
              for (int j = 0; j < 32; ++j) { x0[j] *= rstd; x1[j] *= rstd; }
              if (m < p.M) p.ln_rstd[((long)m * p.ln_heads + n_blk) * 2 + half] = rstd;
            }
            if (lane == 0) tma_store_wait_read<0>();          // the previous tile's boxes have left the slab
            __syncwarp();
            if (p.is_f16) { st_row_16<__half>(slab, lane, x0); st_row_16<__half>(slab + 2048, lane, x1); }
            else { st_row_16<__nv_bfloat16>(slab, lane, x0); st_row_16<__nv_bfloat16>(slab + 2048, lane, x1); }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&map_o16, slab, ncol, mrow);
              tma_store_2d(&map_o16, slab + 2048, ncol + 32, mrow);
              tma_store_commit();
            }
          }
          if (++as == 2) { as = 0; aphase ^= 1u; }
          continue;
        }
      }
      if constexpr (BN == 128) {
        if (p.epilogue == BF_EPI_QKV_LN) {
          // Columns are ordered (head, q | k | v, 64): a tile is two 64-column groups, group gi = n0 / 64 + half belongs
          // to head gi / 3 and is its q (0), k (1) or v (2) part.  The warp pair of a lane quadrant splits the groups;
          // LayerNorm over the 64 columns of a row is thread local (tcgen05.ld hands each thread its row).  Stored:
          // xhat = (x - mean) * rstd without the affine part (v unchanged), plus rstd of q / k for the backward.
          const int mrow = m0 + quad * 32;
          const int gi = (n0 >> 6) + half;
          const int part = gi % 3;
          const int ncol = n0 + 64 * half;
          const bool live = mrow < p.M && ncol < p.N;           // warp uniform
          uint8_t* s = slab + sb * 4096;
          if (lane == 0) tma_store_wait_read<1>();
          __syncwarp();
          float x0[32], x1[32];
          tmem_ld_32x32(t_row + static_cast<uint32_t>(64 * half), x0);
          tmem_ld_32x32(t_row + static_cast<uint32_t>(64 * half + 32), x1);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) release_tmem(as);
          if (live) {
            if (p.bias != nullptr) {
              float b[32];
              load_cols(s_bias + ncol, b);
#pragma unroll
              for (int j = 0; j < 32; ++j) x0[j] += b[j];
              load_cols(s_bias + ncol + 32, b);
#pragma unroll
              for (int j = 0; j < 32; ++j) x1[j] += b[j];
            }
            if (part < 2) {
              float sm_ = 0.f;
#pragma unroll
              for (int j = 0; j < 32; ++j) sm_ += x0[j] + x1[j];
              const float mean = sm_ * (1.f / 64.f);
              float q = 0.f;
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                x0[j] -= mean; x1[j] -= mean;
                q = fmaf(x0[j], x0[j], q); q = fmaf(x1[j], x1[j], q);
              }
              const float rstd = rsqrtf(q * (1.f / 64.f) + 1e-5f);
#pragma unroll
              for (int j = 0; j < 32; ++j) { x0[j] *= rstd; x1[j] *= rstd; }
              if (m < p.M) p.ln_rstd[((long)m * p.ln_heads + gi / 3) * 2 + part] = rstd;
            }
            if (p.is_f16) { st_row_16<__half>(s, lane, x0); st_row_16<__half>(s + 2048, lane, x1); }
            else { st_row_16<__nv_bfloat16>(s, lane, x0); st_row_16<__nv_bfloat16>(s + 2048, lane, x1); }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&map_o16, s, ncol, mrow);
              tma_store_2d(&map_o16, s + 2048, ncol + 32, mrow);
              tma_store_commit();
            }
            sb ^= 1;
          }
          if (++as == 2) { as = 0; aphase ^= 1u; }
          continue;
        }
      }
#pragma unroll 1
      for (int ci = 0; ci < kChunksPerWarp; ++ci) {
        const int c = half * kChunksPerWarp + ci;
        const int n = n0 + 32 * c;
        float acc[32];
        tmem_ld_32x32(t_row + static_cast<uint32_t>(32 * c), acc);
        tmem_ld_wait();
        if (ci == kChunksPerWarp - 1) {
          // accumulator fully read: hand the TMEM stage back to the MMA warp before finishing the stores
          tc_fence_before();
          __syncwarp();
          if (lane == 0) release_tmem(as);
        }
        if (n >= p.N || m0 + quad * 32 >= p.M) continue;     // chunk entirely outside the matrix (warp uniform)
        if (p.bias != nullptr) {
          float b[32];
          load_cols(s_bias + n, b);
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[j] += b[j];
        }
        const int mrow = m0 + quad * 32;                      // first global row of this warp's slab
        switch (p.epilogue) {
          case BF_EPI_STORE16: {
            uint8_t* s = slab + sb * 2048;
            if (lane == 0) tma_store_wait_read<1>();
            __syncwarp();
            if (p.is_f16) st_row_16<__half>(s, lane, acc); else st_row_16<__nv_bfloat16>(s, lane, acc);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) { tma_store_2d(&map_o16, s, n, mrow); tma_store_commit(); }
            if (p.stats_out != nullptr) {
              // per-(image, channel) sums of the stored (rounded) values for the InstanceNorm that follows: lane j
              // owns column j of this warp's 32 x 32 block (conflict-free transposed read of the swizzled rows)
              const int rows_valid = min(32, p.M - mrow);
              float s1 = 0.f, s2q = 0.f;
#pragma unroll 8
              for (int r = 0; r < 32; ++r) {
                const uint16_t h = *reinterpret_cast<const uint16_t*>(s + r * 64 + ((((lane >> 3) ^ ((r >> 1) & 3)) << 4) | ((lane & 7) << 1)));
                float v;
                if (p.is_f16) v = __half2float(*reinterpret_cast<const __half*>(&h));
                else v = __uint_as_float(static_cast<uint32_t>(h) << 16);
                if (r < rows_valid) { s1 += v; s2q = fmaf(v, v, s2q); }
              }
              if (n + lane < p.N) {
                float* dst = p.stats_out + ((long)(mrow / p.rows_per_group) * p.N + n + lane) * 2;
                atomicAdd(dst, s1);
                atomicAdd(dst + 1, s2q);
              }
            }
            sb ^= 1;
            break;
          }
          case BF_EPI_GELU:
          case BF_EPI_GELU_D: {
            uint8_t* s = slab + (kDouble ? sb * 2048 : 0);
            uint8_t* s2 = slab + (kDouble ? 4096 + sb * 2048 : 2048);
            if (lane == 0) { if (kDouble) tma_store_wait_read<1>(); else tma_store_wait_read<0>(); }
            __syncwarp();
            if (p.epilogue == BF_EPI_GELU_D && p.gelu_exact == 0 && p.gelu_h2) {
              // value and derivative in packed half precision (two elements per instruction), stored straight away
              uint32_t ug[16], ud[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                __half2 g2, d2;
                gelu_both_h2(acc[2 * j], acc[2 * j + 1], g2, d2);
                if (p.is_f16) { ug[j] = pack_h2<__half>(g2); ud[j] = pack_h2<__half>(d2); }
                else { ug[j] = pack_h2<__nv_bfloat16>(g2); ud[j] = pack_h2<__nv_bfloat16>(d2); }
              }
              if (p.has_out16b) st_row_16_packed(s2, lane, ud);
              st_row_16_packed(s, lane, ug);
              fence_proxy_async();
              __syncwarp();
              if (lane == 0) {
                tma_store_2d(&map_o16, s, n, mrow);
                if (p.has_out16b) tma_store_2d(&map_o16b, s2, n, mrow);
                tma_store_commit();
              }
              sb ^= 1;
              break;
            }
            if (p.epilogue == BF_EPI_GELU_D) {
              // second output = gelu'(pre): the backward then needs no transcendental (and no second tanh at all)
              float d[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) gelu_both(acc[j], p.gelu_exact, acc[j], d[j]);
              if (p.has_out16b) {
                if (p.is_f16) st_row_16<__half>(s2, lane, d); else st_row_16<__nv_bfloat16>(s2, lane, d);
              }
            } else {
              if (p.has_out16b) {
                if (p.is_f16) st_row_16<__half>(s2, lane, acc); else st_row_16<__nv_bfloat16>(s2, lane, acc);
              }
#pragma unroll
              for (int j = 0; j < 32; ++j) acc[j] = gelu_fwd(acc[j], p.gelu_exact);
            }
            if (p.is_f16) st_row_16<__half>(s, lane, acc); else st_row_16<__nv_bfloat16>(s, lane, acc);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&map_o16, s, n, mrow);
              if (p.has_out16b) tma_store_2d(&map_o16b, s2, n, mrow);
              tma_store_commit();
            }
            sb ^= 1;
            break;
          }
          case BF_EPI_STORE32:
          case BF_EPI_ATOMIC32: {
            uint8_t* s = slab + (kDouble ? sb * 4096 : 0);
            if (lane == 0) { if (kDouble) tma_store_wait_read<1>(); else tma_store_wait_read<0>(); }
            __syncwarp();
            st_row_f32(s, lane, acc);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              if (p.epilogue == BF_EPI_ATOMIC32) tma_reduce_add_2d(&map_o32, s, n, mrow);
              else tma_store_2d(&map_o32, s, n, mrow);
              tma_store_commit();
            }
            sb ^= 1;
            break;
          }
          case BF_EPI_DGELU:
          case BF_EPI_DMUL: {
            uint8_t* box = in_tile + c * (BM * 64);
            float pre[32];
            if (p.is_f16) ld_row_16<__half>(box, row, pre); else ld_row_16<__nv_bfloat16>(box, row, pre);
            if (p.epilogue == BF_EPI_DMUL) {
#pragma unroll
              for (int j = 0; j < 32; ++j) acc[j] *= pre[j];
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) acc[j] *= gelu_bwd(pre[j], p.gelu_exact);
            }
            if (p.is_f16) st_row_16<__half>(box, row, acc); else st_row_16<__nv_bfloat16>(box, row, acc);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) { tma_store_2d(&map_o16, box + quad * 32 * 64, n, mrow); tma_store_commit(); }
            if (p.colsum_out != nullptr) {
              // lane j sums column j of this warp's 32 x 32 block as stored (conflict-free read of the swizzled rows)
              const uint8_t* sl = box + quad * 32 * 64;
              const int rows_valid = min(32, p.M - mrow);
              float cs = 0.f;
#pragma unroll 8
              for (int r = 0; r < 32; ++r) {
                const uint16_t h = *reinterpret_cast<const uint16_t*>(sl + r * 64 + ((((lane >> 3) ^ ((r >> 1) & 3)) << 4) | ((lane & 7) << 1)));
                float v;
                if (p.is_f16) v = __half2float(*reinterpret_cast<const __half*>(&h));
                else v = __uint_as_float(static_cast<uint32_t>(h) << 16);
                if (r < rows_valid) cs += v;
              }
              atomicAdd(s_colsum + n + lane, cs);
            }
            break;
          }
          case BF_EPI_ACC32: {
            uint8_t* box = in_tile + c * (BM * 128);
            float g[32];
            ld_row_f32(box, row, g);
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[j] += g[j];
            st_row_f32(box, row, acc);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) { tma_store_2d(&map_o32, box + quad * 32 * 128, n, mrow); tma_store_commit(); }
            break;
          }
          case BF_EPI_RESID: {
            uint8_t* box = in_tile + c * (BM * 128);
            uint8_t* s = slab + (kDouble ? sb * 2048 : 0);
            uint8_t* s2 = slab + (kDouble ? 4096 + sb * 2048 : 2048);
            if (lane == 0) { if (kDouble) tma_store_wait_read<1>(); else tma_store_wait_read<0>(); }
            __syncwarp();
            if (p.has_out16b) {
              if (p.is_f16) st_row_16<__half>(s2, lane, acc); else st_row_16<__nv_bfloat16>(s2, lane, acc);
            }
            float t[32];
            if (p.col_scale != nullptr) {
              load_cols(s_cscale + n, t);
#pragma unroll
              for (int j = 0; j < 32; ++j) acc[j] *= t[j];
              load_cols(s_cshift + n, t);
#pragma unroll
              for (int j = 0; j < 32; ++j) acc[j] += t[j];
            }
            load_cols(s_cgamma + n, t);
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[j] *= rs * t[j];
            ld_row_f32(box, row, t);
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[j] += t[j];
            st_row_f32(box, row, acc);
            if (p.has_out16) {
              if (p.is_f16) st_row_16<__half>(s, lane, acc); else st_row_16<__nv_bfloat16>(s, lane, acc);
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&map_o32, box + quad * 32 * 128, n, mrow);
              if (p.has_out16) tma_store_2d(&map_o16, s, n, mrow);
              if (p.has_out16b) tma_store_2d(&map_o16b, s2, n, mrow);
              tma_store_commit();
            }
            if (p.stats_out != nullptr) {
              // per-(image, channel) sums of the new residual stream for the next InstanceNorm: lane j owns
              // column j of this warp's 32 x 32 block (conflict-free transposed read of the swizzled rows)
              const uint8_t* sl = box + quad * 32 * 128;
              float s1 = 0.f, s2q = 0.f;
              const int rows_valid = min(32, p.M - mrow);
#pragma unroll 8
              for (int r = 0; r < 32; ++r) {
                const float v = *reinterpret_cast<const float*>(sl + r * 128 + ((((lane >> 2) ^ (r & 7)) << 4) | ((lane & 3) << 2)));
                if (r < rows_valid) { s1 += v; s2q = fmaf(v, v, s2q); }
              }
              if (n + lane < p.N) {
                float* dst = p.stats_out + ((long)(mrow / p.rows_per_group) * p.N + n + lane) * 2;
                atomicAdd(dst, s1);
                atomicAdd(dst + 1, s2q);
              }
            }
            sb ^= 1;
            break;
          }
          case BF_EPI_D2S: {
            // m = (img, y, x), n = (ky, kx, co) -> out[((img*2h + 2y+ky)*2w + 2x+kx)*cout + co]   (direct stores)
            if (m < p.M) {
              const int w = p.d2s_w, h = p.d2s_h, co = p.d2s_cout;
              const int x = m % w, y = (m / w) % h, img = m / (w * h);
              if (co % 32 == 0) {
                const int q = n / co, c0 = n - q * co;
                const long pix = ((long)(img * 2 * h + 2 * y + (q >> 1)) * (2 * w) + 2 * x + (q & 1));
                uint4 u[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  if (p.is_f16) {
                    u[k].x = pack2<__half>(acc[8 * k], acc[8 * k + 1]); u[k].y = pack2<__half>(acc[8 * k + 2], acc[8 * k + 3]);
                    u[k].z = pack2<__half>(acc[8 * k + 4], acc[8 * k + 5]); u[k].w = pack2<__half>(acc[8 * k + 6], acc[8 * k + 7]);
                  } else {
                    u[k].x = pack2<__nv_bfloat16>(acc[8 * k], acc[8 * k + 1]); u[k].y = pack2<__nv_bfloat16>(acc[8 * k + 2], acc[8 * k + 3]);
                    u[k].z = pack2<__nv_bfloat16>(acc[8 * k + 4], acc[8 * k + 5]); u[k].w = pack2<__nv_bfloat16>(acc[8 * k + 6], acc[8 * k + 7]);
                  }
                }
                uint16_t* dst16 = reinterpret_cast<uint16_t*>(p.out16) + pix * co + c0;
                if ((reinterpret_cast<uintptr_t>(dst16) & 31) == 0) {
                  // 64 bytes per thread as two 256-bit stores: whole 32-byte sectors (the lanes of a warp are 2 pixels apart)
#pragma unroll
                  for (int k = 0; k < 4; k += 2)
                    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst16 + 8 * k), "r"(u[k].x),
                                 "r"(u[k].y), "r"(u[k].z), "r"(u[k].w), "r"(u[k + 1].x), "r"(u[k + 1].y), "r"(u[k + 1].z),
                                 "r"(u[k + 1].w) : "memory");
                } else {
                  uint4* dst = reinterpret_cast<uint4*>(dst16);
#pragma unroll
                  for (int k = 0; k < 4; ++k) dst[k] = u[k];
                }
              } else {
                uint16_t* out = reinterpret_cast<uint16_t*>(p.out16);
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  if (n + j < p.N) {
                    const int q = (n + j) / co, c0 = (n + j) - q * co;
                    const long pix = ((long)(img * 2 * h + 2 * y + (q >> 1)) * (2 * w) + 2 * x + (q & 1));
                    if (p.is_f16) { const __half hv = __float2half_rn(acc[j]); out[pix * co + c0] = *reinterpret_cast<const uint16_t*>(&hv); }
                    else { const __nv_bfloat16 bv = __float2bfloat16_rn(acc[j]); out[pix * co + c0] = *reinterpret_cast<const uint16_t*>(&bv); }
                  }
                }
              }
            }
            break;
          }
          default: break;
        }
      }
      if (p.in_kind != 0) {
        // the input tile buffer doubles as the output staging of this tile: release it once TMA has read it
        if (lane == 0) { tma_store_wait_read<0>(); mbar_arrive(in_empty + ib); }
        __syncwarp();
        if (++ib == p.in_bufs) { ib = 0; iphase ^= 1u; }
      }
      if (++as == 2) { as = 0; aphase ^= 1u; }
    }
    if (lane == 0) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync_all();       // the peer has finished reading this CTA's operands / signalling its barriers
  if (warp == 1) {
    tc_fence_after();
    if constexpr (CG == 2) tmem_dealloc_pair(tmem_base, kTmemCols); else tmem_dealloc(tmem_base, kTmemCols);
  }
  if (p.colsum_out != nullptr)
    for (int i = threadIdx.x; i < p.N; i += kThreads) {
      const float v = s_colsum[i];
      if (v != 0.f) atomicAdd(p.colsum_out + i, v);
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

enum MapType { kMap16 = 0, kMapF32 = 1 };

// rank-d tensor map, zero fill out of bounds.  dt: BF_BF16 / BF_F16 / BF_F32.
int make_map(CUtensorMap* map, int dt, const void* base, int rank, const uint64_t* dims,
             const uint64_t* strides_bytes /* rank-1 entries */, const uint32_t* box, CUtensorMapSwizzle swz) {
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
  const CUtensorMapDataType cdt = dt == BF_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                               : (dt == BF_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
  CUresult r = fn(map, cdt, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdim, gstr, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
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

// (rows, cols) row-major matrix with leading dimension ld (elements): 32-column boxes of `box_rows` rows
static int make_epi_map(CUtensorMap* map, int dt, const void* base, long rows, long cols, long ld, int box_rows) {
  const int es = dt == BF_F32 ? 4 : 2;
  BF_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (ld * es) % 16 == 0,
             "bf_gemm: epilogue tensors must be 16-byte aligned with 16-byte row pitch (ld=%ld)", ld);
  uint64_t dims[2] = {(uint64_t)cols, (uint64_t)rows};
  uint64_t str[1] = {(uint64_t)ld * es};
  uint32_t box[2] = {32, (uint32_t)box_rows};
  return make_map(map, dt, base, 2, dims, str, box, dt == BF_F32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
}

struct Maps { CUtensorMap a, b, in, o16, o16b, o32; };

template <int BN, bool A_MN, bool B_MN, int CG = 1>
static int launch(const Maps& mp, GemmParams& p, cudaStream_t st) {
  auto kern = gemm_tcgen05_kernel<BN, A_MN, B_MN, CG>;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [&] {
    attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
  });
  if (attr_err != cudaSuccess) return check_cuda(attr_err, "cudaFuncSetAttribute(gemm)");
  // shared-memory plan: [operand ring][epilogue input tile(s)][per-warp output slabs][barriers]
  const int b_res_bytes = p.b_resident ? p.k_iters * BN * BK * 2 : 0;
  const int stage_bytes = p.b_resident ? BM * BK * 2 : BM * BK * 2 + (BN / CG) * BK * 2;
  const int bar_bytes = (2 * kMaxStages + 9) * 8 + 16;
  const bool slabs = !(p.epilogue == BF_EPI_DGELU || p.epilogue == BF_EPI_DMUL || p.epilogue == BF_EPI_ACC32 ||
                       p.epilogue == BF_EPI_D2S);
  const int slab_bytes = slabs ? epi_warps(BN) * slab_bytes_per_warp(BN)
                               : (p.colsum_out != nullptr ? ((p.num_n_blocks * BN * 4 + 1023) / 1024) * 1024 : 0);
  p.in_bytes = p.in_kind == 0 ? 0 : BM * BN * (p.in_kind == 2 ? 4 : 2);
  int nvec = 0;
  if (p.bias) nvec = 1;
  if (p.col_scale) nvec = 3;
  if (p.col_gamma) nvec = 4;
  p.vec_cols = p.b_resident ? BN : (p.N + 31) / 32 * 32;
  const int vec_bytes = (nvec * p.vec_cols * 4 + 127) / 128 * 128;
  const int avail = kSmemBudget - 1024 - bar_bytes - slab_bytes - b_res_bytes - vec_bytes;
  const int want_stages = p.k_iters < 4 ? (p.k_iters < 2 ? 2 : p.k_iters) : 4;
  p.in_bufs = (p.in_kind != 0 && avail - 2 * p.in_bytes >= (want_stages < 3 ? want_stages : 3) * stage_bytes) ? 2 : 1;
  int stages = (avail - p.in_bufs * p.in_bytes) / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  {
    static int cap = -1;                       // BF_GEMM_STAGES=n caps the operand ring (measurements)
    if (cap < 0) { const char* e = getenv("BF_GEMM_STAGES"); cap = e ? atoi(e) : 0; }
    if (cap >= 2 && stages > cap) stages = cap;
  }
  BF_REQUIRE(stages >= 2, "bf_gemm: shared-memory plan leaves %d stages (BN=%d epilogue=%d)", stages, BN, p.epilogue);
  p.stages = stages;
  p.a_off = b_res_bytes;
  p.in_off = b_res_bytes + stages * stage_bytes;
  p.slab_off = p.in_off + p.in_bufs * p.in_bytes;
  p.vec_off = p.slab_off + slab_bytes;
  p.bar_off = p.vec_off + vec_bytes;
  const int total = p.bar_off + bar_bytes + 1024;
  const int tiles = p.num_m_blocks * p.num_n_blocks * p.split_k;
  int grid = tiles < num_sms() / CG ? tiles * CG : (num_sms() / CG) * CG;   // tiles are per CTA pair when CG = 2
  if (p.b_resident) grid = (num_sms() / p.num_n_blocks) * p.num_n_blocks;   // a CTA keeps one n block for all its tiles
  {
    static int gcap = -1;                      // BF_GEMM_GRID=n caps the persistent grid (measurements)
    if (gcap < 0) { const char* e = getenv("BF_GEMM_GRID"); gcap = e ? atoi(e) : 0; }
    if (gcap >= CG && grid > gcap) grid = (gcap / CG) * CG;
  }
  cudaError_t le;
  if constexpr (CG == 2)
    le = launch_k_cluster(kern, dim3(grid), dim3(gemm_threads(BN)), (size_t)total, st, 2u, mp.a, mp.b, mp.in, mp.o16,
                          mp.o16b, mp.o32, p);
  else
    le = launch_k(kern, dim3(grid), dim3(gemm_threads(BN)), (size_t)total, st, mp.a, mp.b, mp.in, mp.o16, mp.o16b,
                  mp.o32, p);
  count_launch();
  return check_cuda(le != cudaSuccess ? le : cudaGetLastError(), "gemm_tcgen05_kernel launch");
}

// B-resident schedule (see the producer warp) for short contractions (K <= 384) with many row blocks: it cuts the L2
// operand traffic of a 128 x 192 x 384 tile from 240 KB to 96 KB.  Measured on B200 (config-2 shapes) it does NOT pay:
// QKV 57 vs 48 us, fc1 87 vs 74 us, dGELU 91 vs 94 us, out-projection 54 vs 52 us -- these GEMMs are bound by the
// epilogue's hold on the TMEM stage, not by operand loads, and the schedule needs 128-column tiles.  Opt-in
// (BF_GEMM_RESIDENT=1) for measurements.
static bool resident_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("BF_GEMM_RESIDENT");
    on = (e != nullptr && e[0] == '1') ? 1 : 0;
  }
  return on != 0;
}
static int resident_bn(const bf_gemm_args& a) {
  // Depth-to-space GEMMs of the stem / head (K = 96, M = 160 K ... 655 K rows, 32 KB of weights per CTA) re-fetch B for
  // each of their 10 240 tiles, but keeping it resident does not help either (measured 264 vs 245 us on the
  // 655360 x 384 x 96 shape): the scatter epilogue bounds them.  BF_GEMM_D2S_RESIDENT=1 enables it for measurements.
  static int d2s_res = -1;
  if (d2s_res < 0) { const char* e = getenv("BF_GEMM_D2S_RESIDENT"); d2s_res = (e != nullptr && e[0] == '1') ? 1 : 0; }
  const bool small_k_d2s = d2s_res && a.epilogue == BF_EPI_D2S && a.K <= 2 * BK;
  if (!(resident_enabled() || small_k_d2s) || a.bn != 0) return 0;
  if (a.a_mode != BF_A_ROWMAJOR || a.split_k != 1 || a.K > 6 * BK) return 0;
  const int e = a.epilogue;
  if (!(e == BF_EPI_STORE16 || e == BF_EPI_GELU || e == BF_EPI_GELU_D || e == BF_EPI_DGELU || e == BF_EPI_DMUL ||
        e == BF_EPI_RESID || e == BF_EPI_QKV_LN || small_k_d2s)) return 0;
  const int bn = e == BF_EPI_RESID ? 64 : 128;
  if (a.N % bn != 0) return 0;
  const int nnb = a.N / bn, mb = (a.M + BM - 1) / BM;
  if (nnb > num_sms() || mb < 4 * (num_sms() / nnb)) return 0;               // at least 4 tiles per CTA
  return bn;
}

static int pick_bn(const bf_gemm_args& a) {
  if (const int rb = resident_bn(a)) return rb;
  if (a.epilogue == BF_EPI_QKV_LN) return a.bn == 128 ? 128 : 192;            // one head per tile (or two 64-column groups)
  if (a.bn == 64 || a.bn == 128 || a.bn == 192 || a.bn == 256) return a.bn;   // caller override (tuning)
  const int N = a.N;
  if (N <= 64) return 64;
  if (N <= 128 || a.epilogue == BF_EPI_RESID) return 128;                     // fp32 in/out tile: smem bound
  if (a.epilogue == BF_EPI_ACC32) return N % 192 == 0 ? 192 : 128;
  // 16-bit epilogue-input tile (saved derivative / pre-activation): 192 columns leave room for TWO input buffers, so the
  // next tile's input loads while this tile's epilogue drains (256 columns: one 64 KB buffer, -0.25 ms per step measured)
  if ((a.epilogue == BF_EPI_DMUL || a.epilogue == BF_EPI_DGELU) && N % 192 == 0) return 192;
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

// CTA pairs (cta_group::2) for the big row-major GEMMs are OPT-IN (BF_GEMM_CG=2).  Measured on B200 (round 2,
// profiles/r2e_gemm_sweep.txt, r2d_gemm_cg{1,2}.csv, same-box A/B of the replayed step): pairs change nothing -- 42.6 vs
// 49.7 us on the QKV shape, 51.8 vs 52.8 us on fc2, 55.6 vs 54.3 us on fc1, 29.21 vs 29.11 ms per step -- and the L2
// traffic ncu reports (lts__t_bytes) does not drop either: the peer CTA's half of B reaches the tensor core over the same
// SM ingress fabric as a TMA load, so the ~112 GB/s per SM (10-12 TB/s per chip) that bounds this kernel's operand feed at
// 77 FLOP per byte is spent either way.  The residual-epilogue tiles (BN = 128) were ~10 % slower as pairs.
static int cg_mode() {
  static int m = -1;
  if (m < 0) { const char* e = getenv("BF_GEMM_CG"); m = (e != nullptr && e[0] == '2') ? 2 : 1; }
  return m;
}
static int pick_cg(const bf_gemm_args& a, int bn, int resident) {
  if (cg_mode() != 2 || resident) return 1;
  if (a.a_mode != BF_A_ROWMAJOR) return 1;                       // S2D gather and the wgrad form stay single-CTA
  if (bn != 128 && bn != 192 && bn != 256) return 1;
  if (a.b_mode == BF_B_KN && (bn / 2) % 64 != 0) return 1;       // MN-major B halves must be whole 64-column blocks
  if (a.M < 256 * 16) return 1;                                  // too few row blocks to be worth pairing
  return 2;
}

}  // namespace bf

using namespace bf;

namespace bf { int launch_gemm_f32(const bf_gemm_args* a, cudaStream_t st); }

extern "C" int bf_gemm(const bf_gemm_args* a, void* stream) {
  BF_REQUIRE(a != nullptr, "bf_gemm: null args");
  BF_REQUIRE(a->M > 0 && a->N > 0 && a->K > 0, "bf_gemm: empty problem M=%d N=%d K=%d", a->M, a->N, a->K);
  if (a->dtype == BF_F32) {                        // fp32 validation backend (exact.cu)
    BF_REQUIRE(a->A && a->B, "bf_gemm: null operand");
    return launch_gemm_f32(a, static_cast<cudaStream_t>(stream));
  }
  BF_REQUIRE(a->dtype == BF_BF16 || a->dtype == BF_F16, "bf_gemm: dtype %d", a->dtype);
  BF_REQUIRE(a->A && a->B, "bf_gemm: null operand");
  BF_REQUIRE((reinterpret_cast<uintptr_t>(a->A) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->B) & 15) == 0,
             "bf_gemm: operands must be 16-byte aligned");
  const bool a_mn = a->a_mode == BF_A_KM;
  const bool b_s2d = a->b_mode == BF_B_KN_S2D;
  const bool b_mn = a->b_mode == BF_B_KN || b_s2d;
  const bool s2d = a->a_mode == BF_A_S2D;
  const int b_dtype = a->dtype;
  BF_REQUIRE(!b_s2d || (a_mn && a->split_k >= 1), "bf_gemm: B_KN_S2D is the wgrad form (needs a_mode BF_A_KM)");
  BF_REQUIRE(!(a_mn && !b_mn), "bf_gemm: A_KM requires B_KN (wgrad form)");
  BF_REQUIRE(a->split_k >= 1, "bf_gemm: split_k %d", a->split_k);
  BF_REQUIRE(a->split_k == 1 || a->epilogue == BF_EPI_ATOMIC32, "bf_gemm: split_k > 1 needs BF_EPI_ATOMIC32");

  GemmParams p{};
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.split_k = a->split_k;
  p.epilogue = a->epilogue;
  p.is_f16 = a->dtype == BF_F16;
  p.gelu_exact = gelu_exact() ? 1 : 0;
  {
    static int h2 = -1;
    if (h2 < 0) { const char* e = getenv("BF_GELU_H2"); h2 = (e != nullptr && e[0] == '0') ? 0 : 1; }
    p.gelu_h2 = h2;
  }
  p.rows_per_group = a->rows_per_group > 0 ? a->rows_per_group : 1;
  p.d2s_h = a->d2s_h; p.d2s_w = a->d2s_w; p.d2s_cout = a->d2s_cout;
  p.bias = a->bias; p.col_scale = a->col_scale; p.col_shift = a->col_shift; p.col_gamma = a->col_gamma;
  p.row_scale = a->row_scale;
  p.out16 = a->out16;
  p.ldo = a->ldo;
  p.stats_out = a->stats_out;
  p.ln_rstd = a->ln_rstd;
  p.colsum_out = a->colsum_out;
  BF_REQUIRE(a->colsum_out == nullptr || a->epilogue == BF_EPI_DGELU || a->epilogue == BF_EPI_DMUL,
             "bf_gemm: colsum_out only with BF_EPI_DGELU / BF_EPI_DMUL");
  for (const float* v : {a->bias, a->col_scale, a->col_shift, a->col_gamma})
    BF_REQUIRE((reinterpret_cast<uintptr_t>(v) & 15) == 0, "bf_gemm: per-column vectors must be 16-byte aligned");

  // epilogue operand checks
  switch (a->epilogue) {
    case BF_EPI_STORE16: case BF_EPI_GELU: case BF_EPI_GELU_D: BF_REQUIRE(a->out16, "bf_gemm: out16 required"); break;
    case BF_EPI_QKV_LN:
      BF_REQUIRE(a->out16 && a->ln_rstd, "bf_gemm: QKV_LN needs out16 and ln_rstd");
      BF_REQUIRE(a->ln_head_dim == 64 && a->N % 192 == 0,
                 "bf_gemm: QKV_LN is specialised for head_dim 64 (N a multiple of 192), got head_dim=%d N=%d",
                 a->ln_head_dim, a->N);
      BF_REQUIRE(a->bn == 0 || a->bn == 128 || a->bn == 192, "bf_gemm: QKV_LN needs the 192- or 128-column tile");
      break;
    case BF_EPI_STORE32: case BF_EPI_ATOMIC32: BF_REQUIRE(a->out32, "bf_gemm: out32 required"); break;
    case BF_EPI_RESID:
      BF_REQUIRE(a->out32 && a->in32 && a->col_gamma, "bf_gemm: RESID needs in32/out32/col_gamma");
      BF_REQUIRE((a->col_scale == nullptr) == (a->col_shift == nullptr), "bf_gemm: col_scale/col_shift pair");
      BF_REQUIRE(a->stats_out == nullptr || (a->rows_per_group % 32 == 0),
                 "bf_gemm: stats_out needs rows_per_group to be a multiple of 32");
      break;
    case BF_EPI_DGELU: case BF_EPI_DMUL: BF_REQUIRE(a->out16 && a->aux16, "bf_gemm: DGELU / DMUL need out16/aux16"); break;
    case BF_EPI_ACC32: BF_REQUIRE(a->out32 && a->in32, "bf_gemm: ACC32 needs in32/out32"); break;
    case BF_EPI_D2S:
      BF_REQUIRE(a->out16 && a->d2s_h > 0 && a->d2s_w > 0 && a->d2s_cout > 0, "bf_gemm: D2S geometry");
      BF_REQUIRE(a->N == 4 * a->d2s_cout, "bf_gemm: D2S needs N == 4*cout");
      BF_REQUIRE(a->M % (a->d2s_h * a->d2s_w) == 0, "bf_gemm: D2S M not a multiple of h*w");
      BF_REQUIRE(a->d2s_cout % 8 == 0, "bf_gemm: D2S needs cout %% 8 == 0");
      break;
    default: BF_REQUIRE(false, "bf_gemm: unknown epilogue %d", a->epilogue);
  }
  BF_REQUIRE(a->stats_out == nullptr || a->epilogue == BF_EPI_RESID || a->epilogue == BF_EPI_STORE16,
             "bf_gemm: stats_out only with BF_EPI_RESID / BF_EPI_STORE16");
  BF_REQUIRE(a->stats_out == nullptr || (a->rows_per_group % 32 == 0), "bf_gemm: stats_out needs rows_per_group %% 32 == 0");
  if (a->out16 || a->out16b || a->aux16) BF_REQUIRE(a->ldo >= a->N || a->epilogue == BF_EPI_D2S, "bf_gemm: ldo < N");
  if (a->out32 || a->in32) BF_REQUIRE(a->ld32 >= a->N, "bf_gemm: ld32 < N");

  const int bn = pick_bn(*a);
  p.b_resident = resident_bn(*a) == bn ? 1 : 0;
  p.cg = pick_cg(*a, bn, p.b_resident);
  const int cg = p.cg;
  p.ln_heads = a->epilogue == BF_EPI_QKV_LN ? a->N / 192 : 0;
  BF_REQUIRE(!b_mn || bn % 64 == 0, "bf_gemm: internal bn");
  p.num_m_blocks = (a->M + BM * cg - 1) / (BM * cg);       // row blocks of a tile: 128 rows per CTA of the pair
  p.num_n_blocks = (a->N + bn - 1) / bn;
  p.idesc = make_idesc_f16(a->dtype == BF_F16 ? 0 : 1, a_mn ? 1 : 0, b_mn ? 1 : 0, BM * cg, bn);

  Maps mp;
  int st;
  const CUtensorMapSwizzle SW128 = CU_TENSOR_MAP_SWIZZLE_128B;
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
    if ((st = make_map(&mp.a, a->dtype, a->A, 4, dims, str, box, SW128))) return st;
    // B: weight (N, ky, 2C) -> (k = 2C, ky = 2, n = N)
    uint64_t bd[3] = {(uint64_t)seg, 2, (uint64_t)a->N};
    uint64_t bs[2] = {(uint64_t)seg * 2, (uint64_t)a->ldb * 2};
    uint32_t bb[3] = {BK, 1, (uint32_t)bn};
    BF_REQUIRE(a->ldb >= 4 * C && a->ldb % 8 == 0, "bf_gemm: S2D ldb");
    if ((st = make_map(&mp.b, a->dtype, a->B, 3, bd, bs, bb, SW128))) return st;
  } else {
    BF_REQUIRE(a->lda % 8 == 0 && a->ldb % 8 == 0, "bf_gemm: leading dimensions must be multiples of 8 elements");
    const int k_total = (a->K + BK - 1) / BK;
    BF_REQUIRE(a->split_k <= k_total, "bf_gemm: split_k=%d exceeds ceil(K/64)=%d", a->split_k, k_total);
    p.k_iters = (k_total + a->split_k - 1) / a->split_k;   // a ragged last split reads zero-filled (out-of-bounds) K rows
    p.k_seg_iters = k_total;
    if (a_mn) {   // (K, M) row-major: inner dim M
      BF_REQUIRE(a->lda >= a->M, "bf_gemm: lda < M");
      uint64_t dims[2] = {(uint64_t)a->M, (uint64_t)a->K};
      uint64_t str[1] = {(uint64_t)a->lda * 2};
      uint32_t box[2] = {64, BK};
      if ((st = make_map(&mp.a, a->dtype, a->A, 2, dims, str, box, SW128))) return st;
    } else {
      BF_REQUIRE(a->lda >= a->K, "bf_gemm: lda < K");
      uint64_t dims[2] = {(uint64_t)a->K, (uint64_t)a->M};
      uint64_t str[1] = {(uint64_t)a->lda * 2};
      uint32_t box[2] = {BK, BM};
      if ((st = make_map(&mp.a, a->dtype, a->A, 2, dims, str, box, SW128))) return st;
    }
    if (b_s2d) {  // implicit patch gather, K = output pixels: (n in ky segment = 2C, xo = Wo, ky = 2, row = I*Ho)
      const int C = a->s2d_cin, Hin = a->s2d_hin, Win = a->s2d_win, I = a->s2d_images;
      BF_REQUIRE(C > 0 && Hin > 0 && Win > 0 && I > 0 && Hin % 2 == 0 && Win % 2 == 0, "bf_gemm: B_KN_S2D geometry");
      const int Ho = Hin / 2, Wo = Win / 2, seg = 2 * C;
      BF_REQUIRE(a->N == 4 * C && (long)a->K == (long)I * Ho * Wo, "bf_gemm: B_KN_S2D needs N == 4*cin and K == images*ho*wo");
      BF_REQUIRE(Wo % 64 == 0 && seg % 64 == 0, "bf_gemm: B_KN_S2D needs wo %% 64 == 0 and (2*cin) %% 64 == 0 (wo=%d cin=%d)", Wo, C);
      BF_REQUIRE(cg == 1 && !p.b_resident, "bf_gemm: B_KN_S2D runs on the single-CTA streaming schedule");
      p.b_s2d = 1; p.b_seg = seg; p.s2d_wo = Wo;
      uint64_t dims[4] = {(uint64_t)seg, (uint64_t)Wo, 2, (uint64_t)I * Ho};
      uint64_t str[3] = {(uint64_t)seg * 2, (uint64_t)Win * C * 2, (uint64_t)2 * Win * C * 2};
      uint32_t box[4] = {64, BK, 1, 1};
      if ((st = make_map(&mp.b, b_dtype, a->B, 4, dims, str, box, SW128))) return st;
    } else if (b_mn) {   // (K, N) row-major: inner dim N
      BF_REQUIRE(a->ldb >= a->N, "bf_gemm: ldb < N");
      uint64_t dims[2] = {(uint64_t)a->N, (uint64_t)a->K};
      uint64_t str[1] = {(uint64_t)a->ldb * 2};
      uint32_t box[2] = {64, BK};
      if ((st = make_map(&mp.b, b_dtype, a->B, 2, dims, str, box, SW128))) return st;
    } else {
      BF_REQUIRE(a->ldb >= a->K, "bf_gemm: ldb < K");
      uint64_t dims[2] = {(uint64_t)a->K, (uint64_t)a->N};
      uint64_t str[1] = {(uint64_t)a->ldb * 2};
      uint32_t box[2] = {BK, (uint32_t)(bn / cg)};
      if ((st = make_map(&mp.b, b_dtype, a->B, 2, dims, str, box, SW128))) return st;
    }
  }

  // epilogue tensor maps (unused ones alias the A map: never dereferenced)
  mp.in = mp.a; mp.o16 = mp.a; mp.o16b = mp.a; mp.o32 = mp.a;
  const int e = a->epilogue;
  if (e == BF_EPI_DGELU || e == BF_EPI_DMUL) {
    p.in_kind = 1;
    if ((st = make_epi_map(&mp.in, a->dtype, a->aux16, a->M, a->N, a->ldo, BM))) return st;
  } else if (e == BF_EPI_ACC32 || e == BF_EPI_RESID) {
    p.in_kind = 2;
    if ((st = make_epi_map(&mp.in, BF_F32, a->in32, a->M, a->N, a->ld32, BM))) return st;
  }
  if (e != BF_EPI_D2S && a->out16) {
    p.has_out16 = 1;
    if ((st = make_epi_map(&mp.o16, a->dtype, a->out16, a->M, a->N, a->ldo, 32))) return st;
  }
  if ((e == BF_EPI_GELU || e == BF_EPI_GELU_D || e == BF_EPI_RESID) && a->out16b) {
    p.has_out16b = 1;
    if ((st = make_epi_map(&mp.o16b, a->dtype, a->out16b, a->M, a->N, a->ldo, 32))) return st;
  }
  if (a->out32 && (e == BF_EPI_STORE32 || e == BF_EPI_ATOMIC32 || e == BF_EPI_ACC32 || e == BF_EPI_RESID)) {
    if ((st = make_epi_map(&mp.o32, BF_F32, a->out32, a->M, a->N, a->ld32, 32))) return st;
  }

  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (cg == 2) {
    if (bn == 128) return b_mn ? launch<128, false, true, 2>(mp, p, s) : launch<128, false, false, 2>(mp, p, s);
    if (bn == 192 && !b_mn) return launch<192, false, false, 2>(mp, p, s);
    if (bn == 256) return b_mn ? launch<256, false, true, 2>(mp, p, s) : launch<256, false, false, 2>(mp, p, s);
    set_error("bf_gemm: no pair kernel for BN=%d b_mode=%d", bn, (int)b_mn);
    return BF_ERR_INVALID;
  }
#define BF_DISPATCH(BN_)                                                          \
  if (bn == BN_) {                                                                \
    if (a_mn) return launch<BN_, true, true>(mp, p, s);                           \
    if (b_mn) return launch<BN_, false, true>(mp, p, s);                          \
    return launch<BN_, false, false>(mp, p, s);                                   \
  }
  BF_DISPATCH(64)
  BF_DISPATCH(128)
  BF_DISPATCH(192)
  BF_DISPATCH(256)
#undef BF_DISPATCH
  set_error("bf_gemm: no kernel for BN=%d", bn);
  return BF_ERR_INVALID;
}
