// Shared pieces of the pre-normalised fast attention kernels (attention_fast.cu: L <= 32, attention_fast64.cu: L <= 64):
// launch parameters, the per-CTA affine / bias tables, ldmatrix / mma.sync fragment helpers on 128B-swizzled tiles.
#pragma once

#include <cstdlib>

#include "common.cuh"

namespace bf {

using bf16 = __nv_bfloat16;

constexpr int FD = 64;                       // head dim
constexpr int FLP = 32;                      // rows per tile
constexpr int kTile = FLP * FD * 2;          // 4096 B: one 32 x 64 bf16 tile, rows of 128 B, TMA 128B swizzle
constexpr int kFwdWarps = 16;                // one CTA per SM: per-warp tiles fill the shared memory
constexpr int kBwdWarps = 12;

struct FastParams {
  const float* rstd;                 // (tokens, heads, 2)
  int heads, L, G;
  int inner, tiles_per_outer;        // sequences per outer index, tiles (of G sequences) per outer index
  long n_tiles;
  long outer_stride, inner_stride, tok_stride;   // token index = outer*outer_stride + seq*inner_stride + pos*tok_stride
  const float* qn_w; const float* qn_b; const float* kn_w; const float* kn_b;
  const float* bias_emb; const int* bucket; const float* scale_factor;
  float out_scale; int accumulate;
  float* d_qn_w; float* d_qn_b; float* d_kn_w; float* d_kn_b; float* d_bias_emb; float* d_scale_factor;
  float* d_qkv_bias;                 // [heads * 3 * 64] or null: += column sums of the dqkv written by this launch
};

// per-warp shared memory (tiles first: every tile must be 1 KiB aligned for the 128B swizzle)
struct FwdWarp {
  static constexpr int kQ = 0, kK = kTile, kV = 2 * kTile;
  static constexpr int kBar = 3 * kTile;                   // mbarrier (8 B)
  static constexpr int kRowGp = kBar + 16;                 // int[32]
  static constexpr int kBytes = ((kRowGp + FLP * 4) + 1023) / 1024 * 1024;
};
struct BwdWarp {
  static constexpr int kQ = 0, kK = kTile, kV = 2 * kTile, kDo = 3 * kTile;
  static constexpr int kBar = 4 * kTile;
  static constexpr int kRowGp = kBar + 16;
  static constexpr int kRstd = kRowGp + FLP * 4;           // float[32][2]
  static constexpr int kBytes = ((kRstd + FLP * 8) + 1023) / 1024 * 1024;
};
// CTA-level tables (after the per-warp regions): affine parts of the two LayerNorms as packed bf16 (pairs for the
// K-contiguous fragments, splats for the transposed fragments) and fp32 weights for the backward
constexpr int kTabPairs = 0;        // uint32[4][32]: aq, bq, ak, bk  (pair i = columns 2i, 2i+1)
constexpr int kTabSplat = 512;      // uint32[4][64]: aq, bq, ak, bk  (both halves = column i)
constexpr int kTabW = 512 + 1024;   // float[2][64]: wq, wk
constexpr int kTabBrel = 512 + 1024 + 512;   // float[heads][64]: relative-position bias of every head, indexed by rel + L - 1
constexpr int kTabFixed = 512 + 1024 + 512;
__host__ __device__ constexpr int tab_bytes(int heads) { return kTabFixed + heads * 256; }

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

// byte offset of (row, element column) inside a 128B-swizzled 32 x 64 bf16 tile
__device__ __forceinline__ int swz(int row, int col) {
  const int cb = col * 2;
  return row * 128 + ((((cb >> 4) ^ (row & 7)) << 4) | (cb & 15));
}
// A fragment (16 x 16 at rows m0, K columns k0) of a row-major [row][k] tile
__device__ __forceinline__ void frag_a(uint32_t (&a)[4], const uint8_t* tile, int m0, int k0, int lane) {
  ldsm4(a, tile + swz(m0 + (lane & 15), k0 + (lane >> 4) * 8));
}
// A fragment of the TRANSPOSE of a [k][m] tile (A[m][k] = T[k][m]); m0 may include a column offset
__device__ __forceinline__ void frag_a_t(uint32_t (&a)[4], const uint8_t* tile, int m0, int k0, int lane) {
  ldsm4t(a, tile + swz(k0 + (lane & 7) + (lane >> 4) * 8, m0 + ((lane >> 3) & 1) * 8));
}
// B fragments of two adjacent n tiles (n0, n0 + 8) at k0 from a [n][k] tile: b[0..1] tile 0, b[2..3] tile 1
__device__ __forceinline__ void frag_b(uint32_t (&b)[4], const uint8_t* tile, int n0, int k0, int lane) {
  ldsm4(b, tile + swz(n0 + (lane & 7) + (lane >> 4) * 8, k0 + ((lane >> 3) & 1) * 8));
}
// ... from a [k][n] tile
__device__ __forceinline__ void frag_b_t(uint32_t (&b)[4], const uint8_t* tile, int n0, int k0, int lane) {
  ldsm4t(b, tile + swz(k0 + (lane & 7) + ((lane >> 3) & 1) * 8, n0 + (lane >> 4) * 8));
}

__device__ __forceinline__ void fill_tables(const FastParams& p, uint8_t* tab, int tid, int nthreads, int brel_stride = 64) {
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
  float* brel = reinterpret_cast<float*>(tab + kTabBrel);
  for (int i = tid; i < p.heads * brel_stride; i += nthreads) {
    const int h = i / brel_stride, r = i - h * brel_stride;
    brel[i] = r < 2 * p.L - 1 ? p.bias_emb[p.bucket[r] * p.heads + h] : 0.f;
  }
}

struct Item { int head, s_in0, s_out; };

}  // namespace bf
