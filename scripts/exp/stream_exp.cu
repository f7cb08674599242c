// Stand-alone experiment (not part of the library): which load path streams an InstanceNorm-apply-like pass fastest
// on B200 at the config-2 shape (I = 40 images x P = 1024 rows x C = 384 channels)?
//   pass A ("apply"):    out16[r, c] = bf16(x32[r, c] * a[img, c] + b[img, c])             63 MB in, 31 MB out
//   pass B ("bwd2_add"): out32[r, c] = ka*g16 + kb*x32 + kc + add32                       157 MB in, 63 MB out
// Variants:
//   0  per-thread cp.async ring, 8 consecutive channels per thread (the library's norm.cu pattern)
//   1  same, data copies issued before the per-channel parameter loads
//   2  256-bit register loads (LDG.256), U rows in flight per thread, no shared memory
//   3  TMA 1-D bulk copies into a shared-memory ring by one producer thread, consumers read shared memory
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o stream_exp scripts/exp/stream_exp.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

constexpr int I_ = 40, P_ = 1024, C_ = 384;
constexpr int kNT = 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(void* s, const void* g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(s)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack2(uint32_t u) {
  return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}
__device__ __forceinline__ void ldg256(const float* p, float (&v)[8]) {
  asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(p));
}
__device__ __forceinline__ void stg256(float* p, const float (&v)[8]) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
               "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
}
// mbarrier + bulk copy
__device__ __forceinline__ void mbar_init(uint64_t* b, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t* b, uint32_t ph) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(b)), "r"(ph) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t ph) {
  for (uint32_t it = 0; it < (1u << 24); ++it) if (mbar_try(b, ph)) return;
  __trap();
}
__device__ __forceinline__ void bulk_g2s(void* s, const void* g, uint32_t bytes, uint64_t* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(s)), "l"(g),
               "r"(bytes), "r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* g, const void* s, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g), "r"(smem_u32(s)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

struct Args {
  const float* x32; const __nv_bfloat16* g16; const float* add32;
  __nv_bfloat16* out16; float* out32;
  const float* stats; const float* weight; const float* bias; const float* red;
  int splits, rows_per_split;
};

__device__ __forceinline__ void params_apply(const Args& a, int img, int c0, float (&ka)[8], float (&kb)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float2 sq = *reinterpret_cast<const float2*>(a.stats + 2 * ((long)img * C_ + c0 + j));
    const float mean = sq.x * (1.f / P_);
    const float rstd = rsqrtf(fmaxf(sq.y * (1.f / P_) - mean * mean, 0.f) + 1e-5f);
    const float w = a.weight[c0 + j];
    ka[j] = rstd * w; kb[j] = a.bias[c0 + j] - mean * rstd * w;
  }
}
__device__ __forceinline__ void params_bwd(const Args& a, int img, int c0, float (&ka)[8], float (&kb)[8], float (&kc)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const long idx = (long)img * C_ + c0 + j;
    const float2 sq = *reinterpret_cast<const float2*>(a.stats + 2 * idx);
    const float mean = sq.x * (1.f / P_);
    const float rstd = rsqrtf(fmaxf(sq.y * (1.f / P_) - mean * mean, 0.f) + 1e-5f);
    const float k = rstd * a.weight[c0 + j];
    const float m1 = a.red[2 * idx] * (1.f / P_), m2 = a.red[2 * idx + 1] * (1.f / P_);
    ka[j] = k; kb[j] = -k * m2 * rstd; kc[j] = -k * m1 + k * m2 * rstd * mean;
  }
}

// ------------------------------------------------------------------ variant 0 / 1: cp.async ring
template <bool BWD, bool EARLY>
__global__ void __launch_bounds__(kNT, 3) k_ring(Args a) {
  constexpr int TX = C_ / 8, TY = kNT / TX;                 // 48 x 5
  constexpr int NSLOT = BWD ? 5 : 2, S = (65536 / (NSLOT * kNT * 16)) > 8 ? 8 : (65536 / (NSLOT * kNT * 16));
  extern __shared__ __align__(16) uint4 ring[];
  const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
  const bool active = ty < TY;
  const int img = blockIdx.y, c0 = tx * 8;
  const int r0 = blockIdx.x * a.rows_per_split, r1 = min(P_, r0 + a.rows_per_split);
  const int n_it = (a.rows_per_split + TY - 1) / TY;
  const float* xb = a.x32 + (long)img * P_ * C_ + c0;
  const __nv_bfloat16* gb = a.g16 + (long)img * P_ * C_ + c0;
  const float* ab = a.add32 + (long)img * P_ * C_ + c0;
  float ka[8], kb[8], kc[8];
  auto issue = [&](int row, int st) {
    uint4* s = ring + (st * NSLOT) * kNT + threadIdx.x;
    cp_async16(s, xb + (long)row * C_); cp_async16(s + kNT, xb + (long)row * C_ + 4);
    if (BWD) {
      cp_async16(s + 2 * kNT, gb + (long)row * C_);
      cp_async16(s + 3 * kNT, ab + (long)row * C_); cp_async16(s + 4 * kNT, ab + (long)row * C_ + 4);
    }
  };
  if (!EARLY && active) { if (BWD) params_bwd(a, img, c0, ka, kb, kc); else params_apply(a, img, c0, ka, kb); }
#pragma unroll
  for (int it = 0; it < S; ++it) {
    const int row = r0 + ty + it * TY;
    if (active && it < n_it && row < r1) issue(row, it);
    cp_commit();
  }
  if (EARLY && active) { if (BWD) params_bwd(a, img, c0, ka, kb, kc); else params_apply(a, img, c0, ka, kb); }
  int st = 0;
  for (int it = 0; it < n_it; ++it) {
    cp_wait<S - 1>();
    const int row = r0 + ty + it * TY;
    if (active && row < r1) {
      const uint4* s = ring + (st * NSLOT) * kNT + threadIdx.x;
      const float4 x0 = *reinterpret_cast<const float4*>(s), x1 = *reinterpret_cast<const float4*>(s + kNT);
      float xv[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
      if (BWD) {
        const uint4 g = s[2 * kNT];
        const float2 g0 = unpack2(g.x), g1 = unpack2(g.y), g2 = unpack2(g.z), g3 = unpack2(g.w);
        const float gv[8] = {g0.x, g0.y, g1.x, g1.y, g2.x, g2.y, g3.x, g3.y};
        const float4 a0 = *reinterpret_cast<const float4*>(s + 3 * kNT), a1 = *reinterpret_cast<const float4*>(s + 4 * kNT);
        const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = fmaf(ka[j], gv[j], fmaf(kb[j], xv[j], kc[j])) + av[j];
        float* op = a.out32 + ((long)img * P_ + row) * C_ + c0;
        reinterpret_cast<float4*>(op)[0] = make_float4(o[0], o[1], o[2], o[3]);
        reinterpret_cast<float4*>(op)[1] = make_float4(o[4], o[5], o[6], o[7]);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) xv[j] = fmaf(xv[j], ka[j], kb[j]);
        uint4 u; u.x = pack2(xv[0], xv[1]); u.y = pack2(xv[2], xv[3]); u.z = pack2(xv[4], xv[5]); u.w = pack2(xv[6], xv[7]);
        *reinterpret_cast<uint4*>(a.out16 + ((long)img * P_ + row) * C_ + c0) = u;
      }
    }
    const int nrow = row + S * TY;
    if (active && it + S < n_it && nrow < r1) issue(nrow, st);
    cp_commit();
    if (++st == S) st = 0;
  }
  cp_wait<0>();
}

// ------------------------------------------------------------------ variant 2: 256-bit register loads
template <bool BWD, int U, int MINB>
__global__ void __launch_bounds__(kNT, MINB) k_reg(Args a) {
  constexpr int TX = C_ / 8, TY = kNT / TX;
  const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
  if (ty >= TY) return;
  const int img = blockIdx.y, c0 = tx * 8;
  const int r0 = blockIdx.x * a.rows_per_split, r1 = min(P_, r0 + a.rows_per_split);
  const float* xb = a.x32 + (long)img * P_ * C_ + c0;
  const __nv_bfloat16* gb = a.g16 + (long)img * P_ * C_ + c0;
  const float* ab = a.add32 + (long)img * P_ * C_ + c0;
  float xv[U][8], av[BWD ? U : 1][8];
  uint4 gv[BWD ? U : 1];
  auto load = [&](int u, int row) {
    if (row < r1) {
      ldg256(xb + (long)row * C_, xv[u]);
      if (BWD) { ldg256(ab + (long)row * C_, av[u]); gv[u] = __ldg(reinterpret_cast<const uint4*>(gb + (long)row * C_)); }
    }
  };
#pragma unroll
  for (int u = 0; u < U; ++u) load(u, r0 + ty + u * TY);
  float ka[8], kb[8], kc[8];
  if (BWD) params_bwd(a, img, c0, ka, kb, kc); else params_apply(a, img, c0, ka, kb);
  for (int base = r0 + ty; base < r1; base += U * TY) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int row = base + u * TY;
      if (row < r1) {
        if (BWD) {
          const float2 g0 = unpack2(gv[u].x), g1 = unpack2(gv[u].y), g2 = unpack2(gv[u].z), g3 = unpack2(gv[u].w);
          const float g[8] = {g0.x, g0.y, g1.x, g1.y, g2.x, g2.y, g3.x, g3.y};
          float o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = fmaf(ka[j], g[j], fmaf(kb[j], xv[u][j], kc[j])) + av[u][j];
          stg256(a.out32 + ((long)img * P_ + row) * C_ + c0, o);
        } else {
          float o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = fmaf(xv[u][j], ka[j], kb[j]);
          uint4 w; w.x = pack2(o[0], o[1]); w.y = pack2(o[2], o[3]); w.z = pack2(o[4], o[5]); w.w = pack2(o[6], o[7]);
          *reinterpret_cast<uint4*>(a.out16 + ((long)img * P_ + row) * C_ + c0) = w;
        }
      }
      load(u, row + U * TY);
    }
  }
}

// ------------------------------------------------------------------ variant 3: TMA bulk ring
// block = 32 (producer warp) + 256 consumers; stage = RB rows of every operand, contiguous in global memory
template <bool BWD, int RB, int NS>
__global__ void __launch_bounds__(kNT + 32, 1) k_bulk(Args a) {
  constexpr int TX = C_ / 8, TY = kNT / TX;           // 48 x 5 consumers
  constexpr int XB = RB * C_ * 4, GB = BWD ? RB * C_ * 2 : 0, AB = BWD ? RB * C_ * 4 : 0;
  constexpr int STAGE = XB + GB + AB;
  extern __shared__ __align__(128) uint8_t sm[];
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + NS * STAGE);
  uint64_t* empty = full + NS;
  const int img = blockIdx.y;
  const int r0 = blockIdx.x * a.rows_per_split, r1 = min(P_, r0 + a.rows_per_split);
  const int nblk = (r1 - r0 + RB - 1) / RB;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, kNT / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    if (threadIdx.x == 0) {
      int st = 0; uint32_t ph = 0;
      for (int b = 0; b < nblk; ++b) {
        mbar_wait(empty + st, ph ^ 1u);
        const int row = r0 + b * RB;
        const int rows = min(RB, r1 - row);
        uint8_t* s = sm + st * STAGE;
        const long off = ((long)img * P_ + row) * C_;
        mbar_expect(full + st, rows * C_ * (BWD ? 10 : 4));
        bulk_g2s(s, a.x32 + off, rows * C_ * 4, full + st);
        if (BWD) {
          bulk_g2s(s + XB, a.g16 + off, rows * C_ * 2, full + st);
          bulk_g2s(s + XB + GB, a.add32 + off, rows * C_ * 4, full + st);
        }
        if (++st == NS) { st = 0; ph ^= 1u; }
      }
    }
    return;
  }
  const int t = threadIdx.x - 32;
  const int tx = t % TX, ty = t / TX;
  const bool active = ty < TY;
  const int c0 = tx * 8;
  float ka[8], kb[8], kc[8];
  if (active) { if (BWD) params_bwd(a, img, c0, ka, kb, kc); else params_apply(a, img, c0, ka, kb); }
  int st = 0; uint32_t ph = 0;
  for (int b = 0; b < nblk; ++b) {
    mbar_wait(full + st, ph);
    const uint8_t* s = sm + st * STAGE;
    const int row0 = r0 + b * RB;
    if (active) {
#pragma unroll
      for (int k = 0; k < RB / TY; ++k) {
        const int rl = ty + k * TY;
        const int row = row0 + rl;
        if (row < r1) {
          const float4* xp = reinterpret_cast<const float4*>(s + (rl * C_ + c0) * 4);
          const float4 x0 = xp[0], x1 = xp[1];
          const float xv[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
          if (BWD) {
            const uint4 g = *reinterpret_cast<const uint4*>(s + XB + (rl * C_ + c0) * 2);
            const float2 g0 = unpack2(g.x), g1 = unpack2(g.y), g2 = unpack2(g.z), g3 = unpack2(g.w);
            const float gv[8] = {g0.x, g0.y, g1.x, g1.y, g2.x, g2.y, g3.x, g3.y};
            const float4* ap = reinterpret_cast<const float4*>(s + XB + GB + (rl * C_ + c0) * 4);
            const float4 a0 = ap[0], a1 = ap[1];
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = fmaf(ka[j], gv[j], fmaf(kb[j], xv[j], kc[j])) + av[j];
            stg256(a.out32 + ((long)img * P_ + row) * C_ + c0, o);
          } else {
            float o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = fmaf(xv[j], ka[j], kb[j]);
            uint4 w; w.x = pack2(o[0], o[1]); w.y = pack2(o[2], o[3]); w.z = pack2(o[4], o[5]); w.w = pack2(o[6], o[7]);
            *reinterpret_cast<uint4*>(a.out16 + ((long)img * P_ + row) * C_ + c0) = w;
          }
        }
      }
    }
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(empty + st);
    if (++st == NS) { st = 0; ph ^= 1u; }
  }
}

int main(int argc, char** argv) {
  const int NB = 6;                       // buffer sets cycled through (>> L2)
  const long n = (long)I_ * P_ * C_;
  std::vector<Args> sets(NB);
  float *stats, *weight, *bias, *red;
  CK(cudaMalloc(&stats, I_ * C_ * 2 * 4)); CK(cudaMalloc(&red, I_ * C_ * 2 * 4));
  CK(cudaMalloc(&weight, C_ * 4)); CK(cudaMalloc(&bias, C_ * 4));
  {
    std::vector<float> h(I_ * C_ * 2);
    for (int i = 0; i < I_ * C_; ++i) { h[2 * i] = 0.1f * P_; h[2 * i + 1] = 1.5f * P_; }
    CK(cudaMemcpy(stats, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(red, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    std::vector<float> w(C_, 1.f);
    CK(cudaMemcpy(weight, w.data(), C_ * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(bias, w.data(), C_ * 4, cudaMemcpyHostToDevice));
  }
  for (int s = 0; s < NB; ++s) {
    Args& a = sets[s];
    float *x, *ad, *o32; __nv_bfloat16 *g, *o16;
    CK(cudaMalloc(&x, n * 4)); CK(cudaMalloc(&ad, n * 4)); CK(cudaMalloc(&o32, n * 4));
    CK(cudaMalloc(&g, n * 2)); CK(cudaMalloc(&o16, n * 2));
    CK(cudaMemset(x, 0, n * 4)); CK(cudaMemset(ad, 0, n * 4)); CK(cudaMemset(g, 0, n * 2));
    a.x32 = x; a.add32 = ad; a.out32 = o32; a.g16 = g; a.out16 = o16;
    a.stats = stats; a.weight = weight; a.bias = bias; a.red = red;
  }
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  auto timeit = [&](const char* name, auto launch, double mb) {
    for (int i = 0; i < NB; ++i) launch(sets[i % NB]);
    CK(cudaDeviceSynchronize());
    const int iters = 30;
    CK(cudaEventRecord(e0));
    for (int i = 0; i < iters; ++i) launch(sets[i % NB]);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    CK(cudaGetLastError());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    const double us = ms * 1e3 / iters;
    printf("%-44s %7.1f us  %7.0f GB/s\n", name, us, mb / us * 1e-3 * 1e6 / 1e6);
  };
  auto with_split = [&](Args a, int ctas_per_sm) {
    a.splits = sms * ctas_per_sm / I_;
    a.rows_per_split = (P_ + a.splits - 1) / a.splits;
    return a;
  };
  const double mbA = n * 6.0 / 1e6, mbB = n * 14.0 / 1e6;
  CK(cudaFuncSetAttribute(k_ring<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  CK(cudaFuncSetAttribute(k_ring<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  CK(cudaFuncSetAttribute(k_ring<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  CK(cudaFuncSetAttribute(k_ring<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  timeit("A v0 ring", [&](Args a) { a = with_split(a, 3); k_ring<false, false><<<dim3(a.splits, I_), kNT, 65536>>>(a); }, mbA);
  timeit("A v1 ring early-issue", [&](Args a) { a = with_split(a, 3); k_ring<false, true><<<dim3(a.splits, I_), kNT, 65536>>>(a); }, mbA);
  timeit("A v2 ldg256 U=4 x3/SM", [&](Args a) { a = with_split(a, 3); k_reg<false, 4, 3><<<dim3(a.splits, I_), kNT>>>(a); }, mbA);
  timeit("A v2 ldg256 U=4 x6/SM", [&](Args a) { a = with_split(a, 6); k_reg<false, 4, 6><<<dim3(a.splits, I_), kNT>>>(a); }, mbA);
  timeit("A v2 ldg256 U=8 x3/SM", [&](Args a) { a = with_split(a, 3); k_reg<false, 8, 3><<<dim3(a.splits, I_), kNT>>>(a); }, mbA);
  timeit("A v2 ldg256 U=2 x8/SM", [&](Args a) { a = with_split(a, 8); k_reg<false, 2, 8><<<dim3(a.splits, I_), kNT>>>(a); }, mbA);
  timeit("A v2 ldg256 U=4 x12/SM (2.7 waves)", [&](Args a) { a = with_split(a, 12); k_reg<false, 4, 4><<<dim3(a.splits, I_), kNT>>>(a); }, mbA);
#define BULK(BWD, RB, NS, PER_SM, label, mb)                                                                   \
  {                                                                                                            \
    const int smem = NS * RB * C_ * (BWD ? 10 : 4) + 2 * NS * 8;                                               \
    CK(cudaFuncSetAttribute(k_bulk<BWD, RB, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));          \
    timeit(label, [&](Args a) { a = with_split(a, PER_SM); k_bulk<BWD, RB, NS><<<dim3(a.splits, I_), kNT + 32, smem>>>(a); }, mb); \
  }
  BULK(false, 10, 4, 3, "A v3 bulk RB=10 NS=4 x3/SM (61 KB)", mbA)
  BULK(false, 10, 6, 2, "A v3 bulk RB=10 NS=6 x2/SM (92 KB)", mbA)
  BULK(false, 20, 3, 2, "A v3 bulk RB=20 NS=3 x2/SM (92 KB)", mbA)
  BULK(false, 5, 8, 3, "A v3 bulk RB=5 NS=8 x3/SM (61 KB)", mbA)
  BULK(false, 20, 6, 1, "A v3 bulk RB=20 NS=6 x1/SM (184 KB)", mbA)
  timeit("B v0 ring", [&](Args a) { a = with_split(a, 3); k_ring<true, false><<<dim3(a.splits, I_), kNT, 65536>>>(a); }, mbB);
  timeit("B v1 ring early-issue", [&](Args a) { a = with_split(a, 3); k_ring<true, true><<<dim3(a.splits, I_), kNT, 65536>>>(a); }, mbB);
  timeit("B v2 ldg256 U=2 x3/SM", [&](Args a) { a = with_split(a, 3); k_reg<true, 2, 3><<<dim3(a.splits, I_), kNT>>>(a); }, mbB);
  timeit("B v2 ldg256 U=2 x6/SM", [&](Args a) { a = with_split(a, 6); k_reg<true, 2, 6><<<dim3(a.splits, I_), kNT>>>(a); }, mbB);
  timeit("B v2 ldg256 U=4 x3/SM", [&](Args a) { a = with_split(a, 3); k_reg<true, 4, 3><<<dim3(a.splits, I_), kNT>>>(a); }, mbB);
  timeit("B v2 ldg256 U=4 x4/SM", [&](Args a) { a = with_split(a, 4); k_reg<true, 4, 4><<<dim3(a.splits, I_), kNT>>>(a); }, mbB);
  BULK(true, 5, 5, 2, "B v3 bulk RB=5 NS=5 x2/SM (96 KB)", mbB)
  BULK(true, 5, 3, 3, "B v3 bulk RB=5 NS=3 x3/SM (58 KB)", mbB)
  BULK(true, 10, 5, 1, "B v3 bulk RB=10 NS=5 x1/SM (192 KB)", mbB)
  BULK(true, 10, 2, 2, "B v3 bulk RB=10 NS=2 x2/SM (77 KB)", mbB)
  // reference points: device-to-device copies of the same byte counts
  timeit("memcpy 63 MB -> (94 MB moved... as 47+47)", [&](Args a) { cudaMemcpyAsync(a.out32, a.x32, n * 3, cudaMemcpyDeviceToDevice); }, mbA);
  timeit("memcpy 110 MB (220 MB moved)", [&](Args a) { cudaMemcpyAsync(a.out32, a.x32, n * 4, cudaMemcpyDeviceToDevice); cudaMemcpyAsync(a.out16, a.g16, n * 2, cudaMemcpyDeviceToDevice); cudaMemcpyAsync((void*)a.add32, a.x32, n * 1, cudaMemcpyDeviceToDevice); }, mbB);
  return 0;
}
