// Shared device/host helpers for the bubbleformer_b200 kernels (sm_100a only).
//
// Everything here is inline PTX for Blackwell: mbarrier (with bounded waits that trap
// instead of hanging), TMA bulk-tensor loads, tcgen05 MMA / TMEM management.
#pragma once

#include <cuda.h>            // CUtensorMap types only; the driver entry point is fetched at run time
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/bubbleformer_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ != 1000)
#error "bubbleformer_b200 kernels are written for sm_100a only"
#endif

namespace bf {

// ---------------------------------------------------------------------------------------------
// host-side status / error string (thread local, read back through bf_last_error())
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int  check_cuda(cudaError_t e, const char* what);
int  num_sms();
void count_launch(int n = 1);
bool gelu_exact();          // bf_set_gelu_mode / BF_GELU_ERF=1: exact-erf GELU instead of the tanh form
// rank-d TMA tensor map (zero fill out of bounds); dt: BF_BF16 / BF_F16 / BF_F32.  Defined in gemm_sm100.cu.
int  make_map(CUtensorMap* map, int dt, const void* base, int rank, const uint64_t* dims,
              const uint64_t* strides_bytes /* rank-1 entries */, const uint32_t* box, CUtensorMapSwizzle swz);

#define BF_REQUIRE(cond, ...)                                      \
  do {                                                             \
    if (!(cond)) {                                                 \
      bf::set_error(__VA_ARGS__);                                  \
      return BF_ERR_INVALID;                                       \
    }                                                              \
  } while (0)

#define BF_LAUNCH_CHECK(what)                                      \
  do {                                                             \
    int _st = bf::check_cuda(cudaGetLastError(), what);            \
    if (_st) return _st;                                           \
  } while (0)

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch.  Every kernel of this library is launched with the programmatic
// stream-serialization attribute and starts with pdl_prologue_done(): its CTAs may become resident while the
// previous kernel of the stream drains (launch latency, barrier / TMEM / shared-memory set-up overlap that tail),
// but they touch global memory only after griddepcontrol.wait, i.e. after the previous grid has completed and
// flushed.  The dependents of this grid are released right away: they wait the same way.  BF_PDL=0 disables it.
// ---------------------------------------------------------------------------------------------
bool pdl_enabled();

__device__ __forceinline__ void pdl_prologue_done() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                            Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
// same, as thread-block clusters of `cluster_x` CTAs (grid.x must be a multiple of it)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k_cluster(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                    unsigned cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cluster_x;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// ---------------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// exact-erf GELU and its derivative (reference: nn.GELU() default, linear_layers.py:16)
__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f));
}
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
  const float pdf = 0.39894228040143268f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

// tanh-form GELU for the MLP GEMM epilogues and the stem / head norm passes, where the exact erf would make the ALU
// work (not the MMAs / the HBM stream) the bound:
// one MUFU.TANH + ~6 FMA-pipe instructions per element.  |gelu_tanh - gelu_erf| <= 4.8e-4 absolute, 2e-4 rel-L2 on
// N(0,1) pre-activations -- an eighth of the bf16 rounding (1.7e-3) the stored activation carries anyway.
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float gelu_tanh(float x) {
  const float u = x * fmaf(0.0356774081f, x * x, 0.7978845608f);     // sqrt(2/pi) * (x + 0.044715 x^3)
  const float hx = 0.5f * x;
  return fmaf(hx, tanh_approx(u), hx);
}
__device__ __forceinline__ float gelu_tanh_grad(float x) {
  const float x2 = x * x;
  const float th = tanh_approx(x * fmaf(0.0356774081f, x2, 0.7978845608f));
  const float du = fmaf(0.1070322243f, x2, 0.7978845608f);           // d/dx of the tanh argument
  const float hx = 0.5f * x;
  return fmaf(hx * du, fmaf(-th, th, 1.0f), fmaf(0.5f, th, 0.5f));
}

// run-time selected form (kernel parameter, warp uniform): mode 2 = exact erf, otherwise the tanh form
__device__ __forceinline__ float gelu_fwd(float x, int exact) { return exact ? gelu_erf(x) : gelu_tanh(x); }
__device__ __forceinline__ float gelu_bwd(float x, int exact) { return exact ? gelu_erf_grad(x) : gelu_tanh_grad(x); }

// value and derivative together (the tanh / erf is shared): the forward of the MLP stores gelu'(pre) for its backward
__device__ __forceinline__ void gelu_both(float x, int exact, float& g, float& d) {
  if (exact) {
    const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
    const float pdf = 0.39894228040143268f * __expf(-0.5f * x * x);
    g = x * cdf;
    d = fmaf(x, pdf, cdf);
  } else {
    const float x2 = x * x;
    const float th = tanh_approx(x * fmaf(0.0356774081f, x2, 0.7978845608f));
    const float hx = 0.5f * x;
    g = fmaf(hx, th, hx);
    d = fmaf(hx * fmaf(0.1070322243f, x2, 0.7978845608f), fmaf(-th, th, 1.0f), fmaf(0.5f, th, 0.5f));
  }
}

// Packed half-precision evaluation of the tanh-form GELU and its derivative for TWO elements (fc1 epilogue, bf16 / fp16
// storage): 12 HFMA2-class instructions + one MUFU.TANH per pair instead of 24 + 2 in fp32.  The result is rounded to a
// 16-bit storage type anyway (bf16: 8 mantissa bits; the f16 intermediates keep 11).  The polynomial arguments are
// evaluated on x clamped to +-16 (tanh is +-1 in f16 beyond |x| ~ 4.6), so x^2 cannot overflow; value and derivative
// saturate to x / 0 and 1 / 0.
__device__ __forceinline__ __half2 tanh_approx_h2(__half2 x) {
  uint32_t r, xi = *reinterpret_cast<uint32_t*>(&x);
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(r) : "r"(xi));
  return *reinterpret_cast<__half2*>(&r);
}
__device__ __forceinline__ void gelu_both_h2(float x0, float x1, __half2& g, __half2& d) {
  const __half2 h = __floats2half2_rn(x0, x1);
  const __half2 hc = __hmax2(__hmin2(h, __float2half2_rn(16.f)), __float2half2_rn(-16.f));
  const __half2 x2 = __hmul2(hc, hc);
  const __half2 c0 = __float2half2_rn(0.7978845608f);
  const __half2 th = tanh_approx_h2(__hmul2(hc, __hfma2(x2, __float2half2_rn(0.0356774081f), c0)));
  const __half2 half = __float2half2_rn(0.5f);
  const __half2 hx = __hmul2(h, half);
  g = __hfma2(hx, th, hx);
  const __half2 du = __hfma2(x2, __float2half2_rn(0.1070322243f), c0);
  const __half2 t2 = __hfma2(__hneg2(th), th, __float2half2_rn(1.f));
  d = __hfma2(__hmul2(__hmul2(hc, half), du), t2, __hfma2(th, half, half));
}
// derivative only, two elements (InstanceNorm backward through a GELU: stem / head stages)
__device__ __forceinline__ float2 gelu_grad_h2(float y0, float y1) {
  const __half2 h = __floats2half2_rn(y0, y1);
  const __half2 hc = __hmax2(__hmin2(h, __float2half2_rn(16.f)), __float2half2_rn(-16.f));
  const __half2 x2 = __hmul2(hc, hc);
  const __half2 c0 = __float2half2_rn(0.7978845608f);
  const __half2 th = tanh_approx_h2(__hmul2(hc, __hfma2(x2, __float2half2_rn(0.0356774081f), c0)));
  const __half2 half = __float2half2_rn(0.5f);
  const __half2 du = __hfma2(x2, __float2half2_rn(0.1070322243f), c0);
  const __half2 t2 = __hfma2(__hneg2(th), th, __float2half2_rn(1.f));
  return __half22float2(__hfma2(__hmul2(__hmul2(hc, half), du), t2, __hfma2(th, half, half)));
}
template <typename T> __device__ __forceinline__ uint32_t pack_h2(__half2 v);
template <> __device__ __forceinline__ uint32_t pack_h2<__half>(__half2 v) { return *reinterpret_cast<uint32_t*>(&v); }
template <> __device__ __forceinline__ uint32_t pack_h2<__nv_bfloat16>(__half2 v) {
  const float2 f = __half22float2(v);
  __nv_bfloat162 b = __floats2bfloat162_rn(f.x, f.y);
  return *reinterpret_cast<uint32_t*>(&b);
}

// 16-bit storage type helpers: T16 is __nv_bfloat16 (blocks) or __half (stem/head)
template <typename T> struct T16x2;
template <> struct T16x2<__nv_bfloat16> { using type = __nv_bfloat162; };
template <> struct T16x2<__half>        { using type = __half2; };

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }

template <typename T> __device__ __forceinline__ uint32_t pack2(float lo, float hi);
template <> __device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
template <> __device__ __forceinline__ uint32_t pack2<__half>(float lo, float hi) {
  __half2 v = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
template <typename T> __device__ __forceinline__ float2 unpack2(uint32_t u);
template <> __device__ __forceinline__ float2 unpack2<__nv_bfloat16>(uint32_t u) {
  return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u));
}
template <> __device__ __forceinline__ float2 unpack2<__half>(uint32_t u) {
  return __half22float2(*reinterpret_cast<__half2*>(&u));
}

// ---------------------------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must surface as a trapped launch (sticky CUDA error), never as a
// hung GPU.  try_wait suspends in hardware for a bounded time per probe, so ~2^22 probes is seconds.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  for (uint32_t it = 0; it < (1u << 22); ++it)
    if (mbar_try_wait(bar, parity)) return;
  printf("bubbleformer_b200: mbarrier wait timed out (block %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x);
  __trap();
}

// ---------------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor): global -> shared, completion on an mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3)
      : "memory");
}

// TMA stores (shared -> global) of a 2-D box; bulk-group completion.  The writing threads must execute
// fence.proxy.async (fence_proxy_async) and synchronise before one thread issues the store.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem)), "r"(c0), "r"(c1)
               : "memory");
}
// element-wise fp32 add of the box into global memory (split-K accumulation without atomics from registers)
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* m, const void* smem, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* smem, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_4d(const CUtensorMap* m, const void* smem, int c0, int c1, int c2, int c3) {
  asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until at most N of this thread's bulk groups are still READING their shared-memory source
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// tcgen05: TMEM allocation, MMA issue, commit, TMEM loads
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc], kind::f16 (bf16 or fp16 inputs, fp32 accumulate)
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 lanes x 32 consecutive fp32 columns: thread i of the warp gets lane (base_lane + i), v[j] = column j
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, float (&v)[32]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

// ---------------------------------------------------------------------------------------------
// CTA pairs (cta_group::2): the two CTAs of a 2-CTA cluster (one TPC) execute one 256-row MMA together.  Each CTA
// stages its own 128 rows of A and HALF of the B tile; the leader (cluster rank 0) issues the MMA, which reads both
// CTAs' shared memory at the same offsets and writes each CTA's 128 accumulator rows into that CTA's TMEM.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory location in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load issued by either CTA of a pair; completion bytes go to the mbarrier at `bar_cluster_addr` (the leader's)
__device__ __forceinline__ void tma_load_2d_pair(void* smem, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once all previously issued MMAs of this thread have completed) on the barrier at this shared-memory offset in
// every CTA of `cta_mask`
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(cta_mask)
               : "memory");
}

// Shared-memory matrix descriptor (PTX "matrix descriptor", sm_100 version field = 1).
//   bits [0,14)  start address >> 4        bits [16,30) leading-dim byte offset >> 4
//   bits [32,46) stride-dim byte offset >> 4   bits [46,48) version = 1
//   bits [61,64) swizzle mode: 2 = 128-byte swizzle
// K-major, 128B swizzle, rows of 64 16-bit elements (=128 B): 8-row groups are 1024 B apart (SBO);
// LBO is unused.  MN-major, 128B swizzle: 64 MN-elements (128 B) per K row, 8-K groups 1024 B
// apart (SBO), successive 64-element MN blocks LBO bytes apart.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// Instruction descriptor for kind::f16: fp32 accumulate, A/B format (0 = fp16, 1 = bf16), majors, N, M.
// (A and B carry their own format fields, but they must agree: fp16 x bf16 is an illegal instruction on sm_100a.)
__host__ __device__ constexpr uint32_t make_idesc_f16(int fmt, int a_mn_major, int b_mn_major, int M, int N) {
  return (1u << 4) | (static_cast<uint32_t>(fmt) << 7) | (static_cast<uint32_t>(fmt) << 10) |
         (static_cast<uint32_t>(a_mn_major) << 15) | (static_cast<uint32_t>(b_mn_major) << 16) |
         (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}

}  // namespace bf
