// Patch-boundary kernels: the two ends of the hierarchical patch embed / unembed where one side is the
// fp32 NCHW field tensor and the contraction is only 4*fields (= 16) wide -- HBM-bound, so plain SIMT
// fp32 FMAs with coalesced 16-byte accesses rather than a tensor-core tile.
//
//   bf_patch_in   : (I, F, H, W) fp32 NCHW  --2x2/s2 conv-->  (I, H/2, W/2, N) 16-bit channels-last (+ IN sums)
//                   upstream layers/patching.py:37-44 (first Conv2d of HMLPEmbed), and the input-gradient of the
//                   last ConvTranspose2d of HMLPDebed (patching.py:93-99 reversed)
//   bf_patch_out  : (I, h, w, C) 16-bit  --2x2/s2 conv-transpose-->  (I, F, 2h, 2w) fp32 NCHW
//                   upstream layers/patching.py:93-99 (last ConvTranspose2d of HMLPDebed), and the input-gradient of
//                   the first Conv2d of HMLPEmbed
//   bf_patch_wgrad: dW[n][(f,ky,kx)] += sum_pix A[pix][n] * X[img, f, 2y+ky, 2x+kx]   (weight gradient of both)
//   bf_s2d_gather : explicit im2col of 2x2/s2 patches of a channels-last 16-bit image (fallback when the
//                   implicit-GEMM TMA box does not tile the image width, and B operand of the stage wgrads)
#include "common.cuh"

namespace bf {

constexpr int kMaxF = 8;     // fields

// ---------------------------------------------------------------------------------------------
// (I, F, H, W) fp32  ->  (I, H/2, W/2, N) 16-bit,  out[pix][n] = sum_k patch[pix][k] * Wkn[k][n], k = (f, ky, kx)
// Each thread owns 8 output channels and keeps their 4F x 8 weights in registers (F <= 4), so the inner loop
// is pure FMA: 16 coalesced input floats in, one 16-byte store out.
// ---------------------------------------------------------------------------------------------
template <typename T16, int FMAX>
__global__ void __launch_bounds__(256)
patch_in_kernel(const float* __restrict__ x, const float* __restrict__ Wkn, T16* __restrict__ out, float* stats,
                int I, int F, int H, int W, int N, int pix_per_block) {
  pdl_prologue_done();
  extern __shared__ float sm[];
  float* sStat = sm;                 // [N][2]
  for (int i = threadIdx.x; i < 2 * N; i += blockDim.x) sStat[i] = 0.f;
  __syncthreads();
  const int groups = N / 8;
  const int ppi = 256 / groups;                       // pixels in flight per iteration
  const int cg = threadIdx.x % groups, pl = threadIdx.x / groups;
  const int Ho = H / 2, Wo = W / 2;
  const long pix_img = (long)Ho * Wo;
  const int img = blockIdx.y;
  const long p0 = (long)blockIdx.x * pix_per_block;
  const long p1 = min(pix_img, p0 + pix_per_block);
  float wreg[4 * FMAX][8];
#pragma unroll
  for (int k = 0; k < 4 * FMAX; ++k) {
#pragma unroll
    for (int j = 0; j < 8; ++j) wreg[k][j] = (k < 4 * F) ? __ldg(Wkn + (long)k * N + cg * 8 + j) : 0.f;
  }
  float ssum[8], ssq[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { ssum[j] = 0.f; ssq[j] = 0.f; }
  if (pl < ppi) {
    for (long pix = p0 + pl; pix < p1; pix += ppi) {
      const int yo = (int)(pix / Wo), xo = (int)(pix - (long)yo * Wo);
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
      for (int f = 0; f < FMAX; ++f) {
        if (f < F) {
          const float* xp = x + (((long)img * F + f) * H + 2 * yo) * W + 2 * xo;
          const float2 r0 = *reinterpret_cast<const float2*>(xp);
          const float2 r1 = *reinterpret_cast<const float2*>(xp + W);
          const float v[4] = {r0.x, r0.y, r1.x, r1.y};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = fmaf(v[q], wreg[f * 4 + q][j], acc[j]);
          }
        }
      }
      uint4 u;
      u.x = pack2<T16>(acc[0], acc[1]); u.y = pack2<T16>(acc[2], acc[3]);
      u.z = pack2<T16>(acc[4], acc[5]); u.w = pack2<T16>(acc[6], acc[7]);
      *reinterpret_cast<uint4*>(out + ((long)img * pix_img + pix) * N + cg * 8) = u;
      if (stats != nullptr) {
        // statistics of the values as stored (16-bit rounded), so that IN sees exactly its input
        const float2 a = unpack2<T16>(u.x), b = unpack2<T16>(u.y), c = unpack2<T16>(u.z), d = unpack2<T16>(u.w);
        const float r[8] = {a.x, a.y, b.x, b.y, c.x, c.y, d.x, d.y};
#pragma unroll
        for (int j = 0; j < 8; ++j) { ssum[j] += r[j]; ssq[j] = fmaf(r[j], r[j], ssq[j]); }
      }
    }
  }
  if (stats != nullptr) {
    if (pl < ppi) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        atomicAdd(sStat + (cg * 8 + j) * 2, ssum[j]);
        atomicAdd(sStat + (cg * 8 + j) * 2 + 1, ssq[j]);
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * N; i += blockDim.x) atomicAdd(stats + (long)img * N * 2 + i, sStat[i]);
  }
}

// ---------------------------------------------------------------------------------------------
// (I, h, w, C) 16-bit  ->  (I, F, 2h, 2w) fp32,  out[img][f][2y+ky][2x+kx] = sum_c a[pix][c] * Wck[c][(f,ky,kx)]
// One thread per input pixel computes all 4F outputs; the activation row is read from a conflict-free
// (odd word stride) shared tile, the weights are warp-broadcast from shared memory.
// ---------------------------------------------------------------------------------------------
template <typename T16, int FMAX>
__global__ void __launch_bounds__(256)
patch_out_kernel(const T16* __restrict__ a, const float* __restrict__ Wck, float* __restrict__ out,
                 int I, int F, int h, int w, int C, int tile_px) {
  pdl_prologue_done();
  extern __shared__ float sm[];
  const int K = 4 * F;
  float* sW = sm;                                               // [C][4*FMAX] (zero padded)
  uint32_t* sA = reinterpret_cast<uint32_t*>(sW + C * 4 * FMAX);  // [tile_px][C/2 + 1] channel pairs
  const int CW = C / 2 + 1;
  for (int i = threadIdx.x; i < C * 4 * FMAX; i += blockDim.x) {
    const int c = i / (4 * FMAX), k = i - c * (4 * FMAX);
    sW[i] = k < K ? Wck[(long)c * K + k] : 0.f;
  }
  const long pix_img = (long)h * w;
  const int img = blockIdx.y;
  const long p0 = (long)blockIdx.x * tile_px;
  const int npx = (int)min((long)tile_px, pix_img - p0);
  const uint32_t* src = reinterpret_cast<const uint32_t*>(a + ((long)img * pix_img + p0) * C);
  const int wpp = C / 2;                                        // 32-bit words per pixel
  for (int i = threadIdx.x; i < npx * wpp; i += blockDim.x) {
    const int px = i / wpp, cw = i - px * wpp;
    sA[px * CW + cw] = src[i];
  }
  __syncthreads();
  const int px = threadIdx.x;
  if (px < npx) {
    float acc[4 * FMAX];
#pragma unroll
    for (int k = 0; k < 4 * FMAX; ++k) acc[k] = 0.f;
    const uint32_t* ar = sA + px * CW;
    for (int cw = 0; cw < wpp; ++cw) {
      const float2 av = unpack2<T16>(ar[cw]);
      const float* w0 = sW + (2 * cw) * 4 * FMAX;
      const float* w1 = w0 + 4 * FMAX;
#pragma unroll
      for (int k4 = 0; k4 < FMAX; ++k4) {
        const float4 u0 = *reinterpret_cast<const float4*>(w0 + 4 * k4);
        const float4 u1 = *reinterpret_cast<const float4*>(w1 + 4 * k4);
        acc[4 * k4 + 0] = fmaf(av.x, u0.x, fmaf(av.y, u1.x, acc[4 * k4 + 0]));
        acc[4 * k4 + 1] = fmaf(av.x, u0.y, fmaf(av.y, u1.y, acc[4 * k4 + 1]));
        acc[4 * k4 + 2] = fmaf(av.x, u0.z, fmaf(av.y, u1.z, acc[4 * k4 + 2]));
        acc[4 * k4 + 3] = fmaf(av.x, u0.w, fmaf(av.y, u1.w, acc[4 * k4 + 3]));
      }
    }
    const long pix = p0 + px;
    const int y = (int)(pix / w), xx = (int)(pix - (long)y * w);
#pragma unroll
    for (int f = 0; f < FMAX; ++f) {
      if (f < F) {
        float* op = out + (((long)img * F + f) * (2 * h) + 2 * y) * (2 * w) + 2 * xx;
        *reinterpret_cast<float2*>(op) = make_float2(acc[4 * f + 0], acc[4 * f + 1]);
        *reinterpret_cast<float2*>(op + 2 * w) = make_float2(acc[4 * f + 2], acc[4 * f + 3]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// dW[n][(f,ky,kx)] += sum_pix A[pix][n] * X[img][f][2y+ky][2x+kx]
// ---------------------------------------------------------------------------------------------
template <typename T16>
__global__ void __launch_bounds__(1024)
patch_wgrad_kernel(const T16* __restrict__ a, const float* __restrict__ x, float* __restrict__ dW,
                   int I, int F, int H, int W, int N, int lda, int pix_per_block) {
  pdl_prologue_done();
  extern __shared__ float sm[];
  constexpr int TP = 32;                                 // pixels per staging tile
  const int K = 4 * F;
  float* sX = sm;                                        // [TP][K]
  T16* sA = reinterpret_cast<T16*>(sX + TP * K);         // [TP][N]
  const int Ho = H / 2, Wo = W / 2;
  const long pix_img = (long)Ho * Wo;
  const int img = blockIdx.y;
  const long p0 = (long)blockIdx.x * pix_per_block;
  const long p1 = min(pix_img, p0 + pix_per_block);
  const int n = threadIdx.x % N, f = threadIdx.x / N;     // blockDim.x == N * F
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (long t0 = p0; t0 < p1; t0 += TP) {
    const int np = (int)min((long)TP, p1 - t0);
    __syncthreads();
    for (int i = threadIdx.x; i < np * (N / 8); i += blockDim.x) {
      const int px = i / (N / 8), ch = i - px * (N / 8);
      *reinterpret_cast<uint4*>(sA + px * N + ch * 8) =
          *reinterpret_cast<const uint4*>(a + ((long)img * pix_img + t0 + px) * lda + ch * 8);
    }
    for (int i = threadIdx.x; i < np * F * 2; i += blockDim.x) {
      const int px = i / (F * 2), r = i - px * (F * 2);
      const int ff = r >> 1, ky = r & 1;
      const long pix = t0 + px;
      const int yo = (int)(pix / Wo), xo = (int)(pix - (long)yo * Wo);
      const float2 v = *reinterpret_cast<const float2*>(x + (((long)img * F + ff) * H + 2 * yo + ky) * W + 2 * xo);
      sX[px * K + ff * 4 + ky * 2] = v.x;
      sX[px * K + ff * 4 + ky * 2 + 1] = v.y;
    }
    __syncthreads();
    for (int px = 0; px < np; ++px) {
      const float av = to_f32<T16>(sA[px * N + n]);
      const float4 xv = *reinterpret_cast<const float4*>(sX + px * K + f * 4);
      acc[0] = fmaf(av, xv.x, acc[0]); acc[1] = fmaf(av, xv.y, acc[1]);
      acc[2] = fmaf(av, xv.z, acc[2]); acc[3] = fmaf(av, xv.w, acc[3]);
    }
  }
  float* d = dW + (long)n * K + f * 4;
#pragma unroll
  for (int j = 0; j < 4; ++j) atomicAdd(d + j, acc[j]);
}

// ---------------------------------------------------------------------------------------------
// explicit im2col: (I, Hin, Win, C) 16-bit -> (I*Hin/2*Win/2, 4C), K order (ky, kx, ci)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 half8_to_bf16(uint4 v) {
  uint32_t* w = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
  for (int k = 0; k < 4; ++k) { const float2 f = unpack2<__half>(w[k]); w[k] = pack2<__nv_bfloat16>(f.x, f.y); }
  return v;
}
__device__ __forceinline__ uint4 bf16_to_half8(uint4 v) {
  uint32_t* w = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
  for (int k = 0; k < 4; ++k) { const float2 f = unpack2<__nv_bfloat16>(w[k]); w[k] = pack2<__half>(f.x, f.y); }
  return v;
}

// conv: 0 none, 1 fp16 -> bf16, 2 bf16 -> fp16
__global__ void __launch_bounds__(256)
s2d_gather_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, long M, int Ho, int Wo, int C8, int conv) {
  pdl_prologue_done();
  // one 16-byte chunk per thread; a row of the output is 4*C8 chunks = two runs of 2*C8 chunks
  const long total = M * 4 * C8;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long m = i / (4 * C8);
    const int r = (int)(i - m * 4 * C8);
    const int ky = r / (2 * C8), j = r - ky * 2 * C8;
    const long xo = m % Wo, t = m / Wo;
    const long yo = t % Ho, img = t / Ho;
    const long src = (((img * Ho + yo) * 2 + ky) * (2L * Wo) + 2 * xo) * C8 + j;
    uint4 v = in[src];
    if (conv == 1) v = half8_to_bf16(v);
    else if (conv == 2) v = bf16_to_half8(v);
    out[i] = v;
  }
}

__global__ void __launch_bounds__(256)
convert16_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, long n8, int conv) {
  pdl_prologue_done();
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long)gridDim.x * blockDim.x) {
    uint4 v = in[i];
    out[i] = conv == 1 ? half8_to_bf16(v) : bf16_to_half8(v);
  }
}

// flat fp32 -> 16-bit cast (weights), 8 elements per thread
template <typename T16>
__global__ void __launch_bounds__(256)
cast16_kernel(const float* __restrict__ in, T16* __restrict__ out, long n) {
  pdl_prologue_done();
  const long n8 = n / 8;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long)gridDim.x * blockDim.x) {
    const float4 a = reinterpret_cast<const float4*>(in)[2 * i], b = reinterpret_cast<const float4*>(in)[2 * i + 1];
    uint4 u;
    u.x = pack2<T16>(a.x, a.y); u.y = pack2<T16>(a.z, a.w); u.z = pack2<T16>(b.x, b.y); u.w = pack2<T16>(b.z, b.w);
    reinterpret_cast<uint4*>(out)[i] = u;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n - n8 * 8)) out[n8 * 8 + threadIdx.x] = from_f32<T16>(in[n8 * 8 + threadIdx.x]);
}

}  // namespace bf

using namespace bf;

namespace bf {
bool patch_wgrad_mma_ok(int F, int W, int N);
bool patch_out_mma_ok(int F, int C);
int launch_patch_wgrad_mma(const void* a, int dtype, const float* x, float* dW, int I, int F, int H, int W, int N,
                           cudaStream_t s);
int launch_patch_out_mma(const void* a, int dtype, const float* Wck, float* out, int I, int F, int h, int w, int C,
                         cudaStream_t s);
}
namespace bf {
int launch_patch_in_f32(const float* x, const float* Wkn, float* out, float* stats, int I, int F, int H, int W, int N, cudaStream_t st);
int launch_patch_out_f32(const float* a, const float* Wck, float* out, int I, int F, int h, int w, int C, cudaStream_t st);
int launch_patch_wgrad_f32(const float* a, const float* x, float* dW, int I, int F, int H, int W, int N, cudaStream_t st);
int launch_s2d_gather_f32(const float* in, float* out, int I, int Hin, int Win, int C, cudaStream_t st);
}
namespace bf { int launch_patch_in_mma(const float* x, const float* Wkn, void* out, int dtype, float* stats, int I, int F,
                                       int H, int W, int N, cudaStream_t s); }

extern "C" int bf_patch_in(const float* x, const float* Wkn, void* out, int dtype, float* stats, int I, int F, int H,
                           int W, int N, void* stream) {
  BF_REQUIRE(x && Wkn && out, "bf_patch_in: null pointer");
  if (dtype == BF_F32) {                      // fp32 validation backend
    BF_REQUIRE(I > 0 && F > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0 && N > 0, "bf_patch_in (fp32): geometry");
    return launch_patch_in_f32(x, Wkn, static_cast<float*>(out), stats, I, F, H, W, N, static_cast<cudaStream_t>(stream));
  }
  BF_REQUIRE(dtype == BF_BF16 || dtype == BF_F16, "bf_patch_in: dtype");
  BF_REQUIRE(I > 0 && F > 0 && F <= kMaxF && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0, "bf_patch_in: geometry");
  BF_REQUIRE(N % 8 == 0 && N >= 8 && N <= 2048, "bf_patch_in: N=%d must be a multiple of 8 in [8, 2048]", N);
  BF_REQUIRE((reinterpret_cast<uintptr_t>(x) & 7) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0, "bf_patch_in: alignment");
  if (F <= 8 && I <= 65535) return launch_patch_in_mma(x, Wkn, out, dtype, stats, I, F, H, W, N, static_cast<cudaStream_t>(stream));
  const long pix_img = (long)(H / 2) * (W / 2);
  const int ppi = 256 / (N / 8);
  long ppb = (long)ppi * 64;                         // 64 pixels per thread amortise the register-resident weights
  while (ppb > ppi && ((pix_img + ppb - 1) / ppb) * I < 2L * num_sms()) ppb /= 2;
  if (ppb < ppi) ppb = ppi;
  dim3 grid((unsigned)((pix_img + ppb - 1) / ppb), I);
  const size_t sm = (size_t)2 * N * sizeof(float);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
#define BF_PIN(T, FM) launch_k(patch_in_kernel<T, FM>, dim3(grid), dim3(256), (size_t)(sm), s, x, Wkn, (T*)out, stats, I, F, H, W, N, (int)ppb)
  if (dtype == BF_BF16) { if (F <= 4) BF_PIN(__nv_bfloat16, 4); else BF_PIN(__nv_bfloat16, 8); }
  else                  { if (F <= 4) BF_PIN(__half, 4); else BF_PIN(__half, 8); }
#undef BF_PIN
  count_launch();
  BF_LAUNCH_CHECK("patch_in_kernel");
  return BF_OK;
}

template <typename T16, int FM>
static int launch_patch_out(const void* a, const float* Wck, float* out, int I, int F, int h, int w, int C, cudaStream_t s) {
  const long pix_img = (long)h * w;
  int tile_px = pix_img < 256 ? (int)pix_img : 256;
  const size_t sm = (size_t)C * 4 * FM * sizeof(float) + (size_t)tile_px * (C / 2 + 1) * 4;
  BF_REQUIRE(sm <= 220 * 1024, "bf_patch_out: C=%d too large for the shared-memory tile", C);
  static bool done = false;
  if (!done) {
    if (int e = check_cuda(cudaFuncSetAttribute(patch_out_kernel<T16, FM>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                220 * 1024), "cudaFuncSetAttribute(patch_out)")) return e;
    done = true;
  }
  dim3 grid((unsigned)((pix_img + tile_px - 1) / tile_px), I);
  launch_k(patch_out_kernel<T16, FM>, dim3(grid), dim3(256), (size_t)(sm), s, (const T16*)a, Wck, out, I, F, h, w, C, tile_px);
  count_launch();
  BF_LAUNCH_CHECK("patch_out_kernel");
  return BF_OK;
}

extern "C" int bf_patch_out(const void* a, int dtype, const float* Wck, float* out, int I, int F, int h, int w, int C,
                            void* stream) {
  BF_REQUIRE(a && Wck && out, "bf_patch_out: null pointer");
  if (dtype == BF_F32) {
    BF_REQUIRE(I > 0 && F > 0 && h > 0 && w > 0 && C > 0, "bf_patch_out (fp32): geometry");
    return launch_patch_out_f32(static_cast<const float*>(a), Wck, out, I, F, h, w, C, static_cast<cudaStream_t>(stream));
  }
  BF_REQUIRE(dtype == BF_BF16 || dtype == BF_F16, "bf_patch_out: dtype");
  BF_REQUIRE(I > 0 && F > 0 && F <= kMaxF && h > 0 && w > 0 && C % 8 == 0 && C > 0, "bf_patch_out: geometry");
  BF_REQUIRE((reinterpret_cast<uintptr_t>(a) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 7) == 0, "bf_patch_out: alignment");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (patch_out_mma_ok(F, C) && I <= 65535) return launch_patch_out_mma(a, dtype, Wck, out, I, F, h, w, C, s);
  if (dtype == BF_BF16) return F <= 4 ? launch_patch_out<__nv_bfloat16, 4>(a, Wck, out, I, F, h, w, C, s)
                                      : launch_patch_out<__nv_bfloat16, 8>(a, Wck, out, I, F, h, w, C, s);
  return F <= 4 ? launch_patch_out<__half, 4>(a, Wck, out, I, F, h, w, C, s)
                : launch_patch_out<__half, 8>(a, Wck, out, I, F, h, w, C, s);
}

extern "C" int bf_patch_wgrad(const void* a, int dtype, const float* x, float* dW, int I, int F, int H, int W, int N,
                              void* stream) {
  BF_REQUIRE(a && x && dW, "bf_patch_wgrad: null pointer");
  if (dtype == BF_F32) {
    BF_REQUIRE(I > 0 && F > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0 && N > 0 && 4 * F <= 65535, "bf_patch_wgrad (fp32): geometry");
    return launch_patch_wgrad_f32(static_cast<const float*>(a), x, dW, I, F, H, W, N, static_cast<cudaStream_t>(stream));
  }
  BF_REQUIRE(dtype == BF_BF16 || dtype == BF_F16, "bf_patch_wgrad: dtype");
  BF_REQUIRE(I > 0 && F > 0 && F <= kMaxF && H % 2 == 0 && W % 2 == 0 && H > 0 && W > 0, "bf_patch_wgrad: geometry");
  BF_REQUIRE(N % 8 == 0, "bf_patch_wgrad: N=%d must be a multiple of 8", N);
  if (patch_wgrad_mma_ok(F, W, N) && I <= 65535 && (reinterpret_cast<uintptr_t>(a) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(x) & 15) == 0)
    return launch_patch_wgrad_mma(a, dtype, x, dW, I, F, H, W, N, static_cast<cudaStream_t>(stream));
  const long pix_img = (long)(H / 2) * (W / 2);
  long ppb = 1024;
  // enough blocks to fill the machine, few enough that the N*4F atomics per block stay cheap
  while (ppb > 32 && (pix_img + ppb - 1) / ppb * I < 2L * num_sms()) ppb /= 2;
  dim3 grid((unsigned)((pix_img + ppb - 1) / ppb), I);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // column chunks of at most 1024 / F channels (one thread per (channel, field))
  int nc_max = (1024 / F) / 8 * 8;
  for (int n0 = 0; n0 < N; n0 += nc_max) {
    const int nc = N - n0 < nc_max ? N - n0 : nc_max;
    const size_t sm = (size_t)32 * 4 * F * sizeof(float) + (size_t)32 * nc * 2;
    float* dWc = dW + (long)n0 * 4 * F;
    if (dtype == BF_BF16)
      launch_k(patch_wgrad_kernel<__nv_bfloat16>, dim3(grid), dim3(nc * F), (size_t)(sm), s, (const __nv_bfloat16*)a + n0, x, dWc, I, F, H, W, nc, N, (int)ppb);
    else
      launch_k(patch_wgrad_kernel<__half>, dim3(grid), dim3(nc * F), (size_t)(sm), s, (const __half*)a + n0, x, dWc, I, F, H, W, nc, N, (int)ppb);
    count_launch();
    BF_LAUNCH_CHECK("patch_wgrad_kernel");
  }
  return BF_OK;
}

static int conv_code(int in_dtype, int out_dtype) {
  if (in_dtype == out_dtype) return 0;
  return in_dtype == BF_F16 ? 1 : 2;
}

extern "C" int bf_convert16(const void* in, int in_dtype, void* out, int out_dtype, int64_t n, void* stream) {
  BF_REQUIRE(in && out && n > 0 && n % 8 == 0, "bf_convert16: n must be a positive multiple of 8");
  BF_REQUIRE((in_dtype == BF_F16 && out_dtype == BF_BF16) || (in_dtype == BF_BF16 && out_dtype == BF_F16),
             "bf_convert16: dtype pair");
  long blocks = (n / 8 + 255) / 256;
  if (blocks > 8L * num_sms()) blocks = 8L * num_sms();
  launch_k(convert16_kernel, dim3((unsigned)blocks), dim3(256), (size_t)(0), static_cast<cudaStream_t>(stream), 
      (const uint4*)in, (uint4*)out, n / 8, conv_code(in_dtype, out_dtype));
  count_launch();
  BF_LAUNCH_CHECK("convert16_kernel");
  return BF_OK;
}

extern "C" int bf_s2d_gather(const void* in, int in_dtype, void* out, int out_dtype, int I, int Hin, int Win, int C,
                             void* stream) {
  BF_REQUIRE(in && out, "bf_s2d_gather: null pointer");
  if (in_dtype == BF_F32 && out_dtype == BF_F32) {
    BF_REQUIRE(I > 0 && Hin > 0 && Win > 0 && Hin % 2 == 0 && Win % 2 == 0 && C > 0, "bf_s2d_gather (fp32): geometry");
    return launch_s2d_gather_f32(static_cast<const float*>(in), static_cast<float*>(out), I, Hin, Win, C,
                                 static_cast<cudaStream_t>(stream));
  }
  BF_REQUIRE((in_dtype == BF_F16 || in_dtype == BF_BF16) && (out_dtype == BF_F16 || out_dtype == BF_BF16),
             "bf_s2d_gather: dtypes");
  BF_REQUIRE(I > 0 && Hin > 0 && Win > 0 && Hin % 2 == 0 && Win % 2 == 0 && C > 0 && C % 8 == 0,
             "bf_s2d_gather: geometry (C must be a multiple of 8)");
  const long M = (long)I * (Hin / 2) * (Win / 2);
  const long total = M * 4 * (C / 8);
  long blocks = (total + 255) / 256;
  if (blocks > 8L * num_sms()) blocks = 8L * num_sms();
  launch_k(s2d_gather_kernel, dim3((unsigned)blocks), dim3(256), (size_t)(0), static_cast<cudaStream_t>(stream), 
      (const uint4*)in, (uint4*)out, M, Hin / 2, Win / 2, C / 8, conv_code(in_dtype, out_dtype));
  count_launch();
  BF_LAUNCH_CHECK("s2d_gather_kernel");
  return BF_OK;
}

extern "C" int bf_cast16(const float* in, void* out, int dtype, int64_t n, void* stream) {
  BF_REQUIRE(in && out && n > 0, "bf_cast16: bad arguments");
  BF_REQUIRE(dtype == BF_BF16 || dtype == BF_F16, "bf_cast16: dtype");
  BF_REQUIRE((reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0, "bf_cast16: alignment");
  long blocks = (n / 8 + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > 8L * num_sms()) blocks = 8L * num_sms();
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dtype == BF_BF16) launch_k(cast16_kernel<__nv_bfloat16>, dim3((unsigned)blocks), dim3(256), (size_t)(0), s, in, (__nv_bfloat16*)out, n);
  else launch_k(cast16_kernel<__half>, dim3((unsigned)blocks), dim3(256), (size_t)(0), s, in, (__half*)out, n);
  count_launch();
  BF_LAUNCH_CHECK("cast16_kernel");
  return BF_OK;
}
