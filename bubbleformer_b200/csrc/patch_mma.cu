// bf_patch_in on tensor cores: (I, F, H, W) fp32 NCHW  --2x2/s2 conv-->  (I, H/2, W/2, N) 16-bit channels-last,
// optionally with the InstanceNorm sums of the result.  First Conv2d of HMLPEmbed (upstream layers/patching.py:37-44)
// and the input gradient of the last ConvTranspose2d of HMLPDebed (patching.py:93-99 reversed).
//
// The contraction is only K = 4F (= 16) deep, so the kernel must live at the HBM roofline (read 4 B x 16 per pixel,
// write 2 B x N): one warp owns 16 consecutive output pixels of one image row, gathers their 2x2 patches straight
// into TF32 mma.sync A fragments (each lane loads exactly the scalars its fragment needs; a warp instruction covers
// two 64-byte runs of one field), multiplies by weights held in padded shared memory (conflict-free B fragments),
// and stages the 16 x N tile through shared memory so the result leaves as whole 16-byte coalesced stores.
// TF32 keeps the 10-bit operand mantissa the stem needs (its error is amplified by the InstanceNorm that follows).
#include "common.cuh"

namespace bf {

constexpr int kPinWarps = 8;
constexpr int kPinNT = 12;          // n tiles (of 8 channels) per pass: 96 channels

__device__ __forceinline__ uint32_t to_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <typename T16, bool STATS, int KS>
__global__ void __launch_bounds__(kPinWarps * 32, 2)
patch_in_mma_kernel(const float* __restrict__ x, const float* __restrict__ Wkn, T16* __restrict__ out, float* stats,
                    int F, int H, int W, int N, int tiles_per_block) {
  pdl_prologue_done();
  extern __shared__ __align__(16) uint8_t smem[];
  const int K = 4 * F;
  const int WS = N + 8;                                   // padded weight row (words)
  uint32_t* sW = reinterpret_cast<uint32_t*>(smem);       // [KS*8][WS] tf32
  const int nch = N < 8 * kPinNT ? N : 8 * kPinNT;        // channels per pass
  const int SS = nch + 8;                                 // padded stage row (elements)
  T16* sStage = reinterpret_cast<T16*>(sW + KS * 8 * WS) + (size_t)(threadIdx.x >> 5) * 16 * SS;
  float* sStat = reinterpret_cast<float*>(reinterpret_cast<T16*>(sW + KS * 8 * WS) + (size_t)kPinWarps * 16 * SS);   // [N][2]
  for (int i = threadIdx.x; i < KS * 8 * WS; i += blockDim.x) {
    const int k = i / WS, n = i - k * WS;
    sW[i] = (k < K && n < N) ? to_tf32(Wkn[(long)k * N + n]) : 0u;
  }
  if (STATS) for (int i = threadIdx.x; i < 2 * N; i += blockDim.x) sStat[i] = 0.f;
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int Ho = H / 2, Wo = W / 2;
  const int tiles_x = (Wo + 15) / 16;
  const int n_tiles = Ho * tiles_x;
  const int img = blockIdx.y;
  const int tile0 = blockIdx.x * tiles_per_block;
  const int tile1 = min(n_tiles, tile0 + tiles_per_block);
  const float* ximg = x + (long)img * F * H * W;
  T16* oimg = out + (long)img * Ho * Wo * N;
  const int ky = t >> 1, kx = t & 1;

  for (int nc0 = 0; nc0 < N; nc0 += 8 * kPinNT) {
    const int nts = min(kPinNT, (N - nc0) / 8);
    float ssum[kPinNT][2], ssq[kPinNT][2];
#pragma unroll
    for (int nt = 0; nt < kPinNT; ++nt) { ssum[nt][0] = ssum[nt][1] = ssq[nt][0] = ssq[nt][1] = 0.f; }
    // A fragments: a0 (row g, k = t), a1 (row g+8, k = t), a2 (row g, k = t+4), a3 (row g+8, k = t+4); k = f*4 + ky*2 + kx.
    // The gather of the next tile is issued before the current tile is multiplied and stored (its latency hides there).
    auto gather = [&](int tile, float (&r)[KS][4]) {
      const int yo = tile / tiles_x, xo0 = (tile - yo * tiles_x) * 16;
      const bool ok0 = xo0 + g < Wo, ok1 = xo0 + g + 8 < Wo;
#pragma unroll
      for (int s = 0; s < KS; ++s) {
        const int f0 = 2 * s, f1 = 2 * s + 1;
        const float* p0 = ximg + ((long)f0 * H + 2 * yo + ky) * W + 2 * (xo0 + g) + kx;
        const float* p1 = ximg + ((long)f1 * H + 2 * yo + ky) * W + 2 * (xo0 + g) + kx;
        r[s][0] = (f0 < F && ok0) ? __ldg(p0) : 0.f;
        r[s][1] = (f0 < F && ok1) ? __ldg(p0 + 16) : 0.f;
        r[s][2] = (f1 < F && ok0) ? __ldg(p1) : 0.f;
        r[s][3] = (f1 < F && ok1) ? __ldg(p1 + 16) : 0.f;
      }
    };
    float raw[KS][4];
    if (tile0 + warp < tile1) gather(tile0 + warp, raw);
    for (int tile = tile0 + warp; tile < tile1; tile += kPinWarps) {
      const int yo = tile / tiles_x, xo0 = (tile - yo * tiles_x) * 16;
      const bool ok0 = xo0 + g < Wo, ok1 = xo0 + g + 8 < Wo;
      uint32_t a[KS][4];
#pragma unroll
      for (int s = 0; s < KS; ++s)
#pragma unroll
        for (int q = 0; q < 4; ++q) a[s][q] = to_tf32(raw[s][q]);
      if (tile + kPinWarps < tile1) gather(tile + kPinWarps, raw);
      float c[kPinNT][4];
#pragma unroll
      for (int nt = 0; nt < kPinNT; ++nt) { c[nt][0] = c[nt][1] = c[nt][2] = c[nt][3] = 0.f; }
#pragma unroll
      for (int s = 0; s < KS; ++s) {
        {
#pragma unroll
          for (int nt = 0; nt < kPinNT; ++nt) {
            if (nt < nts) {
              const uint32_t b0 = sW[(8 * s + t) * WS + nc0 + nt * 8 + g];
              const uint32_t b1 = sW[(8 * s + 4 + t) * WS + nc0 + nt * 8 + g];
              mma_tf32(c[nt], a[s], b0, b1);
            }
          }
        }
      }
      __syncwarp();                          // previous tile's readers are done with the stage
#pragma unroll
      for (int nt = 0; nt < kPinNT; ++nt) {
        if (nt < nts) {
          // statistics of the values as stored (rounded), so the following norm sees exactly its own input
          const uint32_t lo = pack2<T16>(c[nt][0], c[nt][1]), hi = pack2<T16>(c[nt][2], c[nt][3]);
          *reinterpret_cast<uint32_t*>(sStage + g * SS + nt * 8 + 2 * t) = lo;
          *reinterpret_cast<uint32_t*>(sStage + (g + 8) * SS + nt * 8 + 2 * t) = hi;
          if (STATS) {
            const float2 r0 = unpack2<T16>(lo), r1 = unpack2<T16>(hi);
            if (ok0) { ssum[nt][0] += r0.x; ssum[nt][1] += r0.y; ssq[nt][0] = fmaf(r0.x, r0.x, ssq[nt][0]); ssq[nt][1] = fmaf(r0.y, r0.y, ssq[nt][1]); }
            if (ok1) { ssum[nt][0] += r1.x; ssum[nt][1] += r1.y; ssq[nt][0] = fmaf(r1.x, r1.x, ssq[nt][0]); ssq[nt][1] = fmaf(r1.y, r1.y, ssq[nt][1]); }
          }
        }
      }
      __syncwarp();
      const int cpr = nts;                   // 16-byte chunks per pixel in this pass
      const int rows = min(16, Wo - xo0);
      T16* dst = oimg + ((long)yo * Wo + xo0) * N + nc0;
      // idx = lane + 32 k -> (row, chunk) without a division per store (cpr is a run-time value)
      int r = lane / cpr, ch = lane - r * cpr;
      const int dr = 32 / cpr, dc = 32 - dr * cpr;
      for (int idx = lane; idx < rows * cpr; idx += 32) {
        *reinterpret_cast<uint4*>(dst + (long)r * N + ch * 8) = *reinterpret_cast<const uint4*>(sStage + r * SS + ch * 8);
        r += dr; ch += dc;
        if (ch >= cpr) { ch -= cpr; ++r; }
      }
    }
    if (STATS) {
#pragma unroll
      for (int nt = 0; nt < kPinNT; ++nt) {
        if (nt < nts) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            float s = ssum[nt][e], q = ssq[nt][e];
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) {
              s += __shfl_xor_sync(0xffffffffu, s, o);
              q += __shfl_xor_sync(0xffffffffu, q, o);
            }
            if (g == 0) {
              atomicAdd(sStat + 2 * (nc0 + nt * 8 + 2 * t + e), s);
              atomicAdd(sStat + 2 * (nc0 + nt * 8 + 2 * t + e) + 1, q);
            }
          }
        }
      }
    }
  }
  if (STATS) {
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * N; i += blockDim.x) atomicAdd(stats + (long)img * N * 2 + i, sStat[i]);
  }
}

int launch_patch_in_mma(const float* x, const float* Wkn, void* out, int dtype, float* stats, int I, int F, int H, int W,
                        int N, cudaStream_t s) {
  const int Ho = H / 2, Wo = W / 2;
  const int tiles = Ho * ((Wo + 15) / 16);
  // ~4 blocks per SM in total, at least one tile per warp
  int bpi = (4 * num_sms() + I - 1) / I;
  if (bpi > (tiles + kPinWarps - 1) / kPinWarps) bpi = (tiles + kPinWarps - 1) / kPinWarps;
  if (bpi < 1) bpi = 1;
  const int tpb = (tiles + bpi - 1) / bpi;
  bpi = (tiles + tpb - 1) / tpb;
  int KS = (4 * F + 7) / 8;
  if (KS == 3) KS = 4;
  const int nch = N < 8 * kPinNT ? N : 8 * kPinNT;
  const size_t sm = (size_t)KS * 8 * (N + 8) * 4 + (size_t)kPinWarps * 16 * (nch + 8) * 2 + (size_t)2 * N * 4;
  BF_REQUIRE(sm <= 200 * 1024, "bf_patch_in: N=%d F=%d need %zu bytes of shared memory", N, F, sm);
  dim3 grid(bpi, I);
#define BF_PIN_(T, ST, KS_)                                                                                              \
  do {                                                                                                             \
    static bool done_ = false;                                                                                     \
    if (!done_) {                                                                                                  \
      if (int e_ = check_cuda(cudaFuncSetAttribute(patch_in_mma_kernel<T, ST, KS_>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                                   200 * 1024), "cudaFuncSetAttribute(patch_in)")) return e_;      \
      done_ = true;                                                                                                \
    }                                                                                                              \
    launch_k(patch_in_mma_kernel<T, ST, KS_>, dim3(grid), dim3(kPinWarps * 32), (size_t)(sm), s, x, Wkn, (T*)out, stats, F, H, W, N, tpb);          \
  } while (0)
#define BF_PIN(T, ST) do { if (KS == 1) BF_PIN_(T, ST, 1); else if (KS == 2) BF_PIN_(T, ST, 2); else BF_PIN_(T, ST, 4); } while (0)
  if (dtype == BF_BF16) { if (stats) BF_PIN(__nv_bfloat16, true); else BF_PIN(__nv_bfloat16, false); }
  else                  { if (stats) BF_PIN(__half, true); else BF_PIN(__half, false); }
#undef BF_PIN
#undef BF_PIN_
  count_launch();
  BF_LAUNCH_CHECK("patch_in_mma_kernel");
  return BF_OK;
}

}  // namespace bf
