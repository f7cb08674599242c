// bf_patch_wgrad and bf_patch_out as warp-tile streaming kernels on bf16/fp16 mma.sync.
//
// Both ends of the patch embed / unembed touch the full-resolution stage tensor (I * H/2 * W/2 pixels x N channels,
// 503 MB at config 2) against a contraction that is only 4*fields (= 16) wide, so they must run at the HBM
// roofline: read the 16-bit channels-last tensor once, read or write the fp32 NCHW fields once.
//   bf_patch_wgrad: dW[n][(f,ky,kx)] += sum_pix a[pix][n] * x[img, f, 2y+ky, 2x+kx]
//                   (weight gradient of the first Conv2d of HMLPEmbed, upstream layers/patching.py:37-44, and of the
//                   last ConvTranspose2d of HMLPDebed, patching.py:93-99)
//   bf_patch_out  : out[img, f, 2y+ky, 2x+kx] = sum_c a[pix][c] * Wck[c][(f,ky,kx)]
//                   (last ConvTranspose2d of HMLPDebed; input gradient of the first Conv2d of HMLPEmbed)
// One warp owns a tile of 16 consecutive pixels of one image row.  Its operands arrive with 16-byte cp.async into a
// private double buffer (the next tile is in flight while the current one is multiplied), the channels-last rows
// become mma fragments through ldmatrix, and nothing but the per-warp __syncwarp synchronises the loop.
#include <type_traits>

#include "common.cuh"

namespace bf {

constexpr int kPsWarps = 8;
constexpr int kPsMT = 6;             // 16-channel m tiles per pass: 96 channels
constexpr int kPsSA = 16 * kPsMT + 8;  // padded channel row of the activation tile (elements): 208 B, odd multiple of 16 B
constexpr int kPsXRow = 160;         // bytes per (field, ky) row of the fp32 patch tile: 32 floats + padding

__device__ __forceinline__ void cp_async16_zfill(void* smem, const void* gmem, bool valid) {
  const int n = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem)), "l"(gmem), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
template <typename T16>
__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1);
template <>
__device__ __forceinline__ void mma_16816<__nv_bfloat16>(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <>
__device__ __forceinline__ void mma_16816<__half>(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// activation tile: 16 pixels x nc channels (16-bit, channels-last, row pitch lda) -> sa[px][kPsSA]; pixels past the
// row end are zero filled
template <typename T16>
__device__ __forceinline__ void issue_act_tile(uint8_t* sa, const T16* row0, long lda, int nc, int valid_px, int lane) {
  const int cpp = nc >> 3;                      // 16-byte chunks per pixel
  // idx = lane + 32 k -> (pixel, chunk) without a division per copy (cpp is a run-time value)
  int px = lane / cpp, ch = lane - px * cpp;
  const int dp = 32 / cpp, dc = 32 - dp * cpp;
  for (int idx = lane; idx < 16 * cpp; idx += 32) {
    const bool ok = px < valid_px;
    cp_async16_zfill(sa + (px * kPsSA + ch * 8) * 2, ok ? row0 + (long)px * lda + ch * 8 : row0, ok);
    px += dp; ch += dc;
    if (ch >= cpp) { ch -= cpp; ++px; }
  }
}

// ---------------------------------------------------------------------------------------------
// weight gradient.  A operand = a^T (channels x pixels, from sa through ldmatrix.trans), B operand = the 2x2 patches
// (pixels x (f,ky,kx)) gathered from the fp32 rows as bf16 hi + lo (gradients need the bf16 range; fp16 activations are
// rounded to bf16 for the same reason, the one 2^-9 rounding of this kernel); fp32 accumulators
// live in registers for the whole block and are reduced through shared memory into one atomic per element per block.
// ---------------------------------------------------------------------------------------------
template <typename T16, int NT>
__global__ void __launch_bounds__(kPsWarps * 32, 2)
patch_wgrad_mma_kernel(const T16* __restrict__ a, long lda, const float* __restrict__ x, float* __restrict__ dW, int ldw,
                       int F, int H, int W, int nc, int tiles_per_block) {
  pdl_prologue_done();
  extern __shared__ __align__(16) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int K = 4 * F, rows_x = 2 * F;
  const int a_bytes = 16 * kPsSA * 2, buf_bytes = a_bytes + rows_x * kPsXRow;
  uint8_t* my = smem + (size_t)warp * 2 * buf_bytes;
  const int Ho = H / 2, Wo = W / 2;
  const int tiles_x = (Wo + 15) / 16, n_tiles = Ho * tiles_x;
  const int img = blockIdx.y;
  const int tile0 = blockIdx.x * tiles_per_block, tile1 = min(n_tiles, tile0 + tiles_per_block);
  const int mts = nc >> 4;
  const T16* aimg = a + (long)img * Ho * Wo * lda;
  const float* ximg = x + (long)img * F * H * W;

  auto issue = [&](int tile, int buf) {
    const int yo = tile / tiles_x, xo0 = (tile - yo * tiles_x) * 16;
    uint8_t* sa = my + buf * buf_bytes;
    uint8_t* sx = sa + a_bytes;
    issue_act_tile<T16>(sa, aimg + ((long)yo * Wo + xo0) * lda, lda, nc, Wo - xo0, lane);
    for (int idx = lane; idx < rows_x * 8; idx += 32) {
      const int r = idx >> 3, c = idx & 7;               // row r = (f, ky), chunk c = input columns 4c .. 4c+3
      const bool ok = xo0 + 2 * c < Wo;
      const float* src = ximg + ((long)(r >> 1) * H + 2 * yo + (r & 1)) * W + 2 * xo0 + 4 * c;
      cp_async16_zfill(sx + r * kPsXRow + c * 16, ok ? src : ximg, ok);
    }
  };

  float c[kPsMT][NT][4];
#pragma unroll
  for (int mi = 0; mi < kPsMT; ++mi)
#pragma unroll
    for (int j = 0; j < NT; ++j) c[mi][j][0] = c[mi][j][1] = c[mi][j][2] = c[mi][j][3] = 0.f;

  int buf = 0;
  if (tile0 + warp < tile1) issue(tile0 + warp, 0);
  cp_commit();
  for (int tile = tile0 + warp; tile < tile1; tile += kPsWarps) {
    if (tile + kPsWarps < tile1) issue(tile + kPsWarps, buf ^ 1);
    cp_commit();
    cp_wait<1>();
    __syncwarp();
    const uint8_t* sa = my + buf * buf_bytes;
    const uint8_t* sx = sa + a_bytes;
    // B fragments: b0 = (pixel 2t, 2t+1; column g), b1 = (pixel 2t+8, 2t+9; column g), column kk = 8j + g = f*4 + ky*2 + kx
    // the fp32 field values enter as a bf16 hi + lo pair (two MMAs), i.e. with ~16 significant bits
    uint32_t b[NT][2], bl[NT][2];
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const int kk = 8 * j + g;
      if (kk < K) {
        const float* xr = reinterpret_cast<const float*>(sx + (kk >> 1) * kPsXRow) + (kk & 1);
        const float v[4] = {xr[4 * t], xr[4 * t + 2], xr[4 * t + 16], xr[4 * t + 18]};
        float r[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) r[q] = v[q] - __bfloat162float(__float2bfloat16_rn(v[q]));
        b[j][0] = pack2<__nv_bfloat16>(v[0], v[1]);
        b[j][1] = pack2<__nv_bfloat16>(v[2], v[3]);
        bl[j][0] = pack2<__nv_bfloat16>(r[0], r[1]);
        bl[j][1] = pack2<__nv_bfloat16>(r[2], r[3]);
      } else {
        b[j][0] = b[j][1] = bl[j][0] = bl[j][1] = 0u;
      }
    }
    const uint32_t a_addr = smem_u32(sa) + (((lane & 7) + 8 * (lane >> 4)) * kPsSA + 8 * ((lane >> 3) & 1)) * 2;
#pragma unroll
    for (int mi = 0; mi < kPsMT; ++mi) {
      if (mi < mts) {
        uint32_t af[4];
        ldsm_x4_trans(af, a_addr + mi * 32);
        if constexpr (std::is_same<T16, __half>::value) {
#pragma unroll
          for (int q = 0; q < 4; ++q) { const float2 f2 = unpack2<__half>(af[q]); af[q] = pack2<__nv_bfloat16>(f2.x, f2.y); }
        }
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          mma_16816<__nv_bfloat16>(c[mi][j], af, b[j][0], b[j][1]);
          mma_16816<__nv_bfloat16>(c[mi][j], af, bl[j][0], bl[j][1]);
        }
      }
    }
    __syncwarp();                         // every lane is done with `buf` before the next issue overwrites it
    buf ^= 1;
  }
  cp_wait<0>();
  __syncthreads();
  // block reduction: c0/c1 = (channel 16mi+g, column 8j+2t / +1), c2/c3 = channel +8
  float* sD = reinterpret_cast<float*>(smem);            // [nc][8*NT]
  const int KD = 8 * NT;
  for (int i = threadIdx.x; i < nc * KD; i += blockDim.x) sD[i] = 0.f;
  __syncthreads();
#pragma unroll
  for (int mi = 0; mi < kPsMT; ++mi) {
    if (mi < mts) {
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        float* d0 = sD + (16 * mi + g) * KD + 8 * j + 2 * t;
        atomicAdd(d0, c[mi][j][0]); atomicAdd(d0 + 1, c[mi][j][1]);
        atomicAdd(d0 + 8 * KD, c[mi][j][2]); atomicAdd(d0 + 8 * KD + 1, c[mi][j][3]);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nc * KD; i += blockDim.x) {
    const int n = i / KD, kk = i - n * KD;
    if (kk < K) atomicAdd(dW + (long)n * ldw + kk, sD[i]);
  }
}

// ---------------------------------------------------------------------------------------------
// conv-transpose out.  A operand = the activation tile (pixels x channels, ldmatrix), B operand = the weights, kept in
// registers for the whole kernel as a 16-bit hi + lo pair (two MMAs per fragment), so the fp32 weights are used at
// ~22 bits although the tensor cores see 16-bit operands.  Each accumulator pair is one float2 of an output row.
// ---------------------------------------------------------------------------------------------
template <typename T16, int NT>
__global__ void __launch_bounds__(kPsWarps * 32, 2)
patch_out_mma_kernel(const T16* __restrict__ a, int lda, const float* __restrict__ Wck, float* __restrict__ out,
                     int F, int h, int w, int C, int accumulate, int tiles_per_block) {
  pdl_prologue_done();
  extern __shared__ __align__(16) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int K = 4 * F;
  const int a_bytes = 16 * kPsSA * 2;
  uint8_t* my = smem + (size_t)warp * 2 * a_bytes;
  const int tiles_x = (w + 15) / 16, n_tiles = h * tiles_x;
  const int img = blockIdx.y;
  const int tile0 = blockIdx.x * tiles_per_block, tile1 = min(n_tiles, tile0 + tiles_per_block);
  const int kts = C >> 4;                                // 16-channel k steps (<= kPsMT)
  const T16* aimg = a + (long)img * h * w * lda;      // C channels of this pass out of lda per pixel
  float* oimg = out + (long)img * F * (2 * h) * (2 * w);

  // B fragments for k step ks, n tile j: b0 = (channel 16ks+2t, +1; column 8j+g), b1 = channels +8
  uint32_t bh[kPsMT][NT][2], bl[kPsMT][NT][2];
#pragma unroll
  for (int ks = 0; ks < kPsMT; ++ks) {
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      const int kk = 8 * j + g;
      if (ks < kts && kk < K) {
        const float* wp = Wck + (long)(16 * ks + 2 * t) * K + kk;
        v[0] = __ldg(wp); v[1] = __ldg(wp + K); v[2] = __ldg(wp + 8 * K); v[3] = __ldg(wp + 9 * K);
      }
      float r[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) r[q] = v[q] - to_f32<T16>(from_f32<T16>(v[q]));
      bh[ks][j][0] = pack2<T16>(v[0], v[1]); bh[ks][j][1] = pack2<T16>(v[2], v[3]);
      bl[ks][j][0] = pack2<T16>(r[0], r[1]); bl[ks][j][1] = pack2<T16>(r[2], r[3]);
    }
  }

  int buf = 0;
  auto issue = [&](int tile, int bf_) {
    const int y = tile / tiles_x, x0 = (tile - y * tiles_x) * 16;
    issue_act_tile<T16>(my + bf_ * a_bytes, aimg + ((long)y * w + x0) * lda, lda, C, w - x0, lane);
  };
  if (tile0 + warp < tile1) issue(tile0 + warp, 0);
  cp_commit();
  for (int tile = tile0 + warp; tile < tile1; tile += kPsWarps) {
    if (tile + kPsWarps < tile1) issue(tile + kPsWarps, buf ^ 1);
    cp_commit();
    cp_wait<1>();
    __syncwarp();
    const int y = tile / tiles_x, x0 = (tile - y * tiles_x) * 16;
    float c[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) c[j][0] = c[j][1] = c[j][2] = c[j][3] = 0.f;
    // A fragments (non-transposed): matrices (px 0-7, ch 0-7), (px 8-15, ch 0-7), (px 0-7, ch 8-15), (px 8-15, ch 8-15)
    const uint32_t a_addr = smem_u32(my + buf * a_bytes) + (((lane & 7) + 8 * ((lane >> 3) & 1)) * kPsSA + 8 * (lane >> 4)) * 2;
#pragma unroll
    for (int ks = 0; ks < kPsMT; ++ks) {
      if (ks < kts) {
        uint32_t af[4];
        ldsm_x4(af, a_addr + ks * 32);
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          mma_16816<T16>(c[j], af, bh[ks][j][0], bh[ks][j][1]);
          mma_16816<T16>(c[j], af, bl[ks][j][0], bl[ks][j][1]);
        }
      }
    }
    __syncwarp();
    buf ^= 1;
    // c0,c1 = (px g; column 8j+2t, +1) -> f = 2j + t/2, ky = t%2, kx = 0,1: one float2 of row 2y+ky; c2,c3 = px g+8
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const int f = 2 * j + (t >> 1), ky = t & 1;
      if (f < F) {
        float* op = oimg + ((long)f * (2 * h) + 2 * y + ky) * (2 * w) + 2 * (x0 + g);
        if (accumulate) {             // a later channel pass of the same output (more than 96 channels)
          if (x0 + g < w) { const float2 o = *reinterpret_cast<const float2*>(op); c[j][0] += o.x; c[j][1] += o.y; }
          if (x0 + g + 8 < w) { const float2 o = *reinterpret_cast<const float2*>(op + 16); c[j][2] += o.x; c[j][3] += o.y; }
        }
        if (x0 + g < w) *reinterpret_cast<float2*>(op) = make_float2(c[j][0], c[j][1]);
        if (x0 + g + 8 < w) *reinterpret_cast<float2*>(op + 16) = make_float2(c[j][2], c[j][3]);
      }
    }
  }
  cp_wait<0>();
}

static int plan_blocks(int tiles, int I, int& tpb) {
  int bpi = (2 * num_sms() + I - 1) / I;                 // two resident blocks per SM in total
  const int max_bpi = (tiles + kPsWarps - 1) / kPsWarps;
  if (bpi > max_bpi) bpi = max_bpi;
  if (bpi < 1) bpi = 1;
  tpb = (tiles + bpi - 1) / bpi;
  return (tiles + tpb - 1) / tpb;
}

bool patch_wgrad_mma_ok(int F, int W, int N) { return F >= 1 && F <= 8 && W % 4 == 0 && N % 16 == 0; }
// more than 96 channels (film_avit_big: 192) run as passes of 96 that accumulate into the output
bool patch_out_mma_ok(int F, int C) { return F >= 1 && F <= 8 && C % 16 == 0 && C <= 8 * 16 * kPsMT; }

int launch_patch_wgrad_mma(const void* a, int dtype, const float* x, float* dW, int I, int F, int H, int W, int N,
                           cudaStream_t s) {
  const int Ho = H / 2, Wo = W / 2;
  const int tiles = Ho * ((Wo + 15) / 16);
  int tpb;
  const int bpi = plan_blocks(tiles, I, tpb);
  const int NT = F <= 2 ? 1 : (F <= 4 ? 2 : 4);
  const size_t sm = (size_t)kPsWarps * 2 * (16 * kPsSA * 2 + 2 * F * kPsXRow);
  dim3 grid(bpi, I);
  for (int n0 = 0; n0 < N; n0 += 16 * kPsMT) {
    const int nc = N - n0 < 16 * kPsMT ? N - n0 : 16 * kPsMT;
    float* dWc = dW + (long)n0 * 4 * F;
#define BF_PW_(T, NT_)                                                                                              \
    do {                                                                                                            \
      static bool done_ = false;                                                                                    \
      if (!done_) {                                                                                                 \
        if (int e_ = check_cuda(cudaFuncSetAttribute(patch_wgrad_mma_kernel<T, NT_>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                                     100 * 1024), "cudaFuncSetAttribute(patch_wgrad)")) return e_;  \
        done_ = true;                                                                                               \
      }                                                                                                             \
      launch_k(patch_wgrad_mma_kernel<T, NT_>, grid, dim3(kPsWarps * 32), sm, s, (const T*)a + n0, (long)N, x, dWc, 4 * F, \
               F, H, W, nc, tpb);                                                                                   \
    } while (0)
#define BF_PW(T) do { if (NT == 1) BF_PW_(T, 1); else if (NT == 2) BF_PW_(T, 2); else BF_PW_(T, 4); } while (0)
    if (dtype == BF_BF16) BF_PW(__nv_bfloat16); else BF_PW(__half);
#undef BF_PW
#undef BF_PW_
    count_launch();
    BF_LAUNCH_CHECK("patch_wgrad_mma_kernel");
  }
  return BF_OK;
}

int launch_patch_out_mma(const void* a, int dtype, const float* Wck, float* out, int I, int F, int h, int w, int C,
                         cudaStream_t s) {
  const int tiles = h * ((w + 15) / 16);
  int tpb;
  const int bpi = plan_blocks(tiles, I, tpb);
  const int NT = F <= 2 ? 1 : (F <= 4 ? 2 : 4);
  const size_t sm = (size_t)kPsWarps * 2 * (16 * kPsSA * 2);
  dim3 grid(bpi, I);
#define BF_PO_(T, NT_)                                                                                              \
  do {                                                                                                            \
    static bool done_ = false;                                                                                    \
    if (!done_) {                                                                                                 \
      if (int e_ = check_cuda(cudaFuncSetAttribute(patch_out_mma_kernel<T, NT_>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                                   64 * 1024), "cudaFuncSetAttribute(patch_out)")) return e_;     \
      done_ = true;                                                                                               \
    }                                                                                                             \
    launch_k(patch_out_mma_kernel<T, NT_>, grid, dim3(kPsWarps * 32), sm, s, (const T*)a + c_off, C, Wck + (long)c_off * 4 * F, \
             out, F, h, w, cp, c_off > 0 ? 1 : 0, tpb);                                                           \
  } while (0)
#define BF_PO(T) do { if (NT == 1) BF_PO_(T, 1); else if (NT == 2) BF_PO_(T, 2); else BF_PO_(T, 4); } while (0)
  for (int c_off = 0; c_off < C; c_off += 16 * kPsMT) {
    const int cp = C - c_off < 16 * kPsMT ? C - c_off : 16 * kPsMT;
    if (dtype == BF_BF16) BF_PO(__nv_bfloat16); else BF_PO(__half);
    count_launch();
    BF_LAUNCH_CHECK("patch_out_mma_kernel");
  }
#undef BF_PO
#undef BF_PO_
  return BF_OK;
}

}  // namespace bf
