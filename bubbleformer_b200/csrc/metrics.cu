// Rollout metrics on the device: the physics checks upstream applies to predicted fields.
//   bf_eikonal_sums : sums[s] += sum over the H x W pixels of slab s of (|grad phi| - 1)^2, with torch.gradient's
//                     stencil (central differences inside, one-sided first order at the edges, spacing dx)
//                     -- upstream utils/losses.py:5-15 (eikonal_loss = mean over everything of that quantity)
//   bf_heatflux_rows: flux[t] = mean_x [ heater(x) & dfun[t,0,x] < 0 ] * (heater_temp - temp[t,0,x]) * 0.054 / (dx*lc)
//                     -- upstream utils/heatflux.py:3-38 (wall row y = 0, heater between x = -5 and 5 of a domain
//                     [-8, 8] sampled at cell centres); the caller takes mean and max over t
#include "common.cuh"

namespace bf {

__global__ void __launch_bounds__(256)
eikonal_sums_kernel(const float* __restrict__ phi, float* __restrict__ sums, int H, int W, float inv_dx, int rows_per_block) {
  pdl_prologue_done();
  const int slab = blockIdx.y;
  const float* f = phi + (long)slab * H * W;
  const int y0 = blockIdx.x * rows_per_block, y1 = min(H, y0 + rows_per_block);
  float acc = 0.f;
  for (int idx = y0 * W + threadIdx.x; idx < y1 * W; idx += blockDim.x) {
    const int y = idx / W, x = idx - y * W;
    const int ym = max(y - 1, 0), yp = min(y + 1, H - 1), xm = max(x - 1, 0), xp = min(x + 1, W - 1);
    // central difference over 2 dx inside, one-sided over dx at an edge (the clamped index makes the span 1)
    const float gy = (f[yp * W + x] - f[ym * W + x]) * inv_dx / (float)max(yp - ym, 1);
    const float gx = (f[y * W + xp] - f[y * W + xm]) * inv_dx / (float)max(xp - xm, 1);
    const float e = sqrtf(gy * gy + gx * gx) - 1.f;
    acc = fmaf(e, e, acc);
  }
  acc = warp_sum(acc);
  __shared__ float sh[8];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f;
    for (int i = 0; i < 8; ++i) a += sh[i];
    atomicAdd(sums + slab, a);
  }
}

__global__ void __launch_bounds__(256)
heatflux_rows_kernel(const float* __restrict__ dfun, const float* __restrict__ temp, float* __restrict__ flux, long frame_stride,
                     int W, float heater_temp, float x_min, float dx, float scale) {
  pdl_prologue_done();
  const int t = blockIdx.x;
  const float* d = dfun + (long)t * frame_stride;       // row y = 0 of frame t
  const float* tp = temp + (long)t * frame_stride;
  float acc = 0.f;
  for (int x = threadIdx.x; x < W; x += blockDim.x) {
    const float xc = x_min + ((float)x + 0.5f) * dx;
    if (xc >= -5.0f && xc <= 5.0f && d[x] < 0.f) acc += heater_temp - tp[x];
  }
  acc = warp_sum(acc);
  __shared__ float sh[8];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f;
    for (int i = 0; i < 8; ++i) a += sh[i];
    flux[t] = a * scale / (float)W;
  }
}

}  // namespace bf

using namespace bf;

extern "C" int bf_eikonal_sums(const float* phi, float* sums, int64_t slabs, int H, int W, float dx, void* stream) {
  BF_REQUIRE(phi && sums && slabs > 0 && slabs <= 65535 && H >= 2 && W >= 2 && dx > 0.f, "bf_eikonal_sums: bad arguments");
  int bpy = (int)((4L * num_sms() + slabs - 1) / slabs);
  if (bpy > H) bpy = H;
  if (bpy < 1) bpy = 1;
  const int rpb = (H + bpy - 1) / bpy;
  dim3 grid((H + rpb - 1) / rpb, (unsigned)slabs);
  launch_k(eikonal_sums_kernel, grid, dim3(256), (size_t)0, static_cast<cudaStream_t>(stream), phi, sums, H, W, 1.f / dx, rpb);
  count_launch();
  BF_LAUNCH_CHECK("eikonal_sums_kernel");
  return BF_OK;
}

extern "C" int bf_heatflux_rows(const float* dfun, const float* temp, float* flux, int64_t frames, int64_t frame_stride,
                                int W, float heater_temp, float x_min, float dx, float lc, void* stream) {
  BF_REQUIRE(dfun && temp && flux && frames > 0 && W > 0 && dx > 0.f && lc > 0.f, "bf_heatflux_rows: bad arguments");
  launch_k(heatflux_rows_kernel, dim3((unsigned)frames), dim3(256), (size_t)0, static_cast<cudaStream_t>(stream), dfun, temp,
           flux, (long)frame_stride, W, heater_temp, x_min, dx, 0.054f / (dx * lc));
  count_launch();
  BF_LAUNCH_CHECK("heatflux_rows_kernel");
  return BF_OK;
}
