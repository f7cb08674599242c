// Relative-L2 training loss (upstream utils/losses.py:67-94 as configured at modules.py:50:
// LpLoss(d=2, p=2, reduce_dims=[0,1,2], reductions=[mean, mean, sum])) fused into two streaming passes.
//
//   forward : sums[s] = (sum (pred - tgt)^2, sum tgt^2) over the H*W pixels of slab s = (b, t, c)
//             loss = sum_c mean_b mean_t sqrt(sums[s].0 / sums[s].1)        (finished by the caller: tiny)
//   backward: dpred[s, :] = coef[s] * (pred - tgt),  coef[s] = g / (B*T * sqrt(sums[s].0 * sums[s].1))
// Both kernels are pure HBM streams over fp32 NCHW fields (16-byte loads, grid sized to a few waves).
#include "common.cuh"

namespace bf {

__global__ void __launch_bounds__(256)
lploss_sums_kernel(const float4* __restrict__ pred, const float4* __restrict__ tgt, float* __restrict__ sums,
                   long vec_per_slab, int blocks_per_slab) {
  pdl_prologue_done();
  const int slab = blockIdx.x / blocks_per_slab, part = blockIdx.x - slab * blocks_per_slab;
  const float4* p = pred + (long)slab * vec_per_slab;
  const float4* t = tgt + (long)slab * vec_per_slab;
  float d2 = 0.f, t2 = 0.f;
  for (long i = (long)part * 256 + threadIdx.x; i < vec_per_slab; i += (long)blocks_per_slab * 256) {
    const float4 a = __ldg(p + i), b = __ldg(t + i);
    const float x = a.x - b.x, y = a.y - b.y, z = a.z - b.z, w = a.w - b.w;
    d2 += x * x + y * y + z * z + w * w;
    t2 += b.x * b.x + b.y * b.y + b.z * b.z + b.w * b.w;
  }
  d2 = warp_sum(d2); t2 = warp_sum(t2);
  __shared__ float sh[2][8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sh[0][warp] = d2; sh[1][warp] = t2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
    for (int i = 0; i < 8; ++i) { a += sh[0][i]; b += sh[1][i]; }
    atomicAdd(sums + 2 * slab, a);
    atomicAdd(sums + 2 * slab + 1, b);
  }
}

__global__ void __launch_bounds__(256)
lploss_bwd_kernel(const float4* __restrict__ pred, const float4* __restrict__ tgt, const float* __restrict__ coef,
                  float4* __restrict__ dpred, long vec_per_slab, int blocks_per_slab) {
  pdl_prologue_done();
  const int slab = blockIdx.x / blocks_per_slab, part = blockIdx.x - slab * blocks_per_slab;
  const float c = __ldg(coef + slab);
  const long base = (long)slab * vec_per_slab;
  for (long i = (long)part * 256 + threadIdx.x; i < vec_per_slab; i += (long)blocks_per_slab * 256) {
    const float4 a = __ldg(pred + base + i), b = __ldg(tgt + base + i);
    dpred[base + i] = make_float4(c * (a.x - b.x), c * (a.y - b.y), c * (a.z - b.z), c * (a.w - b.w));
  }
}

static int plan(long slabs, long n_per_slab, int& bps) {
  BF_REQUIRE(slabs > 0 && n_per_slab > 0 && n_per_slab % 4 == 0, "bf_lploss: slab size %ld must be a positive multiple of 4", n_per_slab);
  BF_REQUIRE(slabs < (1l << 24), "bf_lploss: too many slabs");
  long want = 8L * num_sms();
  long b = (want + slabs - 1) / slabs;
  const long maxb = (n_per_slab / 4 + 255) / 256;
  if (b > maxb) b = maxb;
  if (b < 1) b = 1;
  bps = (int)b;
  return BF_OK;
}

}  // namespace bf

using namespace bf;

extern "C" int bf_lploss_sums(const float* pred, const float* tgt, float* sums, int64_t slabs, int64_t n_per_slab,
                              void* stream) {
  BF_REQUIRE(pred && tgt && sums, "bf_lploss_sums: null pointer");
  BF_REQUIRE(((reinterpret_cast<uintptr_t>(pred) | reinterpret_cast<uintptr_t>(tgt)) & 15) == 0, "bf_lploss_sums: alignment");
  int bps;
  if (int st = plan(slabs, n_per_slab, bps)) return st;
  launch_k(lploss_sums_kernel, dim3((unsigned)(slabs * bps)), dim3(256), (size_t)(0), static_cast<cudaStream_t>(stream), 
      reinterpret_cast<const float4*>(pred), reinterpret_cast<const float4*>(tgt), sums, n_per_slab / 4, bps);
  count_launch();
  BF_LAUNCH_CHECK("lploss_sums_kernel");
  return BF_OK;
}

extern "C" int bf_lploss_bwd(const float* pred, const float* tgt, const float* coef, float* dpred, int64_t slabs,
                             int64_t n_per_slab, void* stream) {
  BF_REQUIRE(pred && tgt && coef && dpred, "bf_lploss_bwd: null pointer");
  BF_REQUIRE(((reinterpret_cast<uintptr_t>(pred) | reinterpret_cast<uintptr_t>(tgt) | reinterpret_cast<uintptr_t>(dpred)) & 15) == 0,
             "bf_lploss_bwd: alignment");
  int bps;
  if (int st = plan(slabs, n_per_slab, bps)) return st;
  launch_k(lploss_bwd_kernel, dim3((unsigned)(slabs * bps)), dim3(256), (size_t)(0), static_cast<cudaStream_t>(stream), 
      reinterpret_cast<const float4*>(pred), reinterpret_cast<const float4*>(tgt), coef, reinterpret_cast<float4*>(dpred),
      n_per_slab / 4, bps);
  count_launch();
  BF_LAUNCH_CHECK("lploss_bwd_kernel");
  return BF_OK;
}
