// Optimiser step over the flat parameter / gradient buffers (one launch for all 28.9 M parameters), fused with the
// refresh of the 16-bit operand mirror the GEMMs read.  Replaces, for the hot path's parameters, what upstream gets
// from torch.optim.AdamW / Adam and lion_pytorch.Lion (bubbleformer/modules.py:132-142, config/optim_cfg/*.yaml).
//   LION  : p *= 1 - lr*wd;  p -= lr * sign(b1*m + (1-b1)*g);  m = b2*m + (1-b2)*g          (lion_pytorch update_fn)
//   ADAMW : p *= 1 - lr*wd;  m = b1*m + (1-b1)*g;  v = b2*v + (1-b2)*g^2;
//           p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)                               (torch.optim.AdamW)
//   ADAM  : g += wd*p (L2), then the Adam update without decoupled decay                      (torch.optim.Adam)
// A pure HBM stream: 16-byte loads / stores, grid-stride, ~4 waves.
#include "common.cuh"

namespace bf {

struct OptimArgs {
  float* p; const float* g; float* m; float* v; __nv_bfloat16* p16;
  long n4;
  float lr, b1, b2, eps, wd, bc1, bc2s;      // bc1 = 1 - b1^t, bc2s = sqrt(1 - b2^t)
};

template <int KIND>
__global__ void __launch_bounds__(256) optim_kernel(OptimArgs a) {
  pdl_prologue_done();
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n4; i += (long)gridDim.x * blockDim.x) {
    float4 p4 = reinterpret_cast<float4*>(a.p)[i];
    const float4 g4 = reinterpret_cast<const float4*>(a.g)[i];
    float4 m4 = reinterpret_cast<float4*>(a.m)[i];
    float4 v4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (KIND != BF_OPT_LION) v4 = reinterpret_cast<float4*>(a.v)[i];
    float* p = reinterpret_cast<float*>(&p4);
    const float* g = reinterpret_cast<const float*>(&g4);
    float* m = reinterpret_cast<float*>(&m4);
    float* v = reinterpret_cast<float*>(&v4);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (KIND == BF_OPT_LION) {
        float w = p[k] * (1.f - a.lr * a.wd);
        const float u = a.b1 * m[k] + (1.f - a.b1) * g[k];
        w -= a.lr * (u > 0.f ? 1.f : (u < 0.f ? -1.f : 0.f));
        m[k] = a.b2 * m[k] + (1.f - a.b2) * g[k];
        p[k] = w;
      } else {
        float gk = g[k], w = p[k];
        if (KIND == BF_OPT_ADAM) gk = fmaf(a.wd, w, gk);
        else w *= (1.f - a.lr * a.wd);
        m[k] = a.b1 * m[k] + (1.f - a.b1) * gk;
        v[k] = a.b2 * v[k] + (1.f - a.b2) * gk * gk;
        const float denom = sqrtf(v[k]) / a.bc2s + a.eps;
        p[k] = w - (a.lr / a.bc1) * (m[k] / denom);
      }
    }
    reinterpret_cast<float4*>(a.p)[i] = p4;
    reinterpret_cast<float4*>(a.m)[i] = m4;
    if (KIND != BF_OPT_LION) reinterpret_cast<float4*>(a.v)[i] = v4;
    if (a.p16 != nullptr) {
      uint2 u;
      u.x = pack2<__nv_bfloat16>(p[0], p[1]);
      u.y = pack2<__nv_bfloat16>(p[2], p[3]);
      reinterpret_cast<uint2*>(a.p16)[i] = u;
    }
  }
}

}  // namespace bf

using namespace bf;

extern "C" int bf_optim_step(int kind, float* p, const float* g, float* m, float* v, void* p16, int64_t n, float lr,
                             float beta1, float beta2, float eps, float weight_decay, int64_t step, void* stream) {
  BF_REQUIRE(kind == BF_OPT_LION || kind == BF_OPT_ADAMW || kind == BF_OPT_ADAM, "bf_optim_step: kind %d", kind);
  BF_REQUIRE(p && g && m && (kind == BF_OPT_LION || v), "bf_optim_step: null buffer");
  BF_REQUIRE(n > 0 && n % 4 == 0, "bf_optim_step: n=%ld must be a positive multiple of 4", (long)n);
  BF_REQUIRE(((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
               reinterpret_cast<uintptr_t>(v)) & 15) == 0 && (reinterpret_cast<uintptr_t>(p16) & 7) == 0,
             "bf_optim_step: buffers must be 16-byte aligned");
  BF_REQUIRE(step >= 1, "bf_optim_step: step counts from 1");
  OptimArgs a{p, g, m, v, static_cast<__nv_bfloat16*>(p16), n / 4, lr, beta1, beta2, eps, weight_decay,
              1.f - powf(beta1, (float)step), sqrtf(1.f - powf(beta2, (float)step))};
  long blocks = (a.n4 + 255) / 256;
  if (blocks > 8L * num_sms()) blocks = 8L * num_sms();
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (kind == BF_OPT_LION) launch_k(optim_kernel<BF_OPT_LION>, dim3((unsigned)blocks), dim3(256), (size_t)0, s, a);
  else if (kind == BF_OPT_ADAMW) launch_k(optim_kernel<BF_OPT_ADAMW>, dim3((unsigned)blocks), dim3(256), (size_t)0, s, a);
  else launch_k(optim_kernel<BF_OPT_ADAM>, dim3((unsigned)blocks), dim3(256), (size_t)0, s, a);
  count_launch();
  BF_LAUNCH_CHECK("optim_kernel");
  return BF_OK;
}
