// Per-channel parameter arithmetic of the residual branches, as two small kernels instead of ~30 element-wise torch
// launches per block:
//   bf_feat_consts       : the feature-scaling constants of the axial block (upstream layers/attention.py:302-307);
//                          mean_img(z) of z = IN(o) W^T + b is exactly c = W b_norm2 + b_out, so the op is the affine
//                          z*c1 + c0 with c1 = 1 + high, c0 = c*(low - high)
//   bf_branch_param_grads: gradients of gamma / low / high / output_head.bias (and, through c, of output_head.weight
//                          and norm2.bias) from the per-image sums S0 = sum mask*dX_out, S1 = sum mask*dX_out*Z that
//                          bf_resid_bwd produced (layers/attention.py:123,309 reversed)
#include "common.cuh"

namespace bf {

// one block per output channel j: c[j] = sum_i W[j,i]*nb[i] + b[j]
__global__ void __launch_bounds__(128)
feat_consts_kernel(const float* __restrict__ W, const float* __restrict__ nb, const float* __restrict__ bout,
                   const float* __restrict__ lo, const float* __restrict__ hi, int E, float* __restrict__ c,
                   float* __restrict__ c1, float* __restrict__ c0) {
  pdl_prologue_done();
  __shared__ float sh[4];
  const int j = blockIdx.x;
  float acc = 0.f;
  for (int i = threadIdx.x; i < E; i += blockDim.x) acc = fmaf(W[(long)j * E + i], nb[i], acc);
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    const float cj = sh[0] + sh[1] + sh[2] + sh[3] + bout[j];
    c[j] = cj;
    c1[j] = 1.f + hi[j];
    c0[j] = cj * (lo[j] - hi[j]);
  }
}

struct BranchGradArgs {
  const float* S01;        // (2, I, E)
  int I, E;
  const float* gamma;      // layer scale of the branch
  // feature scaling (all null when off)
  const float* c; const float* c1; const float* c0; const float* lo; const float* hi;
  const float* W;          // output_head.weight (E, E) fp32
  const float* nb;         // norm2.bias
  float* d_gamma; float* d_bout; float* d_lo; float* d_hi; float* d_W; float* d_nb;
};

// one block per channel j
__global__ void __launch_bounds__(128) branch_param_grads_kernel(BranchGradArgs a) {
  pdl_prologue_done();
  __shared__ float sh[2][4];
  __shared__ float s_dc;
  const int j = blockIdx.x, E = a.E;
  float s0 = 0.f, s1 = 0.f;
  for (int img = threadIdx.x; img < a.I; img += blockDim.x) {
    s0 += a.S01[(long)img * E + j];
    s1 += a.S01[((long)a.I + img) * E + j];
  }
  s0 = warp_sum(s0); s1 = warp_sum(s1);
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = s0; sh[1][threadIdx.x >> 5] = s1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    const float S0 = sh[0][0] + sh[0][1] + sh[0][2] + sh[0][3];
    const float S1 = sh[1][0] + sh[1][1] + sh[1][2] + sh[1][3];
    const float ga = a.gamma[j];
    if (a.c == nullptr) {
      a.d_gamma[j] += S1;
      a.d_bout[j] += ga * S0;
      s_dc = 0.f;
    } else {
      const float c = a.c[j], c1 = a.c1[j], c0 = a.c0[j];
      const float dc = ga * (a.lo[j] - a.hi[j]) * S0;           // gradient reaching c = W b_norm2 + b_out
      a.d_gamma[j] += c1 * S1 + c0 * S0;
      a.d_hi[j] += ga * (S1 - c * S0);
      a.d_lo[j] += ga * c * S0;
      a.d_bout[j] += ga * c1 * S0 + dc;
      s_dc = dc;
    }
  }
  __syncthreads();
  if (a.c != nullptr) {
    const float dc = s_dc;
    for (int i = threadIdx.x; i < E; i += blockDim.x) {
      a.d_W[(long)j * E + i] += dc * a.nb[i];
      atomicAdd(a.d_nb + i, a.W[(long)j * E + i] * dc);
    }
  }
}

}  // namespace bf

using namespace bf;

extern "C" int bf_feat_consts(const float* W, const float* norm2_bias, const float* out_bias, const float* low,
                              const float* high, int E, float* c, float* c1, float* c0, void* stream) {
  BF_REQUIRE(W && norm2_bias && out_bias && low && high && c && c1 && c0 && E > 0, "bf_feat_consts: bad arguments");
  launch_k(feat_consts_kernel, dim3(E), dim3(128), (size_t)0, static_cast<cudaStream_t>(stream), W, norm2_bias, out_bias,
           low, high, E, c, c1, c0);
  count_launch();
  BF_LAUNCH_CHECK("feat_consts_kernel");
  return BF_OK;
}

extern "C" int bf_branch_param_grads(const bf_branch_grad_args* g, void* stream) {
  BF_REQUIRE(g && g->S01 && g->gamma && g->d_gamma && g->d_out_bias && g->I > 0 && g->E > 0,
             "bf_branch_param_grads: bad arguments");
  const bool fs = g->c != nullptr;
  BF_REQUIRE(!fs || (g->c1 && g->c0 && g->low && g->high && g->W && g->norm2_bias && g->d_low && g->d_high && g->d_W &&
                     g->d_norm2_bias),
             "bf_branch_param_grads: feature scaling needs all of c/c1/c0/low/high/W/norm2_bias and their gradients");
  BranchGradArgs a{g->S01, g->I, g->E, g->gamma, g->c, g->c1, g->c0, g->low, g->high, g->W, g->norm2_bias,
                   g->d_gamma, g->d_out_bias, g->d_low, g->d_high, g->d_W, g->d_norm2_bias};
  launch_k(branch_param_grads_kernel, dim3(g->E), dim3(128), (size_t)0, static_cast<cudaStream_t>(stream), a);
  count_launch();
  BF_LAUNCH_CHECK("branch_param_grads_kernel");
  return BF_OK;
}
