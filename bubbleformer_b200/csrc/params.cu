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
                   const float* __restrict__ lo, const float* __restrict__ hi, const float* __restrict__ gamma, int E,
                   float* __restrict__ c, float* __restrict__ c1, float* __restrict__ c0, float* __restrict__ coef) {
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
    if (coef != nullptr) coef[j] = gamma[j] * (1.f + hi[j]);
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

// ---------------------------------------------------------------------------------------------
// FiLM conditioning vector (upstream layers/linear_layers.py:58-61,71-72):  gb[b] = Linear(LayerNorm_F(cond[b]))
// F = number of fluid parameters (9 in every shipped config), 2E outputs.  One block per sample.
// ---------------------------------------------------------------------------------------------
constexpr int kFilmMaxF = 32;

__device__ __forceinline__ void film_ln_row(const float* cond, int F, const float* lw, const float* lb, float* xhat,
                                            float* c, float& rstd_out) {
  float mean = 0.f;
  for (int f = 0; f < F; ++f) mean += cond[f];
  mean /= (float)F;
  float var = 0.f;
  for (int f = 0; f < F; ++f) { const float d = cond[f] - mean; var = fmaf(d, d, var); }
  const float rstd = rsqrtf(var / (float)F + 1e-5f);
  for (int f = 0; f < F; ++f) {
    xhat[f] = (cond[f] - mean) * rstd;
    c[f] = fmaf(xhat[f], lw[f], lb[f]);
  }
  rstd_out = rstd;
}

__global__ void __launch_bounds__(256)
film_fwd_kernel(const float* __restrict__ cond, int F, const float* __restrict__ lw, const float* __restrict__ lb,
                const float* __restrict__ W, const float* __restrict__ bias, int E2, float* __restrict__ gb) {
  pdl_prologue_done();
  __shared__ float s_c[kFilmMaxF], s_x[kFilmMaxF];
  const int b = blockIdx.x;
  if (threadIdx.x == 0) {
    float rstd;
    film_ln_row(cond + (long)b * F, F, lw, lb, s_x, s_c, rstd);
  }
  __syncthreads();
  for (int j = threadIdx.x; j < E2; j += blockDim.x) {
    float acc = bias[j];
    for (int f = 0; f < F; ++f) acc = fmaf(W[(long)j * F + f], s_c[f], acc);
    gb[(long)b * E2 + j] = acc;
  }
}

// backward, two tiny launches:
//  (1) one thread per output j (grid over j): d_W[j,f] += sum_b dgb[b,j] c[b,f], d_bias[j] += sum_b dgb[b,j], and the
//      block's share of dc[b,f] = sum_j dgb[b,j] W[j,f] reduced through shuffles, one atomic per (b, f) per block
//  (2) LayerNorm backward over F for the affine parameters: d_lw[f] += sum_b dc*xhat, d_lb[f] += sum_b dc
__global__ void __launch_bounds__(128)
film_bwd_w_kernel(const float* __restrict__ dgb, const float* __restrict__ cond, int B, int F,
                  const float* __restrict__ lw, const float* __restrict__ lb, const float* __restrict__ W, int E2,
                  float* __restrict__ d_W, float* __restrict__ d_bias, float* __restrict__ dc) {
  pdl_prologue_done();
  extern __shared__ float sm[];            // c[B][F], xhat scratch [F]
  float* s_c = sm;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    float xh[kFilmMaxF];
    float rstd;
    film_ln_row(cond + (long)b * F, F, lw, lb, xh, s_c + b * F, rstd);
  }
  __syncthreads();
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = j < E2;
  float w[kFilmMaxF], dw[kFilmMaxF];
#pragma unroll
  for (int f = 0; f < kFilmMaxF; ++f) { w[f] = (live && f < F) ? W[(long)j * F + f] : 0.f; dw[f] = 0.f; }
  float db = 0.f;
  for (int b = 0; b < B; ++b) {
    const float g = live ? dgb[(long)b * E2 + j] : 0.f;
    db += g;
#pragma unroll
    for (int f = 0; f < kFilmMaxF; ++f) {
      if (f < F) {
        dw[f] = fmaf(g, s_c[b * F + f], dw[f]);
        const float v = warp_sum(g * w[f]);
        if ((threadIdx.x & 31) == 0) atomicAdd(dc + b * F + f, v);
      }
    }
  }
  if (live) {
    d_bias[j] += db;
#pragma unroll
    for (int f = 0; f < kFilmMaxF; ++f)
      if (f < F) d_W[(long)j * F + f] += dw[f];
  }
}

__global__ void __launch_bounds__(32)
film_bwd_ln_kernel(const float* __restrict__ dc, const float* __restrict__ cond, int B, int F,
                   const float* __restrict__ lw, const float* __restrict__ lb, float* __restrict__ d_lw,
                   float* __restrict__ d_lb) {
  pdl_prologue_done();
  const int f = threadIdx.x;
  if (f >= F) return;
  float a = 0.f, c = 0.f;
  for (int b = 0; b < B; ++b) {
    float xh[kFilmMaxF], cc[kFilmMaxF];
    float rstd;
    film_ln_row(cond + (long)b * F, F, lw, lb, xh, cc, rstd);
    a = fmaf(dc[b * F + f], xh[f], a);
    c += dc[b * F + f];
  }
  d_lw[f] += a;
  d_lb[f] += c;
}

}  // namespace bf

using namespace bf;

extern "C" int bf_film_fwd(const float* cond, int B, int F, const float* ln_w, const float* ln_b, const float* W,
                           const float* bias, int E2, float* gb, void* stream) {
  BF_REQUIRE(cond && ln_w && ln_b && W && bias && gb, "bf_film_fwd: null pointer");
  BF_REQUIRE(B > 0 && E2 > 0 && F > 0 && F <= kFilmMaxF, "bf_film_fwd: B=%d F=%d (<= %d) E2=%d", B, F, kFilmMaxF, E2);
  launch_k(film_fwd_kernel, dim3(B), dim3(256), (size_t)0, static_cast<cudaStream_t>(stream), cond, F, ln_w, ln_b, W,
           bias, E2, gb);
  count_launch();
  BF_LAUNCH_CHECK("film_fwd_kernel");
  return BF_OK;
}

extern "C" int bf_film_bwd(const float* dgb, const float* cond, int B, int F, const float* ln_w, const float* ln_b,
                           const float* W, int E2, float* d_ln_w, float* d_ln_b, float* d_W, float* d_bias,
                           float* dc_scratch, void* stream) {
  BF_REQUIRE(dgb && cond && ln_w && ln_b && W && d_ln_w && d_ln_b && d_W && d_bias && dc_scratch, "bf_film_bwd: null pointer");
  BF_REQUIRE(B > 0 && E2 > 0 && F > 0 && F <= kFilmMaxF, "bf_film_bwd: B=%d F=%d (<= %d) E2=%d", B, F, kFilmMaxF, E2);
  const size_t smem = (size_t)B * F * sizeof(float);
  BF_REQUIRE(smem <= 48 * 1024, "bf_film_bwd: batch %d too large", B);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  launch_k(film_bwd_w_kernel, dim3((E2 + 127) / 128), dim3(128), smem, st, dgb, cond, B, F, ln_w, ln_b, W, E2, d_W, d_bias,
           dc_scratch);
  launch_k(film_bwd_ln_kernel, dim3(1), dim3(32), (size_t)0, st, (const float*)dc_scratch, cond, B, F, ln_w, ln_b, d_ln_w,
           d_ln_b);
  count_launch(2);
  BF_LAUNCH_CHECK("film_bwd kernels");
  return BF_OK;
}
extern "C" int bf_feat_consts(const float* W, const float* norm2_bias, const float* out_bias, const float* low,
                              const float* high, const float* gamma, int E, float* c, float* c1, float* c0, float* coef,
                              void* stream) {
  BF_REQUIRE(W && norm2_bias && out_bias && low && high && c && c1 && c0 && E > 0, "bf_feat_consts: bad arguments");
  BF_REQUIRE((gamma == nullptr) == (coef == nullptr), "bf_feat_consts: gamma and coef go together");
  launch_k(feat_consts_kernel, dim3(E), dim3(128), (size_t)0, static_cast<cudaStream_t>(stream), W, norm2_bias, out_bias,
           low, high, gamma, E, c, c1, c0, coef);
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
