// InstanceNorm over the token grid of one image (per image, per channel), token-major data.
//
// Replaces nn.InstanceNorm2d(affine=True) at upstream layers/attention.py:39-40,76,120,153-154,197,208,298,316
// and layers/patching.py:45,102 (forward and autograd backward), fused with what follows it:
// GELU (patching.py:47,103), FiLM (linear_layers.py:71-77), layer-scale * drop-path + residual
// (attention.py:317).  Also the residual-branch backward helper and column sums for bias gradients.
//
// All kernels are HBM-bound streaming passes: 16-byte vector accesses, each thread owns a fixed set of
// channels and strides over rows, so consecutive threads touch consecutive addresses; block partials are
// combined in shared memory and flushed with one fp32 atomic per (image, channel).
// Statistics are kept as raw sums (sum x, sum x^2) per (image, channel) so producers can accumulate them.
#include "common.cuh"

namespace bf {

constexpr int kNormThreads = 256;

template <typename T> struct Vec;          // 16-byte vector of T
template <> struct Vec<float> { static constexpr int N = 4; };
template <> struct Vec<__nv_bfloat16> { static constexpr int N = 8; };
template <> struct Vec<__half> { static constexpr int N = 8; };

template <typename T, int N>
__device__ __forceinline__ void load_vec(const T* p, float (&v)[N], bool vec) {
  if (vec) {
    if constexpr (sizeof(T) == 4) {
      float4 u = *reinterpret_cast<const float4*>(p);
      v[0] = u.x; v[1] = u.y; v[2] = u.z; v[3] = u.w;
    } else {
      uint4 u = *reinterpret_cast<const uint4*>(p);
      float2 a = unpack2<T>(u.x), b = unpack2<T>(u.y), c = unpack2<T>(u.z), d = unpack2<T>(u.w);
      v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
    }
  } else {
#pragma unroll
    for (int j = 0; j < N; ++j) v[j] = to_f32<T>(p[j]);
  }
}
template <typename T, int N>
__device__ __forceinline__ void store_vec(T* p, const float (&v)[N], bool vec) {
  if (vec) {
    if constexpr (sizeof(T) == 4) {
      *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
      uint4 u;
      u.x = pack2<T>(v[0], v[1]); u.y = pack2<T>(v[2], v[3]); u.z = pack2<T>(v[4], v[5]); u.w = pack2<T>(v[6], v[7]);
      *reinterpret_cast<uint4*>(p) = u;
    }
  } else {
#pragma unroll
    for (int j = 0; j < N; ++j) p[j] = from_f32<T>(v[j]);
  }
}

// 8 consecutive channels per thread regardless of the storage type (16 B for 16-bit, 2 x 16 B for fp32)
template <typename T>
__device__ __forceinline__ void load8(const T* p, float (&v)[8]) {
  if constexpr (sizeof(T) == 4) {
    const float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
    load_vec<T, 8>(p, v, true);
  }
}
template <typename T>
__device__ __forceinline__ void store8(T* p, const float (&v)[8]) {
  if constexpr (sizeof(T) == 4) {
    reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
  } else {
    store_vec<T, 8>(p, v, true);
  }
}

// Thread geometry shared by all row-streaming kernels: a block covers `tx_n` vector columns and
// `ty_n` rows at a time; grid = (column chunks, row splits, images).
struct RowGeom {
  int vc;       // vector columns in a row (C / N)
  int tx_n;     // vector columns per block
  int ty_n;     // rows in flight per block
  int chunks;   // column chunks
  int splits;   // row splits per image
};
static RowGeom make_geom(int C, int vecn, int P, int I) {
  RowGeom g;
  g.vc = C / vecn;
  g.tx_n = g.vc < 128 ? g.vc : 128;
  g.chunks = (g.vc + g.tx_n - 1) / g.tx_n;
  g.ty_n = kNormThreads / g.tx_n;
  const int want_blocks = 4 * num_sms();
  int splits = (want_blocks + I * g.chunks - 1) / (I * g.chunks);
  const int max_splits = (P + 8 * g.ty_n - 1) / (8 * g.ty_n);     // keep >= 8 rows per thread (unrolled loops)
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  g.splits = splits;
  return g;
}

struct NormCommon {
  int I, P, C;
  long ldx;
  RowGeom g;
};

// ---------------------------------------------------------------------------------------------
// statistics: stats[img][c] += (sum x, sum x^2)
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kNormThreads)
inorm_stats_kernel(const T* __restrict__ x, NormCommon nc, float* __restrict__ stats) {
  constexpr int N = Vec<T>::N;
  extern __shared__ float red[];                       // [ty_n][tx_n][2N]
  const RowGeom& g = nc.g;
  const int tx = threadIdx.x % g.tx_n, ty = threadIdx.x / g.tx_n;
  const int vcol = blockIdx.x * g.tx_n + tx;
  const int img = blockIdx.z;
  const bool active = ty < g.ty_n && vcol < g.vc;
  const int rows_per_split = (nc.P + g.splits - 1) / g.splits;
  const int r0 = blockIdx.y * rows_per_split;
  const int r1 = min(nc.P, r0 + rows_per_split);
  float s[N], q[N];
#pragma unroll
  for (int j = 0; j < N; ++j) { s[j] = 0.f; q[j] = 0.f; }
  if (active) {
    const T* base = x + ((long)img * nc.P) * nc.ldx + (long)vcol * N;
    constexpr int UN = 4;
    for (int r = r0 + ty; r < r1; r += UN * g.ty_n) {
      float v[UN][N];
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const int rr = r + u * g.ty_n;
        if (rr < r1) load_vec<T, N>(base + (long)rr * nc.ldx, v[u], true);
        else {
#pragma unroll
          for (int j = 0; j < N; ++j) v[u][j] = 0.f;
        }
      }
#pragma unroll
      for (int u = 0; u < UN; ++u) {
#pragma unroll
        for (int j = 0; j < N; ++j) { s[j] += v[u][j]; q[j] = fmaf(v[u][j], v[u][j], q[j]); }
      }
    }
    float* my = red + ((long)ty * g.tx_n + tx) * 2 * N;
#pragma unroll
    for (int j = 0; j < N; ++j) { my[j] = s[j]; my[N + j] = q[j]; }
  }
  __syncthreads();
  if (active && ty == 0) {
    for (int t = 1; t < g.ty_n; ++t) {
      const float* o = red + ((long)t * g.tx_n + tx) * 2 * N;
#pragma unroll
      for (int j = 0; j < N; ++j) { s[j] += o[j]; q[j] += o[N + j]; }
    }
    float* dst = stats + ((long)img * nc.C + (long)vcol * N) * 2;
#pragma unroll
    for (int j = 0; j < N; ++j) {
      atomicAdd(dst + 2 * j, s[j]);
      atomicAdd(dst + 2 * j + 1, q[j]);
    }
  }
}

__device__ __forceinline__ void mean_rstd(const float* stats, long idx, float inv_p, float& mean, float& rstd) {
  const float s = stats[2 * idx], q = stats[2 * idx + 1];
  mean = s * inv_p;
  const float var = fmaxf(q * inv_p - mean * mean, 0.f);
  rstd = rsqrtf(var + 1e-5f);
}

// ---------------------------------------------------------------------------------------------
// apply: y = IN(x) [gelu] [film]; out = y | resid_in + row_scale*col_gamma*y
// ---------------------------------------------------------------------------------------------
struct ApplyParams {
  NormCommon nc;
  const float* stats;
  const float* weight;
  const float* bias;
  int gelu;
  const float* film_gamma;   // [I / film_T][C]
  const float* film_beta;
  int film_T;
  const float* resid_in;     // fp32 (I*P, C) ld = ldo, or null
  const float* row_scale;    // [I] or null
  const float* col_gamma;    // [C] (with resid_in)
  long ldo;
};

template <typename TI, typename TO>
__global__ void __launch_bounds__(kNormThreads)
inorm_apply_kernel(const TI* __restrict__ x, TO* __restrict__ out, ApplyParams p) {
  constexpr int N = Vec<TI>::N;       // work unit = the input vector width
  const RowGeom& g = p.nc.g;
  const int tx = threadIdx.x % g.tx_n, ty = threadIdx.x / g.tx_n;
  const int vcol = blockIdx.x * g.tx_n + tx;
  const int img = blockIdx.z;
  if (!(ty < g.ty_n && vcol < g.vc)) return;
  const int c0 = vcol * N;
  const float inv_p = 1.f / (float)p.nc.P;
  float a[N], b[N], fg[N], fb[N], cg[N];
#pragma unroll
  for (int j = 0; j < N; ++j) {
    float mean, rstd;
    mean_rstd(p.stats, (long)img * p.nc.C + c0 + j, inv_p, mean, rstd);
    const float w = p.weight[c0 + j];
    a[j] = rstd * w;
    b[j] = p.bias[c0 + j] - mean * rstd * w;
    if (p.film_gamma != nullptr) {
      const long fi = (long)(img / p.film_T) * p.nc.C + c0 + j;
      fg[j] = p.film_gamma[fi];
      fb[j] = p.film_beta[fi];
    }
    if (p.resid_in != nullptr)
      cg[j] = p.col_gamma[c0 + j] * (p.row_scale != nullptr ? p.row_scale[img] : 1.f);
  }
  const int rows_per_split = (p.nc.P + g.splits - 1) / g.splits;
  const int r0 = blockIdx.y * rows_per_split;
  const int r1 = min(p.nc.P, r0 + rows_per_split);
  const TI* xb = x + ((long)img * p.nc.P) * p.nc.ldx + c0;
  TO* ob = out + ((long)img * p.nc.P) * p.ldo + c0;
  const float* rb = p.resid_in != nullptr ? p.resid_in + ((long)img * p.nc.P) * p.ldo + c0 : nullptr;
  constexpr int UN = 2;
  for (int r = r0 + ty; r < r1; r += UN * g.ty_n) {
    float vv[UN][N];
    float xr[UN][N];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int rr = r + u * g.ty_n;
      if (rr < r1) {
        load_vec<TI, N>(xb + (long)rr * p.nc.ldx, vv[u], true);
        if (rb != nullptr) {
#pragma unroll
          for (int h = 0; h < N / 4; ++h) {
            float t4[4];
            load_vec<float, 4>(rb + (long)rr * p.ldo + 4 * h, t4, true);
#pragma unroll
            for (int j = 0; j < 4; ++j) xr[u][4 * h + j] = t4[j];
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int rr = r + u * g.ty_n;
      if (rr >= r1) continue;
      float (&v)[N] = vv[u];
#pragma unroll
      for (int j = 0; j < N; ++j) {
        float y = fmaf(v[j], a[j], b[j]);
        if (p.gelu) y = gelu_erf(y);
        if (p.film_gamma != nullptr) y = fmaf(fg[j], y, fb[j]);
        if (rb != nullptr) y = fmaf(cg[j], y, xr[u][j]);
        v[j] = y;
      }
      if constexpr (sizeof(TO) == 4 && N == 8) {
        float lo[4] = {v[0], v[1], v[2], v[3]}, hi[4] = {v[4], v[5], v[6], v[7]};
        store_vec<float, 4>(reinterpret_cast<float*>(ob + (long)rr * p.ldo), lo, true);
        store_vec<float, 4>(reinterpret_cast<float*>(ob + (long)rr * p.ldo) + 4, hi, true);
      } else if constexpr (sizeof(TO) == 2 && N == 4) {
        uint2 u2;
        u2.x = pack2<TO>(v[0], v[1]); u2.y = pack2<TO>(v[2], v[3]);
        *reinterpret_cast<uint2*>(ob + (long)rr * p.ldo) = u2;
      } else {
        store_vec<TO, N>(ob + (long)rr * p.ldo, v, true);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// backward, pass 1: red[img][c] += (sum g, sum g * xhat), g = gin [* gelu'(xhat*w+b)]
// ---------------------------------------------------------------------------------------------
struct BwdParams {
  NormCommon nc;
  const float* stats;
  const float* weight;
  const float* bias;
  int gelu;
  long ldg;
  float* red;                // [I][C][2]
  // pass 2 only
  const float* row_scale;    // [I] or null
  const float* col_scale;    // [C] or null
  const float* film_gamma;   // [I / film_T][C] or null
  int film_T;
  const float* add32;        // fp32 tensor added to the result (residual-stream gradient), ld = ldo
  long ldo;
};

template <typename TG, typename TX>
__global__ void __launch_bounds__(kNormThreads)
inorm_bwd_reduce_kernel(const TG* __restrict__ gin, const TX* __restrict__ x, BwdParams p) {
  constexpr int N = 8;                 // 8 channels per thread for every dtype combination
  constexpr int UN = 2;
  extern __shared__ float red[];       // [ty_n][tx_n][2N]
  const RowGeom& g = p.nc.g;
  const int tx = threadIdx.x % g.tx_n, ty = threadIdx.x / g.tx_n;
  const int vcol = blockIdx.x * g.tx_n + tx;
  const int img = blockIdx.z;
  const bool active = ty < g.ty_n && vcol < g.vc;
  const int c0 = vcol * N;
  const float inv_p = 1.f / (float)p.nc.P;
  float s[N], q[N], mean[N], rstd[N], w[N], b[N];
#pragma unroll
  for (int j = 0; j < N; ++j) { s[j] = 0.f; q[j] = 0.f; }
  if (active) {
#pragma unroll
    for (int j = 0; j < N; ++j) {
      mean_rstd(p.stats, (long)img * p.nc.C + c0 + j, inv_p, mean[j], rstd[j]);
      w[j] = p.weight[c0 + j];
      b[j] = p.bias[c0 + j];
    }
    const int rows_per_split = (p.nc.P + g.splits - 1) / g.splits;
    const int r0 = blockIdx.y * rows_per_split;
    const int r1 = min(p.nc.P, r0 + rows_per_split);
    const TG* gb = gin + ((long)img * p.nc.P) * p.ldg + c0;
    const TX* xb = x + ((long)img * p.nc.P) * p.nc.ldx + c0;
    for (int r = r0 + ty; r < r1; r += UN * g.ty_n) {
      float gv[UN][N], xv[UN][N];
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const int rr = r + u * g.ty_n;
        if (rr < r1) {
          load8<TG>(gb + (long)rr * p.ldg, gv[u]);
          load8<TX>(xb + (long)rr * p.nc.ldx, xv[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        if (r + u * g.ty_n >= r1) continue;
#pragma unroll
        for (int j = 0; j < N; ++j) {
          const float xh = (xv[u][j] - mean[j]) * rstd[j];
          float gg = gv[u][j];
          if (p.gelu) gg *= gelu_erf_grad(fmaf(xh, w[j], b[j]));
          s[j] += gg;
          q[j] = fmaf(gg, xh, q[j]);
        }
      }
    }
    float* my = red + ((long)ty * g.tx_n + tx) * 2 * N;
#pragma unroll
    for (int j = 0; j < N; ++j) { my[j] = s[j]; my[N + j] = q[j]; }
  }
  __syncthreads();
  if (active && ty == 0) {
    for (int t = 1; t < g.ty_n; ++t) {
      const float* o = red + ((long)t * g.tx_n + tx) * 2 * N;
#pragma unroll
      for (int j = 0; j < N; ++j) { s[j] += o[j]; q[j] += o[N + j]; }
    }
    float* dst = p.red + ((long)img * p.nc.C + c0) * 2;
#pragma unroll
    for (int j = 0; j < N; ++j) {
      atomicAdd(dst + 2 * j, s[j]);
      atomicAdd(dst + 2 * j + 1, q[j]);
    }
  }
}

// backward, pass 2: dx = rstd*w*cs*(g - R1/P - xhat*R2/P) [+ add32]
template <typename TG, typename TX, typename TO>
__global__ void __launch_bounds__(kNormThreads)
inorm_bwd_apply_kernel(const TG* __restrict__ gin, const TX* __restrict__ x, TO* __restrict__ out, BwdParams p) {
  constexpr int N = 8;
  constexpr int UN = 2;
  const RowGeom& g = p.nc.g;
  const int tx = threadIdx.x % g.tx_n, ty = threadIdx.x / g.tx_n;
  const int vcol = blockIdx.x * g.tx_n + tx;
  const int img = blockIdx.z;
  if (!(ty < g.ty_n && vcol < g.vc)) return;
  const int c0 = vcol * N;
  const float inv_p = 1.f / (float)p.nc.P;
  float mean[N], rstd[N], w[N], b[N], k[N], m1[N], m2[N];
#pragma unroll
  for (int j = 0; j < N; ++j) {
    const long idx = (long)img * p.nc.C + c0 + j;
    mean_rstd(p.stats, idx, inv_p, mean[j], rstd[j]);
    w[j] = p.weight[c0 + j];
    b[j] = p.bias[c0 + j];
    float cs = 1.f;
    if (p.row_scale != nullptr) cs *= p.row_scale[img];
    if (p.col_scale != nullptr) cs *= p.col_scale[c0 + j];
    if (p.film_gamma != nullptr) cs *= p.film_gamma[(long)(img / p.film_T) * p.nc.C + c0 + j];
    k[j] = rstd[j] * w[j] * cs;
    m1[j] = p.red[2 * idx] * inv_p;
    m2[j] = p.red[2 * idx + 1] * inv_p;
  }
  const int rows_per_split = (p.nc.P + g.splits - 1) / g.splits;
  const int r0 = blockIdx.y * rows_per_split;
  const int r1 = min(p.nc.P, r0 + rows_per_split);
  const TG* gb = gin + ((long)img * p.nc.P) * p.ldg + c0;
  const TX* xb = x + ((long)img * p.nc.P) * p.nc.ldx + c0;
  TO* ob = out + ((long)img * p.nc.P) * p.ldo + c0;
  const float* ab = p.add32 != nullptr ? p.add32 + ((long)img * p.nc.P) * p.ldo + c0 : nullptr;
  for (int r = r0 + ty; r < r1; r += UN * g.ty_n) {
    float gv[UN][N], xv[UN][N], av[UN][N];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int rr = r + u * g.ty_n;
      if (rr < r1) {
        load8<TG>(gb + (long)rr * p.ldg, gv[u]);
        load8<TX>(xb + (long)rr * p.nc.ldx, xv[u]);
        if (ab != nullptr) load8<float>(ab + (long)rr * p.ldo, av[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int rr = r + u * g.ty_n;
      if (rr >= r1) continue;
      float o[N];
#pragma unroll
      for (int j = 0; j < N; ++j) {
        const float xh = (xv[u][j] - mean[j]) * rstd[j];
        float gg = gv[u][j];
        if (p.gelu) gg *= gelu_erf_grad(fmaf(xh, w[j], b[j]));
        o[j] = k[j] * (gg - m1[j] - xh * m2[j]);
        if (ab != nullptr) o[j] += av[u][j];
      }
      store8<TO>(ob + (long)rr * p.ldo, o);
    }
  }
}

// parameter gradients from the per-(image, channel) reductions: 32 channels x 8 image lanes per block
struct BwdParamArgs {
  const float* red; const float* stats;
  int I, P, C;
  const float* row_scale; const float* col_scale; const float* film_gamma; int film_T;
  const float* weight; const float* bias;
  float* dweight; float* dbias; float* dcol_scale; float* dfilm_gamma; float* dfilm_beta;
};
__global__ void __launch_bounds__(256) inorm_bwd_params_kernel(BwdParamArgs a) {
  __shared__ float sh[3][8][33];
  const int cl = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  float dw = 0.f, db = 0.f, dcs = 0.f;
  float w = 0.f, b = 0.f;
  if (c < a.C) {
    w = a.weight[c]; b = a.bias[c];
    for (int img = grp; img < a.I; img += 8) {
      const long idx = (long)img * a.C + c;
      const float R1 = a.red[2 * idx], R2 = a.red[2 * idx + 1];
      const float rs = a.row_scale != nullptr ? a.row_scale[img] : 1.f;
      float cs = rs;
      if (a.col_scale != nullptr) cs *= a.col_scale[c];
      if (a.film_gamma != nullptr) cs *= a.film_gamma[(long)(img / a.film_T) * a.C + c];
      dw = fmaf(cs, R2, dw);
      db = fmaf(cs, R1, db);
      dcs = fmaf(rs, fmaf(w, R2, b * R1), dcs);
    }
  }
  sh[0][grp][cl] = dw; sh[1][grp][cl] = db; sh[2][grp][cl] = dcs;
  __syncthreads();
  if (grp == 0 && c < a.C) {
    for (int t = 1; t < 8; ++t) { dw += sh[0][t][cl]; db += sh[1][t][cl]; dcs += sh[2][t][cl]; }
    if (a.dweight) a.dweight[c] += dw;
    if (a.dbias) a.dbias[c] += db;
    if (a.dcol_scale) a.dcol_scale[c] += dcs;
    if (a.dfilm_gamma) {
      const int nb = a.I / a.film_T;
      for (int bi = 0; bi < nb; ++bi) {
        float dg = 0.f, dbt = 0.f;
        for (int t = 0; t < a.film_T; ++t) {
          const long idx = (long)(bi * a.film_T + t) * a.C + c;
          const float R1 = a.red[2 * idx], R2 = a.red[2 * idx + 1];
          dg += fmaf(w, R2, b * R1);
          dbt += R1;
        }
        a.dfilm_gamma[(long)bi * a.C + c] = dg;
        a.dfilm_beta[(long)bi * a.C + c] = dbt;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// residual-branch backward helper (attention.py:123,309 reversed):
//   dz16 = row_scale[img] * coef[c] * dx ;  S0[c] += sum rs*dx ;  S1[c] += sum rs*dx*z
// ---------------------------------------------------------------------------------------------
struct ResidBwdParams {
  NormCommon nc;             // ldx = ld of dx32
  const float* dx;
  const void* z16;           // (I*P, C) 16-bit, ld = ldz, may be null (then S1 untouched)
  long ldz;
  const float* row_scale;    // [I] or null
  const float* coef;         // [C]
  void* dz16;                // out, ld = ldz, may be null
  float* S0; float* S1;      // [C]
};
template <typename T16>
__global__ void __launch_bounds__(kNormThreads)
resid_bwd_kernel(ResidBwdParams p) {
  constexpr int N = 8;
  constexpr int UN = 2;
  extern __shared__ float red[];
  const RowGeom& g = p.nc.g;
  const int tx = threadIdx.x % g.tx_n, ty = threadIdx.x / g.tx_n;
  const int vcol = blockIdx.x * g.tx_n + tx;
  const int img = blockIdx.z;
  const bool active = ty < g.ty_n && vcol < g.vc;
  const int c0 = vcol * N;
  float s[N], q[N];
#pragma unroll
  for (int j = 0; j < N; ++j) { s[j] = 0.f; q[j] = 0.f; }
  if (active) {
    const float rs = p.row_scale != nullptr ? p.row_scale[img] : 1.f;
    float cf[N];
#pragma unroll
    for (int j = 0; j < N; ++j) cf[j] = p.coef[c0 + j] * rs;
    const int rows_per_split = (p.nc.P + g.splits - 1) / g.splits;
    const int r0 = blockIdx.y * rows_per_split;
    const int r1 = min(p.nc.P, r0 + rows_per_split);
    const float* db = p.dx + ((long)img * p.nc.P) * p.nc.ldx + c0;
    const T16* zb = p.z16 ? reinterpret_cast<const T16*>(p.z16) + ((long)img * p.nc.P) * p.ldz + c0 : nullptr;
    T16* ob = p.dz16 ? reinterpret_cast<T16*>(p.dz16) + ((long)img * p.nc.P) * p.ldz + c0 : nullptr;
    for (int r = r0 + ty; r < r1; r += UN * g.ty_n) {
      float dv[UN][N], zv[UN][N];
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const int rr = r + u * g.ty_n;
        if (rr < r1) {
          load8<float>(db + (long)rr * p.nc.ldx, dv[u]);
          if (zb != nullptr) load8<T16>(zb + (long)rr * p.ldz, zv[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const int rr = r + u * g.ty_n;
        if (rr >= r1) continue;
        float o[N];
#pragma unroll
        for (int j = 0; j < N; ++j) {
          if (zb != nullptr) q[j] = fmaf(rs * dv[u][j], zv[u][j], q[j]);
          s[j] = fmaf(rs, dv[u][j], s[j]);
          o[j] = cf[j] * dv[u][j];
        }
        if (ob != nullptr) store8<T16>(ob + (long)rr * p.ldz, o);
      }
    }
    float* my = red + ((long)ty * g.tx_n + tx) * 2 * N;
#pragma unroll
    for (int j = 0; j < N; ++j) { my[j] = s[j]; my[N + j] = q[j]; }
  }
  __syncthreads();
  if (active && ty == 0) {
    for (int t = 1; t < g.ty_n; ++t) {
      const float* o = red + ((long)t * g.tx_n + tx) * 2 * N;
#pragma unroll
      for (int j = 0; j < N; ++j) { s[j] += o[j]; q[j] += o[N + j]; }
    }
#pragma unroll
    for (int j = 0; j < N; ++j) {
      atomicAdd(p.S0 + c0 + j, s[j]);
      if (p.z16 != nullptr) atomicAdd(p.S1 + c0 + j, q[j]);
    }
  }
}

// column sums of a 16-bit matrix: out[c] += sum_rows x[r, c]   (bias gradients)
template <typename T16>
__global__ void __launch_bounds__(kNormThreads)
colsum16_kernel(const T16* __restrict__ x, NormCommon nc, float* __restrict__ out) {
  constexpr int N = 8;
  extern __shared__ float red[];
  const RowGeom& g = nc.g;
  const int tx = threadIdx.x % g.tx_n, ty = threadIdx.x / g.tx_n;
  const int vcol = blockIdx.x * g.tx_n + tx;
  const bool active = ty < g.ty_n && vcol < g.vc;
  float s[N];
#pragma unroll
  for (int j = 0; j < N; ++j) s[j] = 0.f;
  if (active) {
    const int rows_per_split = (nc.P + g.splits - 1) / g.splits;
    const int r0 = blockIdx.y * rows_per_split;
    const int r1 = min(nc.P, r0 + rows_per_split);
    const T16* base = x + (long)vcol * N;
    constexpr int UN = 4;
    for (int r = r0 + ty; r < r1; r += UN * g.ty_n) {
      float v[UN][N];
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const int rr = r + u * g.ty_n;
        if (rr < r1) load_vec<T16, N>(base + (long)rr * nc.ldx, v[u], true);
        else {
#pragma unroll
          for (int j = 0; j < N; ++j) v[u][j] = 0.f;
        }
      }
#pragma unroll
      for (int u = 0; u < UN; ++u) {
#pragma unroll
        for (int j = 0; j < N; ++j) s[j] += v[u][j];
      }
    }
    float* my = red + ((long)ty * g.tx_n + tx) * N;
#pragma unroll
    for (int j = 0; j < N; ++j) my[j] = s[j];
  }
  __syncthreads();
  if (active && ty == 0) {
    for (int t = 1; t < g.ty_n; ++t) {
      const float* o = red + ((long)t * g.tx_n + tx) * N;
#pragma unroll
      for (int j = 0; j < N; ++j) s[j] += o[j];
    }
#pragma unroll
    for (int j = 0; j < N; ++j) atomicAdd(out + (long)vcol * N + j, s[j]);
  }
}

static int check_common(const char* fn, int I, int P, int C, long ld, int vecn, const void* x) {
  BF_REQUIRE(I > 0 && P > 0 && C > 0, "%s: empty problem I=%d P=%d C=%d", fn, I, P, C);
  BF_REQUIRE(C % vecn == 0, "%s: C=%d must be a multiple of %d", fn, C, vecn);
  BF_REQUIRE(ld >= C && ld % vecn == 0, "%s: ld=%ld", fn, ld);
  BF_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "%s: pointer must be 16-byte aligned", fn);
  return BF_OK;
}

}  // namespace bf

using namespace bf;

static inline int dtype_vec(int dt) { return dt == BF_F32 ? 4 : 8; }

extern "C" int bf_inorm_stats(const void* x, int x_dtype, int I, int P, int C, int64_t ldx, float* stats,
                              void* stream) {
  BF_REQUIRE(x && stats, "bf_inorm_stats: null pointer");
  BF_REQUIRE(x_dtype == BF_BF16 || x_dtype == BF_F16 || x_dtype == BF_F32, "bf_inorm_stats: dtype %d", x_dtype);
  const int vn = dtype_vec(x_dtype);
  if (int st = check_common("bf_inorm_stats", I, P, C, ldx, vn, x)) return st;
  NormCommon nc{I, P, C, ldx, make_geom(C, vn, P, I)};
  dim3 grid(nc.g.chunks, nc.g.splits, I);
  const size_t sm = (size_t)nc.g.ty_n * nc.g.tx_n * 2 * vn * sizeof(float);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (x_dtype == BF_F32) inorm_stats_kernel<float><<<grid, kNormThreads, sm, s>>>((const float*)x, nc, stats);
  else if (x_dtype == BF_BF16) inorm_stats_kernel<__nv_bfloat16><<<grid, kNormThreads, sm, s>>>((const __nv_bfloat16*)x, nc, stats);
  else inorm_stats_kernel<__half><<<grid, kNormThreads, sm, s>>>((const __half*)x, nc, stats);
  count_launch();
  BF_LAUNCH_CHECK("inorm_stats_kernel");
  return BF_OK;
}

extern "C" int bf_inorm_apply(const bf_inorm_apply_args* a, void* stream) {
  BF_REQUIRE(a && a->x && a->out && a->stats && a->weight && a->bias, "bf_inorm_apply: null pointer");
  const int vn = dtype_vec(a->x_dtype);
  if (int st = check_common("bf_inorm_apply", a->I, a->P, a->C, a->ldx, vn, a->x)) return st;
  BF_REQUIRE(a->ldo >= a->C && a->ldo % vn == 0, "bf_inorm_apply: ldo");
  BF_REQUIRE((a->film_gamma == nullptr) == (a->film_beta == nullptr), "bf_inorm_apply: film pair");
  BF_REQUIRE(a->film_gamma == nullptr || (a->film_T > 0 && a->I % a->film_T == 0), "bf_inorm_apply: film_T");
  BF_REQUIRE(a->resid_in == nullptr || (a->out_dtype == BF_F32 && a->col_gamma), "bf_inorm_apply: residual needs f32 out");
  ApplyParams p{};
  p.nc = NormCommon{a->I, a->P, a->C, a->ldx, make_geom(a->C, vn, a->P, a->I)};
  p.stats = a->stats; p.weight = a->weight; p.bias = a->bias; p.gelu = a->gelu;
  p.film_gamma = a->film_gamma; p.film_beta = a->film_beta; p.film_T = a->film_T > 0 ? a->film_T : 1;
  p.resid_in = a->resid_in; p.row_scale = a->row_scale; p.col_gamma = a->col_gamma; p.ldo = a->ldo;
  dim3 grid(p.nc.g.chunks, p.nc.g.splits, a->I);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
#define BF_APPLY(TI, TO) inorm_apply_kernel<TI, TO><<<grid, kNormThreads, 0, s>>>((const TI*)a->x, (TO*)a->out, p)
  const int xi = a->x_dtype, xo = a->out_dtype;
  if (xi == BF_F32 && xo == BF_BF16) BF_APPLY(float, __nv_bfloat16);
  else if (xi == BF_F32 && xo == BF_F16) BF_APPLY(float, __half);
  else if (xi == BF_F32 && xo == BF_F32) BF_APPLY(float, float);
  else if (xi == BF_BF16 && xo == BF_BF16) BF_APPLY(__nv_bfloat16, __nv_bfloat16);
  else if (xi == BF_BF16 && xo == BF_F32) BF_APPLY(__nv_bfloat16, float);
  else if (xi == BF_F16 && xo == BF_F16) BF_APPLY(__half, __half);
  else if (xi == BF_F16 && xo == BF_F32) BF_APPLY(__half, float);
  else BF_REQUIRE(false, "bf_inorm_apply: unsupported dtype pair %d -> %d", xi, xo);
#undef BF_APPLY
  count_launch();
  BF_LAUNCH_CHECK("inorm_apply_kernel");
  return BF_OK;
}

extern "C" int bf_inorm_bwd(const bf_inorm_bwd_args* a, void* stream) {
  BF_REQUIRE(a && a->gin && a->x && a->stats && a->weight && a->bias && a->red, "bf_inorm_bwd: null pointer");
  if (int st = check_common("bf_inorm_bwd", a->I, a->P, a->C, a->ldx, 8, a->x)) return st;
  BF_REQUIRE(a->ldg >= a->C && a->ldg % 8 == 0, "bf_inorm_bwd: ldg");
  BF_REQUIRE(a->phase == 1 || a->phase == 2, "bf_inorm_bwd: phase %d", a->phase);
  BwdParams p{};
  p.nc = NormCommon{a->I, a->P, a->C, a->ldx, make_geom(a->C, 8, a->P, a->I)};
  p.stats = a->stats; p.weight = a->weight; p.bias = a->bias; p.gelu = a->gelu; p.ldg = a->ldg; p.red = a->red;
  p.row_scale = a->row_scale; p.col_scale = a->col_scale; p.film_gamma = a->film_gamma;
  p.film_T = a->film_T > 0 ? a->film_T : 1; p.add32 = a->add32; p.ldo = a->ldo;
  dim3 grid(p.nc.g.chunks, p.nc.g.splits, a->I);
  const size_t sm = (size_t)p.nc.g.ty_n * p.nc.g.tx_n * 16 * sizeof(float);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int gd = a->g_dtype, xd = a->x_dtype, od = a->out_dtype;
  if (a->phase == 1) {
#define BF_RED(TG, TX) inorm_bwd_reduce_kernel<TG, TX><<<grid, kNormThreads, sm, s>>>((const TG*)a->gin, (const TX*)a->x, p)
    if (gd == BF_F32 && xd == BF_F32) BF_RED(float, float);
    else if (gd == BF_F32 && xd == BF_BF16) BF_RED(float, __nv_bfloat16);
    else if (gd == BF_F32 && xd == BF_F16) BF_RED(float, __half);
    else if (gd == BF_BF16 && xd == BF_F32) BF_RED(__nv_bfloat16, float);
    else if (gd == BF_BF16 && xd == BF_BF16) BF_RED(__nv_bfloat16, __nv_bfloat16);
    else if (gd == BF_F16 && xd == BF_F16) BF_RED(__half, __half);
    else if (gd == BF_BF16 && xd == BF_F16) BF_RED(__nv_bfloat16, __half);
    else BF_REQUIRE(false, "bf_inorm_bwd: unsupported dtype pair g=%d x=%d", gd, xd);
#undef BF_RED
    count_launch();
    BF_LAUNCH_CHECK("inorm_bwd_reduce_kernel");
    return BF_OK;
  }
  BF_REQUIRE(a->out, "bf_inorm_bwd: out required in phase 2");
  BF_REQUIRE(a->ldo >= a->C && a->ldo % 8 == 0, "bf_inorm_bwd: ldo");
  BF_REQUIRE(a->add32 == nullptr || od == BF_F32, "bf_inorm_bwd: add32 needs f32 out");
#define BF_APP(TG, TX, TO) inorm_bwd_apply_kernel<TG, TX, TO><<<grid, kNormThreads, 0, s>>>((const TG*)a->gin, (const TX*)a->x, (TO*)a->out, p)
  if (gd == BF_F32 && xd == BF_F32 && od == BF_F32) BF_APP(float, float, float);
  else if (gd == BF_F32 && xd == BF_BF16 && od == BF_BF16) BF_APP(float, __nv_bfloat16, __nv_bfloat16);
  else if (gd == BF_F32 && xd == BF_F16 && od == BF_F16) BF_APP(float, __half, __half);
  else if (gd == BF_BF16 && xd == BF_F32 && od == BF_F32) BF_APP(__nv_bfloat16, float, float);
  else if (gd == BF_BF16 && xd == BF_BF16 && od == BF_BF16) BF_APP(__nv_bfloat16, __nv_bfloat16, __nv_bfloat16);
  else if (gd == BF_F16 && xd == BF_F16 && od == BF_F16) BF_APP(__half, __half, __half);
  else if (gd == BF_BF16 && xd == BF_F16 && od == BF_BF16) BF_APP(__nv_bfloat16, __half, __nv_bfloat16);
  else if (gd == BF_F32 && xd == BF_F16 && od == BF_BF16) BF_APP(float, __half, __nv_bfloat16);
  else BF_REQUIRE(false, "bf_inorm_bwd: unsupported dtype triple g=%d x=%d out=%d", gd, xd, od);
#undef BF_APP
  count_launch();
  BF_LAUNCH_CHECK("inorm_bwd_apply_kernel");
  return BF_OK;
}

extern "C" int bf_inorm_bwd_params(const bf_inorm_bwd_params_args* a, void* stream) {
  BF_REQUIRE(a && a->red && a->weight && a->bias, "bf_inorm_bwd_params: null pointer");
  BF_REQUIRE(a->I > 0 && a->C > 0, "bf_inorm_bwd_params: empty");
  BF_REQUIRE((a->dfilm_gamma == nullptr) == (a->dfilm_beta == nullptr), "bf_inorm_bwd_params: film pair");
  BF_REQUIRE(a->dfilm_gamma == nullptr || (a->film_T > 0 && a->I % a->film_T == 0), "bf_inorm_bwd_params: film_T");
  BwdParamArgs k{a->red, nullptr, a->I, a->P, a->C, a->row_scale, a->col_scale, a->film_gamma,
                 a->film_T > 0 ? a->film_T : 1, a->weight, a->bias, a->dweight, a->dbias, a->dcol_scale,
                 a->dfilm_gamma, a->dfilm_beta};
  inorm_bwd_params_kernel<<<(a->C + 31) / 32, 256, 0, static_cast<cudaStream_t>(stream)>>>(k);
  count_launch();
  BF_LAUNCH_CHECK("inorm_bwd_params_kernel");
  return BF_OK;
}

extern "C" int bf_resid_bwd(const float* dx, int64_t lddx, const void* z16, void* dz16, int64_t ldz, int dtype,
                            int I, int P, int C, const float* row_scale, const float* coef, float* S0, float* S1,
                            void* stream) {
  BF_REQUIRE(dx && coef && S0, "bf_resid_bwd: null pointer");
  BF_REQUIRE(z16 == nullptr || S1 != nullptr, "bf_resid_bwd: S1 required with z16");
  BF_REQUIRE(dtype == BF_BF16 || dtype == BF_F16, "bf_resid_bwd: dtype");
  if (int st = check_common("bf_resid_bwd", I, P, C, lddx, 8, dx)) return st;
  BF_REQUIRE((z16 == nullptr && dz16 == nullptr) || (ldz >= C && ldz % 8 == 0), "bf_resid_bwd: ldz");
  ResidBwdParams p{};
  p.nc = NormCommon{I, P, C, lddx, make_geom(C, 8, P, I)};
  p.dx = dx; p.z16 = z16; p.ldz = ldz; p.row_scale = row_scale; p.coef = coef; p.dz16 = dz16; p.S0 = S0; p.S1 = S1;
  dim3 grid(p.nc.g.chunks, p.nc.g.splits, I);
  const size_t sm = (size_t)p.nc.g.ty_n * p.nc.g.tx_n * 16 * sizeof(float);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dtype == BF_BF16) resid_bwd_kernel<__nv_bfloat16><<<grid, kNormThreads, sm, s>>>(p);
  else resid_bwd_kernel<__half><<<grid, kNormThreads, sm, s>>>(p);
  count_launch();
  BF_LAUNCH_CHECK("resid_bwd_kernel");
  return BF_OK;
}

extern "C" int bf_colsum16(const void* x, int dtype, int64_t rows, int C, int64_t ldx, float* out, void* stream) {
  BF_REQUIRE(x && out, "bf_colsum16: null pointer");
  BF_REQUIRE(dtype == BF_BF16 || dtype == BF_F16, "bf_colsum16: dtype");
  BF_REQUIRE(rows > 0 && rows < (1ll << 31), "bf_colsum16: rows");
  if (int st = check_common("bf_colsum16", 1, (int)rows, C, ldx, 8, x)) return st;
  NormCommon nc{1, (int)rows, C, ldx, make_geom(C, 8, (int)rows, 1)};
  dim3 grid(nc.g.chunks, nc.g.splits, 1);
  const size_t sm = (size_t)nc.g.ty_n * nc.g.tx_n * 8 * sizeof(float);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dtype == BF_BF16) colsum16_kernel<__nv_bfloat16><<<grid, kNormThreads, sm, s>>>((const __nv_bfloat16*)x, nc, out);
  else colsum16_kernel<__half><<<grid, kNormThreads, sm, s>>>((const __half*)x, nc, out);
  count_launch();
  BF_LAUNCH_CHECK("colsum16_kernel");
  return BF_OK;
}
