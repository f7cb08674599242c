// InstanceNorm over the token grid of one image (per image, per channel), token-major data.
//
// Replaces nn.InstanceNorm2d(affine=True) at upstream layers/attention.py:39-40,76,120,153-154,197,208,298,316
// and layers/patching.py:45,102 (forward and autograd backward), fused with what follows it:
// GELU (patching.py:47,103), FiLM (linear_layers.py:71-77), layer-scale * drop-path + residual
// (attention.py:317).  Also the residual-branch backward helper and column sums for bias gradients.
//
// All kernels are HBM-bound streaming passes built on one pattern (measured against the per-thread cp.async ring it
// replaces -- scripts/exp/stream_exp.cu, profiles/r2o_stream_exp.txt: 30.2 -> 18.6 us for the fp32 -> bf16 apply pass and
// 58.6 -> 36.2 us for the norm1 backward apply at config 2, i.e. 0.78 / 0.92 of the measured HBM copy rate):
//   * a block owns a slab of rows of one image (x one column chunk); a dedicated producer warp moves that slab through
//     a shared-memory ring with TMA 1-D bulk copies (cp.async.bulk, full lines, one instruction per operand and stage
//     when the rows are contiguous, one per row otherwise), completion on a per-stage mbarrier;
//   * the 256 consumer threads each own 8 consecutive channels: they read their 16 / 32 bytes per row from shared memory,
//     keep the per-(image, channel) coefficients in registers and store results straight to global memory (128- or
//     256-bit stores), then hand the stage back through a second mbarrier -- no block-level synchronisation in the loop;
//   * the grid is one resident wave (2 blocks per SM, 96 KiB of ring each); per-(image, channel) partial sums are combined
//     in shared memory and flushed with one fp32 atomic per block.
// Statistics are kept as raw sums (sum x, sum x^2) per (image, channel) so producers can accumulate them.
#include "common.cuh"

namespace bf {

constexpr int kNT = 256;                 // consumer threads per block
constexpr int kThreads = kNT + 32;       // + the producer warp (warp 8)
constexpr int kRingBytes = 96 * 1024;    // bulk-copy ring per block
constexpr int kMaxStages = 8;
constexpr int kSmemBytes = kRingBytes + 2 * kMaxStages * 8;
// fused two-phase kernels (thread-block clusters): a smaller ring plus this block's per-channel partial sums
constexpr int kFusedRingBytes = 80 * 1024;
constexpr int kFusedMaxC = 1024;
constexpr int kFusedSmemBytes = kRingBytes + 2 * kMaxStages * 8;   // ring (88 KiB) + partials (8 KiB) + barriers
static_assert(kFusedRingBytes + 2 * kFusedMaxC * 2 * 4 <= kRingBytes, "fused layout");   // partials + totals
constexpr int kBlocksPerSM = 2;
constexpr int kStageTarget = 20 * 1024;  // bytes per stage aimed for (all operands)

__device__ __forceinline__ void bulk_g2s(void* smem, const void* gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem)), "l"(gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// this thread's 8 channels of one row of an operand staged in shared memory
template <typename T>
__device__ __forceinline__ void ld8(const uint8_t* p, float (&v)[8]) {
  if constexpr (sizeof(T) == 4) {
    const float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const float2 a = unpack2<T>(u.x), b = unpack2<T>(u.y), c = unpack2<T>(u.z), d = unpack2<T>(u.w);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
  }
}
// 8 results to global memory: one 128-bit store (16-bit types), one 256-bit store (fp32, 32-byte aligned rows) or two
// 128-bit stores (fp32 at 16-byte alignment; `wide` is block uniform)
template <typename T>
__device__ __forceinline__ void store8(T* p, const float (&v)[8], bool wide) {
  if constexpr (sizeof(T) == 4) {
    if (wide) {
      asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]),
                   "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
    } else {
      reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
      reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
  } else {
    uint4 u;
    u.x = pack2<T>(v[0], v[1]); u.y = pack2<T>(v[2], v[3]); u.z = pack2<T>(v[4], v[5]); u.w = pack2<T>(v[6], v[7]);
    *reinterpret_cast<uint4*>(p) = u;
  }
}

// N consecutive floats added to global memory: 128-bit vector reductions (red.global.add.v4.f32: a quarter of the L2
// atomic operations -- the column sums of one launch all land on the same few addresses) when the destination is
// 16-byte aligned, scalar atomics otherwise
template <int N>
__device__ __forceinline__ void red_add(float* dst, const float (&v)[N]) {
  static_assert(N % 4 == 0, "red_add");
  if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
    for (int j = 0; j < N; j += 4)
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(v[j]), "f"(v[j + 1]), "f"(v[j + 2]),
                   "f"(v[j + 3]) : "memory");
  } else {
#pragma unroll
    for (int j = 0; j < N; ++j) atomicAdd(dst + j, v[j]);
  }
}

// Geometry shared by all kernels: block = tx_n vector columns x ty_n rows of consumers; grid = (splits, images, column
// chunks); a stage of the ring holds rb = kk * ty_n rows of every input operand
struct Geom {
  int I, P, C;
  int vc, tx_n, ty_n, chunks, splits, rows_per_split;
  int kk, rb, ns, stage_bytes;
};
// `sum_es`: bytes per element summed over the streamed input operands
static Geom make_geom(int I, int P, int C, int sum_es, int ring_bytes = kRingBytes, int max_splits_cap = 1 << 30) {
  Geom g;
  g.I = I; g.P = P; g.C = C;
  g.vc = C / 8;
  // Fewest column chunks: whole rows (ld == C) are contiguous, so a stage of an operand is ONE bulk copy; splitting the
  // columns to keep more consumer threads busy would turn it into one copy per row (measured on the 1152-column sums:
  // 35 vs 27 us), and these passes are bound by the copies, not by the consumers.
  g.chunks = (g.vc + kNT - 1) / kNT;
  g.tx_n = (g.vc + g.chunks - 1) / g.chunks;
  g.ty_n = kNT / g.tx_n;
  const int target = num_sms() * kBlocksPerSM;                  // one resident wave
  int splits = target / (I * g.chunks);
  const int max_splits = (P + g.ty_n - 1) / g.ty_n;             // at least one row per thread row
  if (splits > max_splits) splits = max_splits;
  if (splits > max_splits_cap) splits = max_splits_cap;
  if (splits < 1) splits = 1;
  g.splits = splits;
  g.rows_per_split = (P + splits - 1) / splits;
  const int group_bytes = g.ty_n * g.tx_n * 8 * sum_es;         // one row per consumer thread row
  int kk = kStageTarget / group_bytes;
  const int kk_max = (g.rows_per_split + g.ty_n - 1) / g.ty_n;
  if (kk > kk_max) kk = kk_max;
  if (kk < 1) kk = 1;
  g.kk = kk;
  g.rb = kk * g.ty_n;
  g.stage_bytes = g.rb * g.tx_n * 8 * sum_es;
  int ns = ring_bytes / g.stage_bytes;
  if (ns > kMaxStages) ns = kMaxStages;
  g.ns = ns;
  return g;
}

struct Ctx {            // per-thread view of the geometry
  int tx, ty, vcol, img, r0, r1;
  bool active;
};
__device__ __forceinline__ Ctx make_ctx(const Geom& g) {
  Ctx c;
  c.tx = threadIdx.x % g.tx_n;
  c.ty = threadIdx.x / g.tx_n;
  c.vcol = blockIdx.z * g.tx_n + c.tx;
  c.img = blockIdx.y;
  c.r0 = blockIdx.x * g.rows_per_split;
  c.r1 = min(g.P, c.r0 + g.rows_per_split);
  c.active = threadIdx.x < kNT && c.ty < g.ty_n && c.vcol < g.vc;
  return c;
}

// block reduction of per-thread column partials v[NV] over ty, result in ty == 0 threads.  `red` >= ty_n*tx_n*NV floats.
template <int NV>
__device__ __forceinline__ void reduce_over_ty(float (&v)[NV], float* red, const Geom& g, const Ctx& c) {
  __syncthreads();                                  // the ring is dead: reuse it
  if (c.active) {
    float* my = red + ((long)c.ty * g.tx_n + c.tx) * NV;
#pragma unroll
    for (int j = 0; j < NV; ++j) my[j] = v[j];
  }
  __syncthreads();
  if (c.active && c.ty == 0) {
    for (int t = 1; t < g.ty_n; ++t) {
      const float* o = red + ((long)t * g.tx_n + c.tx) * NV;
#pragma unroll
      for (int j = 0; j < NV; ++j) v[j] += o[j];
    }
  }
}

__device__ __forceinline__ void mean_rstd(const float* stats, long idx, float inv_p, float& mean, float& rstd) {
  const float2 sq = *reinterpret_cast<const float2*>(stats + 2 * idx);
  mean = sq.x * inv_p;
  const float var = fmaxf(sq.y * inv_p - mean * mean, 0.f);
  rstd = rsqrtf(var + 1e-5f);
}

// Up to three streamed input operands of a kernel.  base: first row of this block's image at the block's first channel;
// pitch: row pitch in bytes; es: element size.  Offsets inside a stage follow from the geometry.
struct StreamOps {
  const uint8_t* base[3];
  long pitch[3];
  int es[3];
  int n;
};

// Barrier set-up.  Runs BEFORE griddepcontrol.wait, i.e. while the previous kernel of the stream still drains.
#define BF_STREAM_SETUP()                                                                  \
  extern __shared__ __align__(128) uint8_t ring[];                                         \
  uint64_t* full_ = reinterpret_cast<uint64_t*>(ring + kRingBytes);                        \
  uint64_t* empty_ = full_ + kMaxStages;                                                   \
  if (threadIdx.x == 0) {                                                                  \
    for (int s_ = 0; s_ < kMaxStages; ++s_) { mbar_init(full_ + s_, 1); mbar_init(empty_ + s_, kNT / 32); } \
    fence_mbar_init();                                                                     \
  }                                                                                        \
  __syncthreads();                                                                         \
  int st_ = 0; uint32_t ph_ = 0;        /* ring position: continues across consecutive loops */ \
  pdl_prologue_done();

// The streaming loop.  `ops_` describes the inputs; BODY(row, s0, s1, s2) consumes one row: s0..s2 point at this thread's
// 8 channels of each operand in shared memory.  The producer warp falls through to whatever follows the loop.
#define BF_STREAM_LOOP(ops_, BODY) BF_STREAM_LOOP_(ops_, BODY, false)
// REV_: walk the slab from its last row block to its first (second pass of a fused kernel: the rows read last in the
// first pass are the most likely to still be in L2)
#define BF_STREAM_LOOP_(ops_, BODY, REV_)                                                  \
  {                                                                                        \
    const int cw_ = g.tx_n * 8;                           /* channels per block row */     \
    const int nblk_ = c.r1 > c.r0 ? (c.r1 - c.r0 + g.rb - 1) / g.rb : 0;                   \
    int off_[3];                                                                           \
    off_[0] = 0;                                                                           \
    off_[1] = g.rb * cw_ * (ops_).es[0];                                                   \
    off_[2] = off_[1] + g.rb * cw_ * ((ops_).n > 1 ? (ops_).es[1] : 0);                    \
    const int sum_es_ = (ops_).es[0] + (ops_).es[1] + (ops_).es[2];   /* unused operands have es = 0 */ \
    if (threadIdx.x >= kNT) {                                                              \
      const int lane_ = threadIdx.x & 31;                                                  \
      for (int bb_ = 0; bb_ < nblk_; ++bb_) {                                              \
        const int b_ = (REV_) ? nblk_ - 1 - bb_ : bb_;                                     \
        const int row_ = c.r0 + b_ * g.rb;                                                 \
        const int rows_ = min(g.rb, c.r1 - row_);                                          \
        if (lane_ == 0) {                                                                  \
          mbar_wait(empty_ + st_, ph_ ^ 1u);                                               \
          mbar_arrive_expect_tx(full_ + st_, (uint32_t)(rows_ * cw_ * sum_es_));           \
        }                                                                                  \
        __syncwarp();                                                                      \
        uint8_t* sp_ = ring + st_ * g.stage_bytes;                                         \
        _Pragma("unroll") for (int k_ = 0; k_ < 3; ++k_) {                                 \
          if (k_ >= (ops_).n) break;                                                       \
          const int rowb_ = cw_ * (ops_).es[k_];                                           \
          const uint8_t* src_ = (ops_).base[k_] + (long)row_ * (ops_).pitch[k_];           \
          if ((ops_).pitch[k_] == rowb_) {                                                 \
            if (lane_ == k_) bulk_g2s(sp_ + off_[k_], src_, (uint32_t)(rows_ * rowb_), full_ + st_); \
          } else {                                                                         \
            for (int r_ = lane_; r_ < rows_; r_ += 32)                                     \
              bulk_g2s(sp_ + off_[k_] + r_ * rowb_, src_ + (long)r_ * (ops_).pitch[k_], (uint32_t)rowb_, full_ + st_); \
          }                                                                                \
        }                                                                                  \
        if (++st_ == g.ns) { st_ = 0; ph_ ^= 1u; }                                         \
      }                                                                                    \
    } else {                                                                               \
      for (int bb_ = 0; bb_ < nblk_; ++bb_) {                                              \
        const int b_ = (REV_) ? nblk_ - 1 - bb_ : bb_;                                     \
        mbar_wait(full_ + st_, ph_);                                                       \
        if (c.active) {                                                                    \
          const uint8_t* sp_ = ring + st_ * g.stage_bytes;                                 \
          for (int k_ = 0; k_ < g.kk; ++k_) {                                              \
            const int rl_ = c.ty + k_ * g.ty_n;                                            \
            const int row_ = c.r0 + b_ * g.rb + rl_;                                       \
            if (row_ < c.r1) {                                                             \
              const int e_ = rl_ * cw_ + c.tx * 8;                                         \
              BODY(row_, (sp_ + off_[0] + e_ * (ops_).es[0]), (sp_ + off_[1] + e_ * (ops_).es[1]), \
                   (sp_ + off_[2] + e_ * (ops_).es[2]))                                    \
            }                                                                              \
          }                                                                                \
        }                                                                                  \
        __syncwarp();                                                                      \
        if ((threadIdx.x & 31) == 0) mbar_arrive(empty_ + st_);                            \
        if (++st_ == g.ns) { st_ = 0; ph_ ^= 1u; }                                         \
      }                                                                                    \
    }                                                                                      \
  }

// ---------------------------------------------------------------------------------------------
// statistics: stats[img][c] += (sum x, sum x^2)
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads, kBlocksPerSM)
inorm_stats_kernel(const T* __restrict__ x, long ldx, Geom g, float* __restrict__ stats) {
  BF_STREAM_SETUP()
  const Ctx c = make_ctx(g);
  StreamOps ops{};
  ops.n = 1;
  ops.base[0] = reinterpret_cast<const uint8_t*>(x + ((long)c.img * g.P) * ldx + (long)blockIdx.z * g.tx_n * 8);
  ops.pitch[0] = ldx * (long)sizeof(T); ops.es[0] = sizeof(T);
  float acc[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) acc[j] = 0.f;
#define BODY(row, s0, s1, s2)                                                \
  float v[8];                                                                \
  ld8<T>(s0, v);                                                             \
  _Pragma("unroll") for (int j = 0; j < 8; ++j) { acc[j] += v[j]; acc[8 + j] = fmaf(v[j], v[j], acc[8 + j]); }
  BF_STREAM_LOOP(ops, BODY)
#undef BODY
  reduce_over_ty<16>(acc, reinterpret_cast<float*>(ring), g, c);
  if (c.active && c.ty == 0) {
    float* dst = stats + ((long)c.img * g.C + (long)c.vcol * 8) * 2;
    float il[16];
#pragma unroll
    for (int j = 0; j < 8; ++j) { il[2 * j] = acc[j]; il[2 * j + 1] = acc[8 + j]; }
    red_add<16>(dst, il);
  }
}

// ---------------------------------------------------------------------------------------------
// apply: y = IN(x) [gelu] [film]; out = y | resid_in + row_scale*col_gamma*y ; optional statistics of `out`
// ---------------------------------------------------------------------------------------------
struct ApplyParams {
  Geom g;
  long ldx, ldo;
  const float* stats;
  const float* weight;
  const float* bias;
  int gelu;
  const float* film_gamma;   // [I / film_T][film_ld]
  const float* film_beta;
  int film_T;
  int film_ld;               // row pitch of film_gamma / film_beta (C, or 2C when both live in one (B, 2C) matrix)
  const float* resid_in;     // fp32 (I*P, C) ld = ldo, or null
  const float* row_scale;    // [I] or null
  const float* col_gamma;    // [C] (with resid_in)
  float* stats_out;          // [I][C][2] or null: += (sum, sum^2) of the values written
  int wide;                  // fp32 output rows are 32-byte aligned (256-bit stores)
};

template <typename TI, typename TO, bool RESID>
__global__ void __launch_bounds__(kThreads, kBlocksPerSM)
inorm_apply_kernel(const TI* __restrict__ x, TO* __restrict__ out, ApplyParams p) {
  BF_STREAM_SETUP()
  const Geom& g = p.g;
  const Ctx c = make_ctx(g);
  const int c0 = c.vcol * 8;
  const float inv_p = 1.f / (float)g.P;
  float a[8], b[8], fg[8], fb[8], cg[8];
  if (c.active) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float mean, rstd;
      mean_rstd(p.stats, (long)c.img * g.C + c0 + j, inv_p, mean, rstd);
      const float w = p.weight[c0 + j];
      a[j] = rstd * w;
      b[j] = p.bias[c0 + j] - mean * rstd * w;
      fg[j] = 1.f; fb[j] = 0.f; cg[j] = 1.f;
      if (p.film_gamma != nullptr) {
        const long fi = (long)(c.img / p.film_T) * p.film_ld + c0 + j;
        fg[j] = p.film_gamma[fi];
        fb[j] = p.film_beta[fi];
      }
      if (RESID) cg[j] = p.col_gamma[c0 + j] * (p.row_scale != nullptr ? p.row_scale[c.img] : 1.f);
    }
    if (!p.gelu && p.film_gamma != nullptr) {   // fold FiLM into the affine map
#pragma unroll
      for (int j = 0; j < 8; ++j) { a[j] *= fg[j]; b[j] = fmaf(fg[j], b[j], fb[j]); }
    }
    if (RESID) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { a[j] *= cg[j]; b[j] *= cg[j]; }
    }
  }
  const bool post_film = p.gelu && p.film_gamma != nullptr;
  const long cb = (long)blockIdx.z * g.tx_n * 8;        // first channel of this block's column chunk
  StreamOps ops{};
  ops.n = RESID ? 2 : 1;
  ops.base[0] = reinterpret_cast<const uint8_t*>(x + ((long)c.img * g.P) * p.ldx + cb);
  ops.pitch[0] = p.ldx * (long)sizeof(TI); ops.es[0] = sizeof(TI);
  if (RESID) {
    ops.base[1] = reinterpret_cast<const uint8_t*>(p.resid_in + ((long)c.img * g.P) * p.ldo + cb);
    ops.pitch[1] = p.ldo * 4; ops.es[1] = 4;
  }
  TO* ob = out + ((long)c.img * g.P) * p.ldo + c0;
  float acc[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) acc[j] = 0.f;
  const bool want_stats = p.stats_out != nullptr;
  const bool wide = p.wide != 0;
#define BODY(row, s0, s1, s2)                                                            \
  float v[8];                                                                            \
  ld8<TI>(s0, v);                                                                        \
  _Pragma("unroll") for (int j = 0; j < 8; ++j) v[j] = fmaf(v[j], a[j], b[j]);          \
  if (p.gelu) { _Pragma("unroll") for (int j = 0; j < 8; ++j) v[j] = gelu_fwd(v[j], p.gelu == 2); }  \
  if (post_film) { _Pragma("unroll") for (int j = 0; j < 8; ++j) v[j] = fmaf(fg[j], v[j], fb[j]); } \
  if (RESID) {                                                                           \
    float xr[8];                                                                         \
    ld8<float>(s1, xr);                                                                  \
    _Pragma("unroll") for (int j = 0; j < 8; ++j) v[j] += xr[j];                         \
  }                                                                                      \
  if (want_stats) { _Pragma("unroll") for (int j = 0; j < 8; ++j) { acc[j] += v[j]; acc[8 + j] = fmaf(v[j], v[j], acc[8 + j]); } } \
  store8<TO>(ob + (long)(row) * p.ldo, v, wide);
  BF_STREAM_LOOP(ops, BODY)
#undef BODY
  if (want_stats) {
    reduce_over_ty<16>(acc, reinterpret_cast<float*>(ring), g, c);
    if (c.active && c.ty == 0) {
      float* dst = p.stats_out + ((long)c.img * g.C + c0) * 2;
      float il[16];
#pragma unroll
      for (int j = 0; j < 8; ++j) { il[2 * j] = acc[j]; il[2 * j + 1] = acc[8 + j]; }
      red_add<16>(dst, il);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// backward, pass 1: red[img][c] += (sum g, sum g * xhat), g = gin [* gelu'(xhat*w+b)]
// ---------------------------------------------------------------------------------------------
struct BwdParams {
  Geom g;
  long ldg, ldx, ldo;
  const float* stats;
  const float* weight;
  const float* bias;
  int gelu;
  float* red;                // [I][C][2]
  // pass 2 only
  const float* row_scale;    // [I] or null
  const float* col_scale;    // [C] or null
  const float* film_gamma;   // [I / film_T][film_ld] or null
  int film_T;
  int film_ld;               // row pitch of film_gamma / dfilm_gamma / dfilm_beta
  const float* add32;        // fp32 tensor added to the result (residual-stream gradient), ld = ldo
  // pass 2, optional: parameter gradients from `red` (fp32 atomics by the first row split of every image)
  float* dweight; float* dbias; float* dcol_scale; float* dfilm_gamma; float* dfilm_beta;
  int wide;                  // fp32 output rows are 32-byte aligned (256-bit stores)
};

template <typename TG, typename TX, bool GELU>
__global__ void __launch_bounds__(kThreads, kBlocksPerSM)
inorm_bwd_reduce_kernel(const TG* __restrict__ gin, const TX* __restrict__ x, BwdParams p) {
  BF_STREAM_SETUP()
  const Geom& g = p.g;
  const Ctx c = make_ctx(g);
  const int c0 = c.vcol * 8;
  const float inv_p = 1.f / (float)g.P;
  float mean[8], rstd[8], wa[8], wb[8];
  if (c.active) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      mean_rstd(p.stats, (long)c.img * g.C + c0 + j, inv_p, mean[j], rstd[j]);
      if (GELU) {   // y = x*wa + wb
        const float w = p.weight[c0 + j];
        wa[j] = rstd[j] * w;
        wb[j] = p.bias[c0 + j] - mean[j] * rstd[j] * w;
      }
    }
  }
  const long cb = (long)blockIdx.z * g.tx_n * 8;
  StreamOps ops{};
  ops.n = 2;
  ops.base[0] = reinterpret_cast<const uint8_t*>(gin + ((long)c.img * g.P) * p.ldg + cb);
  ops.pitch[0] = p.ldg * (long)sizeof(TG); ops.es[0] = sizeof(TG);
  ops.base[1] = reinterpret_cast<const uint8_t*>(x + ((long)c.img * g.P) * p.ldx + cb);
  ops.pitch[1] = p.ldx * (long)sizeof(TX); ops.es[1] = sizeof(TX);
  float acc[16];                                  // [0..8): sum g, [8..16): sum g*x
#pragma unroll
  for (int j = 0; j < 16; ++j) acc[j] = 0.f;
#define BODY(row, s0, s1, s2)                                                           \
  float gv[8], xv[8];                                                                   \
  ld8<TG>(s0, gv);                                                                      \
  ld8<TX>(s1, xv);                                                                      \
  if (GELU && p.gelu == 1) {      /* tanh form: derivative of two elements per packed half-precision evaluation */ \
    _Pragma("unroll") for (int j = 0; j < 8; j += 2) {                                  \
      const float2 d2 = gelu_grad_h2(fmaf(xv[j], wa[j], wb[j]), fmaf(xv[j + 1], wa[j + 1], wb[j + 1])); \
      gv[j] *= d2.x; gv[j + 1] *= d2.y;                                                 \
    }                                                                                   \
  }                                                                                     \
  _Pragma("unroll") for (int j = 0; j < 8; ++j) {                                       \
    float gg = gv[j];                                                                   \
    if (GELU && p.gelu != 1) gg *= gelu_bwd(fmaf(xv[j], wa[j], wb[j]), p.gelu == 2);     \
    acc[j] += gg;                                                                       \
    acc[8 + j] = fmaf(gg, xv[j], acc[8 + j]);                                           \
  }
  BF_STREAM_LOOP(ops, BODY)
#undef BODY
  reduce_over_ty<16>(acc, reinterpret_cast<float*>(ring), g, c);
  if (c.active && c.ty == 0) {
    float* dst = p.red + ((long)c.img * g.C + c0) * 2;
    float il[16];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      // sum g*xhat = rstd * (sum g*x - mean * sum g)
      il[2 * j] = acc[j];
      il[2 * j + 1] = rstd[j] * (acc[8 + j] - mean[j] * acc[j]);
    }
    red_add<16>(dst, il);
  }
}

// backward, pass 2: dx = rstd*w*cs*(g - R1/P - xhat*R2/P) [+ add32]   ( = A*g + B*x + C0 per channel )
template <typename TG, typename TX, typename TO, bool GELU, bool ADD>
__global__ void __launch_bounds__(kThreads, kBlocksPerSM)
inorm_bwd_apply_kernel(const TG* __restrict__ gin, const TX* __restrict__ x, TO* __restrict__ out, BwdParams p) {
  BF_STREAM_SETUP()
  const Geom& g = p.g;
  const Ctx c = make_ctx(g);
  const int c0 = c.vcol * 8;
  const float inv_p = 1.f / (float)g.P;
  float ka[8], kb[8], kc[8], wa[8], wb[8];
  if (c.active) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const long idx = (long)c.img * g.C + c0 + j;
      float mean, rstd;
      mean_rstd(p.stats, idx, inv_p, mean, rstd);
      const float w = p.weight[c0 + j];
      float cs = 1.f;
      if (p.row_scale != nullptr) cs *= p.row_scale[c.img];
      if (p.col_scale != nullptr) cs *= p.col_scale[c0 + j];
      if (p.film_gamma != nullptr) cs *= p.film_gamma[(long)(c.img / p.film_T) * p.film_ld + c0 + j];
      const float k = rstd * w * cs;
      const float m1 = p.red[2 * idx] * inv_p, m2 = p.red[2 * idx + 1] * inv_p;
      ka[j] = k;
      kb[j] = -k * m2 * rstd;
      kc[j] = -k * m1 + k * m2 * rstd * mean;
      if (GELU) { wa[j] = rstd * w; wb[j] = p.bias[c0 + j] - mean * rstd * w; }
      if (blockIdx.x == 0 && c.ty == 0 && p.dweight != nullptr) {
        // dweight += cs*R2, dbias += cs*R1, dcol_scale += rs*(w*R2 + b*R1), dfilm_gamma[b] += w*R2 + b*R1, dfilm_beta[b] += R1
        const float R1 = p.red[2 * idx], R2 = p.red[2 * idx + 1];
        const float b = p.bias[c0 + j];
        atomicAdd(p.dweight + c0 + j, cs * R2);
        atomicAdd(p.dbias + c0 + j, cs * R1);
        if (p.dcol_scale != nullptr)
          atomicAdd(p.dcol_scale + c0 + j, (p.row_scale != nullptr ? p.row_scale[c.img] : 1.f) * fmaf(w, R2, b * R1));
        if (p.dfilm_gamma != nullptr) {
          const long fi = (long)(c.img / p.film_T) * p.film_ld + c0 + j;
          atomicAdd(p.dfilm_gamma + fi, fmaf(w, R2, b * R1));
          atomicAdd(p.dfilm_beta + fi, R1);
        }
      }
    }
  }
  const long cb = (long)blockIdx.z * g.tx_n * 8;
  StreamOps ops{};
  ops.n = ADD ? 3 : 2;
  ops.base[0] = reinterpret_cast<const uint8_t*>(gin + ((long)c.img * g.P) * p.ldg + cb);
  ops.pitch[0] = p.ldg * (long)sizeof(TG); ops.es[0] = sizeof(TG);
  ops.base[1] = reinterpret_cast<const uint8_t*>(x + ((long)c.img * g.P) * p.ldx + cb);
  ops.pitch[1] = p.ldx * (long)sizeof(TX); ops.es[1] = sizeof(TX);
  if (ADD) {
    ops.base[2] = reinterpret_cast<const uint8_t*>(p.add32 + ((long)c.img * g.P) * p.ldo + cb);
    ops.pitch[2] = p.ldo * 4; ops.es[2] = 4;
  }
  TO* ob = out + ((long)c.img * g.P) * p.ldo + c0;
  const bool wide = p.wide != 0;
#define BODY(row, s0, s1, s2)                                                           \
  float gv[8], xv[8], o[8];                                                             \
  ld8<TG>(s0, gv);                                                                      \
  ld8<TX>(s1, xv);                                                                      \
  if (GELU && p.gelu == 1) {                                                            \
    _Pragma("unroll") for (int j = 0; j < 8; j += 2) {                                  \
      const float2 d2 = gelu_grad_h2(fmaf(xv[j], wa[j], wb[j]), fmaf(xv[j + 1], wa[j + 1], wb[j + 1])); \
      gv[j] *= d2.x; gv[j + 1] *= d2.y;                                                 \
    }                                                                                   \
  }                                                                                     \
  _Pragma("unroll") for (int j = 0; j < 8; ++j) {                                       \
    float gg = gv[j];                                                                   \
    if (GELU && p.gelu != 1) gg *= gelu_bwd(fmaf(xv[j], wa[j], wb[j]), p.gelu == 2);     \
    o[j] = fmaf(ka[j], gg, fmaf(kb[j], xv[j], kc[j]));                                  \
  }                                                                                     \
  if (ADD) {                                                                            \
    float av[8];                                                                        \
    ld8<float>(s2, av);                                                                 \
    _Pragma("unroll") for (int j = 0; j < 8; ++j) o[j] += av[j];                        \
  }                                                                                     \
  store8<TO>(ob + (long)(row) * p.ldo, o, wide);
  BF_STREAM_LOOP(ops, BODY)
#undef BODY
}


// ---------------------------------------------------------------------------------------------
// Fused two-phase kernels.  The slabs of ONE image form a thread-block cluster (<= 8 blocks, co-scheduled by the
// hardware): every block reduces its slab, leaves its per-channel partial sums in its own shared memory, the cluster
// synchronises, every block adds up all partials through distributed shared memory and streams its slab a second time
// (back to front: out of L2, not HBM) to apply the result.  One launch, no global atomics, no zeroed buffers; the
// reduced values are also written to global memory (by the first block of the cluster) for later consumers.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ld_cluster_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
// acc[16] of the ty == 0 threads -> this block's partial table -> sums over the cluster in tot[16] of every active thread.
// Only the ty == 0 threads read the peers' tables (distributed shared memory); the block's other threads pick the totals
// up from local shared memory (every thread reading every peer directly cost 25 us per launch at config 2).
__device__ __forceinline__ void cluster_total16(float (&acc)[16], float (&tot)[16], float* part, const Geom& g, const Ctx& c) {
  fence_proxy_async();        // the ring was used as reduction scratch (generic proxy); bulk copies refill it next
  float* mine = part + c.tx * 16;
  float* total = part + g.tx_n * 16 + c.tx * 16;
  if (c.active && c.ty == 0) {
#pragma unroll
    for (int j = 0; j < 16; j += 4)
      *reinterpret_cast<float4*>(mine + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
  }
  cluster_sync_all();
#pragma unroll
  for (int j = 0; j < 16; ++j) tot[j] = 0.f;
  if (c.active && c.ty == 0) {
    const uint32_t local = smem_u32(mine);
    for (int r = 0; r < g.splits; ++r) {
      const uint32_t remote = mapa_shared(local, (uint32_t)r);
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        const float4 v = ld_cluster_f4(remote + j * 4);
        tot[j] += v.x; tot[j + 1] += v.y; tot[j + 2] += v.z; tot[j + 3] += v.w;
      }
    }
#pragma unroll
    for (int j = 0; j < 16; j += 4)
      *reinterpret_cast<float4*>(total + j) = make_float4(tot[j], tot[j + 1], tot[j + 2], tot[j + 3]);
  }
  __syncthreads();
  if (c.active && c.ty != 0) {
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
      const float4 v = *reinterpret_cast<const float4*>(total + j);
      tot[j] = v.x; tot[j + 1] = v.y; tot[j + 2] = v.z; tot[j + 3] = v.w;
    }
  }
}

// forward: statistics of x, then out = IN(x) * weight + bias
template <typename TI, typename TO>
__global__ void __launch_bounds__(kThreads, kBlocksPerSM)
inorm_fwd_fused_kernel(const TI* __restrict__ x, TO* __restrict__ out, ApplyParams p, float* __restrict__ stats_w) {
  BF_STREAM_SETUP()
  float* part = reinterpret_cast<float*>(ring + kFusedRingBytes);
  const Geom& g = p.g;
  const Ctx c = make_ctx(g);
  const int c0 = c.vcol * 8;
  StreamOps ops{};
  ops.n = 1;
  ops.base[0] = reinterpret_cast<const uint8_t*>(x + ((long)c.img * g.P) * p.ldx);
  ops.pitch[0] = p.ldx * (long)sizeof(TI); ops.es[0] = sizeof(TI);
  float acc[16], tot[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) acc[j] = 0.f;
#define BODY(row, s0, s1, s2)                                                \
  float v[8];                                                                \
  ld8<TI>(s0, v);                                                            \
  _Pragma("unroll") for (int j = 0; j < 8; ++j) { acc[j] += v[j]; acc[8 + j] = fmaf(v[j], v[j], acc[8 + j]); }
  BF_STREAM_LOOP(ops, BODY)
#undef BODY
  reduce_over_ty<16>(acc, reinterpret_cast<float*>(ring), g, c);
  cluster_total16(acc, tot, part, g, c);
  const float inv_p = 1.f / (float)g.P;
  float a[8], b[8];
  if (c.active) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float mean = tot[j] * inv_p;
      const float rstd = rsqrtf(fmaxf(tot[8 + j] * inv_p - mean * mean, 0.f) + 1e-5f);
      const float w = p.weight[c0 + j];
      a[j] = rstd * w;
      b[j] = p.bias[c0 + j] - mean * rstd * w;
    }
    if (blockIdx.x == 0 && c.ty == 0) {            // raw sums for the backward pass: [img][c][2]
      float4* dst = reinterpret_cast<float4*>(stats_w + ((long)c.img * g.C + c0) * 2);
#pragma unroll
      for (int j = 0; j < 8; j += 2) dst[j / 2] = make_float4(tot[j], tot[8 + j], tot[j + 1], tot[9 + j]);
    }
  }
  TO* ob = out + ((long)c.img * g.P) * p.ldo + c0;
  const bool wide = p.wide != 0;
#define BODY(row, s0, s1, s2)                                                \
  float v[8];                                                                \
  ld8<TI>(s0, v);                                                            \
  _Pragma("unroll") for (int j = 0; j < 8; ++j) v[j] = fmaf(v[j], a[j], b[j]); \
  store8<TO>(ob + (long)(row) * p.ldo, v, wide);
  BF_STREAM_LOOP_(ops, BODY, true)
#undef BODY
  cluster_sync_all();                              // nobody leaves while a peer may still read its partials
}

// backward: (sum g, sum g*xhat) per (image, channel), then dx = rstd*w*cs*(g - R1/P - xhat*R2/P) [+ add32]
template <typename TG, typename TX, typename TO, bool ADD>
__global__ void __launch_bounds__(kThreads, kBlocksPerSM)
inorm_bwd_fused_kernel(const TG* __restrict__ gin, const TX* __restrict__ x, TO* __restrict__ out, BwdParams p) {
  BF_STREAM_SETUP()
  float* part = reinterpret_cast<float*>(ring + kFusedRingBytes);
  const Geom& g = p.g;
  const Ctx c = make_ctx(g);
  const int c0 = c.vcol * 8;
  StreamOps ops{};
  ops.n = 2;
  ops.base[0] = reinterpret_cast<const uint8_t*>(gin + ((long)c.img * g.P) * p.ldg);
  ops.pitch[0] = p.ldg * (long)sizeof(TG); ops.es[0] = sizeof(TG);
  ops.base[1] = reinterpret_cast<const uint8_t*>(x + ((long)c.img * g.P) * p.ldx);
  ops.pitch[1] = p.ldx * (long)sizeof(TX); ops.es[1] = sizeof(TX);
  float acc[16], tot[16];                          // [0..8): sum g, [8..16): sum g*x
#pragma unroll
  for (int j = 0; j < 16; ++j) acc[j] = 0.f;
#define BODY(row, s0, s1, s2)                                                \
  float gv[8], xv[8];                                                        \
  ld8<TG>(s0, gv);                                                           \
  ld8<TX>(s1, xv);                                                           \
  _Pragma("unroll") for (int j = 0; j < 8; ++j) { acc[j] += gv[j]; acc[8 + j] = fmaf(gv[j], xv[j], acc[8 + j]); }
  BF_STREAM_LOOP(ops, BODY)
#undef BODY
  reduce_over_ty<16>(acc, reinterpret_cast<float*>(ring), g, c);
  cluster_total16(acc, tot, part, g, c);
  const float inv_p = 1.f / (float)g.P;
  float ka[8], kb[8], kc[8];
  if (c.active) {
    float r1v[8], r2v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const long idx = (long)c.img * g.C + c0 + j;
      float mean, rstd;
      mean_rstd(p.stats, idx, inv_p, mean, rstd);
      const float R1 = tot[j], R2 = rstd * (tot[8 + j] - mean * tot[j]);
      r1v[j] = R1; r2v[j] = R2;
      const float w = p.weight[c0 + j];
      float cs = 1.f;
      if (p.row_scale != nullptr) cs *= p.row_scale[c.img];
      if (p.col_scale != nullptr) cs *= p.col_scale[c0 + j];
      const float k = rstd * w * cs;
      const float m1 = R1 * inv_p, m2 = R2 * inv_p;
      ka[j] = k;
      kb[j] = -k * m2 * rstd;
      kc[j] = -k * m1 + k * m2 * rstd * mean;
      if (blockIdx.x == 0 && c.ty == 0 && p.dweight != nullptr) {
        const float b = p.bias[c0 + j];
        atomicAdd(p.dweight + c0 + j, cs * R2);
        atomicAdd(p.dbias + c0 + j, cs * R1);
        if (p.dcol_scale != nullptr)
          atomicAdd(p.dcol_scale + c0 + j, (p.row_scale != nullptr ? p.row_scale[c.img] : 1.f) * fmaf(w, R2, b * R1));
      }
    }
    if (blockIdx.x == 0 && c.ty == 0) {
      float4* dst = reinterpret_cast<float4*>(p.red + ((long)c.img * g.C + c0) * 2);
#pragma unroll
      for (int j = 0; j < 8; j += 2) dst[j / 2] = make_float4(r1v[j], r2v[j], r1v[j + 1], r2v[j + 1]);
    }
  }
  if (ADD) {
    ops.n = 3;
    ops.base[2] = reinterpret_cast<const uint8_t*>(p.add32 + ((long)c.img * g.P) * p.ldo);
    ops.pitch[2] = p.ldo * 4; ops.es[2] = 4;
  }
  TO* ob = out + ((long)c.img * g.P) * p.ldo + c0;
  const bool wide = p.wide != 0;
#define BODY(row, s0, s1, s2)                                                \
  float gv[8], xv[8], o[8];                                                  \
  ld8<TG>(s0, gv);                                                           \
  ld8<TX>(s1, xv);                                                           \
  _Pragma("unroll") for (int j = 0; j < 8; ++j) o[j] = fmaf(ka[j], gv[j], fmaf(kb[j], xv[j], kc[j])); \
  if (ADD) {                                                                 \
    float av[8];                                                             \
    ld8<float>(s2, av);                                                      \
    _Pragma("unroll") for (int j = 0; j < 8; ++j) o[j] += av[j];             \
  }                                                                          \
  store8<TO>(ob + (long)(row) * p.ldo, o, wide);
  BF_STREAM_LOOP_(ops, BODY, true)
#undef BODY
  cluster_sync_all();
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
  pdl_prologue_done();
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
  Geom g;
  long lddx, ldz;
  const float* dx;
  const void* z16;           // (I*P, C) 16-bit, ld = ldz, may be null (then S1 untouched)
  const float* row_scale;    // [I] or null
  const float* coef;         // [C]
  void* dz16;                // out, ld = ldz, may be null
  float* S0; float* S1;      // [I][C] per-image sums
  int wide;                  // fp32 validation backend: dz rows are 32-byte aligned
};
template <typename T16, bool HASZ>
__global__ void __launch_bounds__(kThreads, kBlocksPerSM)
resid_bwd_kernel(ResidBwdParams p) {
  BF_STREAM_SETUP()
  const Geom& g = p.g;
  const Ctx c = make_ctx(g);
  const int c0 = c.vcol * 8;
  const float rs = (p.row_scale != nullptr && c.active) ? p.row_scale[c.img] : 1.f;
  float cf[8];
  if (c.active) {
#pragma unroll
    for (int j = 0; j < 8; ++j) cf[j] = p.coef[c0 + j] * rs;
  }
  const long cb = (long)blockIdx.z * g.tx_n * 8;
  StreamOps ops{};
  ops.n = HASZ ? 2 : 1;
  ops.base[0] = reinterpret_cast<const uint8_t*>(p.dx + ((long)c.img * g.P) * p.lddx + cb);
  ops.pitch[0] = p.lddx * 4; ops.es[0] = 4;
  if (HASZ) {
    ops.base[1] = reinterpret_cast<const uint8_t*>(reinterpret_cast<const T16*>(p.z16) + ((long)c.img * g.P) * p.ldz + cb);
    ops.pitch[1] = p.ldz * (long)sizeof(T16); ops.es[1] = sizeof(T16);
  }
  T16* ob = p.dz16 ? reinterpret_cast<T16*>(p.dz16) + ((long)c.img * g.P) * p.ldz + c0 : nullptr;
  float acc[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) acc[j] = 0.f;
#define BODY(row, s0, s1, s2)                                                           \
  float dv[8], o[8];                                                                    \
  ld8<float>(s0, dv);                                                                   \
  if (HASZ) {                                                                           \
    float zv[8];                                                                        \
    ld8<T16>(s1, zv);                                                                   \
    _Pragma("unroll") for (int j = 0; j < 8; ++j) acc[8 + j] = fmaf(dv[j], zv[j], acc[8 + j]); \
  }                                                                                     \
  _Pragma("unroll") for (int j = 0; j < 8; ++j) { acc[j] += dv[j]; o[j] = cf[j] * dv[j]; } \
  if (ob != nullptr) store8<T16>(ob + (long)(row) * p.ldz, o, p.wide != 0);
  BF_STREAM_LOOP(ops, BODY)
#undef BODY
  reduce_over_ty<16>(acc, reinterpret_cast<float*>(ring), g, c);
  if (c.active && c.ty == 0) {
    float s0[8], s1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s0[j] = rs * acc[j]; s1[j] = rs * acc[8 + j]; }
    red_add<8>(p.S0 + (long)c.img * g.C + c0, s0);
    if (HASZ) red_add<8>(p.S1 + (long)c.img * g.C + c0, s1);
  }
}

// column sums of a 16-bit matrix: out[c] += sum_rows x[r, c]   (bias gradients)
template <typename T16>
__global__ void __launch_bounds__(kThreads, kBlocksPerSM)
colsum16_kernel(const T16* __restrict__ x, long ldx, Geom g, float* __restrict__ out) {
  BF_STREAM_SETUP()
  const Ctx c = make_ctx(g);
  StreamOps ops{};
  ops.n = 1;
  ops.base[0] = reinterpret_cast<const uint8_t*>(x + ((long)c.img * g.P) * ldx + (long)blockIdx.z * g.tx_n * 8);
  ops.pitch[0] = ldx * (long)sizeof(T16); ops.es[0] = sizeof(T16);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#define BODY(row, s0, s1, s2)                                        \
  float v[8];                                                        \
  ld8<T16>(s0, v);                                                   \
  _Pragma("unroll") for (int j = 0; j < 8; ++j) acc[j] += v[j];
  BF_STREAM_LOOP(ops, BODY)
#undef BODY
  reduce_over_ty<8>(acc, reinterpret_cast<float*>(ring), g, c);
  if (c.active && c.ty == 0) {
    red_add<8>(out + (long)c.vcol * 8, acc);
  }
}

static int check_common(const char* fn, int I, int P, int C, long ld, const void* x) {
  BF_REQUIRE(I > 0 && P > 0 && C > 0, "%s: empty problem I=%d P=%d C=%d", fn, I, P, C);
  BF_REQUIRE(C % 8 == 0, "%s: C=%d must be a multiple of 8", fn, C);
  BF_REQUIRE(ld >= C && ld % 8 == 0, "%s: ld=%ld must be a multiple of 8 and >= C", fn, ld);
  BF_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "%s: pointer must be 16-byte aligned", fn);
  BF_REQUIRE(I <= 65535, "%s: more than 65535 images", fn);
  return BF_OK;
}

static inline int es_of(int dtype) { return dtype == BF_F32 ? 4 : 2; }
// fp32 rows that start on 32-byte boundaries take 256-bit stores
static inline int wide_ok(const void* out, long ld) { return ((reinterpret_cast<uintptr_t>(out) & 31) == 0 && ld % 8 == 0) ? 1 : 0; }

template <typename K>
static int set_smem(K kern) {
  return check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes),
                    "cudaFuncSetAttribute(norm)");
}
// launch with the bulk-copy ring; the attribute is set once per kernel instantiation
#define BF_NORM_LAUNCH(kern, grid, stream, ...)                                   \
  do {                                                                            \
    static bool done_ = false;                                                    \
    if (!done_) { if (int e_ = set_smem(kern)) return e_; done_ = true; }         \
    launch_k(kern, dim3(grid), dim3(kThreads), (size_t)(kSmemBytes), stream, __VA_ARGS__);                    \
  } while (0)

// same, as one thread-block cluster per image (fused two-phase kernels)
template <typename K>
static void debug_cluster_occupancy(K kern, dim3 grid, unsigned cluster_x) {
  if (getenv("BF_DEBUG_OCC") == nullptr) return;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = kSmemBytes;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cluster_x; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int n = -1;
  cudaError_t e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
  int per_sm = -1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, kSmemBytes);
  cudaFuncAttributes fa{};
  cudaFuncGetAttributes(&fa, kern);
  fprintf(stderr, "[bf] cluster kernel: grid %u x %u cluster %u -> max active clusters %d (%s), blocks/SM %d, regs %d\n", grid.x, grid.y,
          cluster_x, n, cudaGetErrorString(e), per_sm, fa.numRegs);
}
#define BF_NORM_LAUNCH_CLUSTER(kern, grid, cluster_x, stream, ...)                \
  do {                                                                            \
    static bool done_ = false;                                                    \
    if (!done_) { if (int e_ = set_smem(kern)) return e_; done_ = true; debug_cluster_occupancy(kern, dim3(grid), cluster_x); } \
    launch_k_cluster(kern, dim3(grid), dim3(kThreads), (size_t)(kSmemBytes), stream, (unsigned)(cluster_x), __VA_ARGS__); \
  } while (0)

// How many clusters of `cs` fused-kernel blocks the device can hold at once (cudaOccupancyMaxActiveClusters; all fused
// instantiations share block size, shared memory and the 2-blocks-per-SM register bound).  Cached per cluster size.
static int max_active_clusters(int cs) {
  static int cache[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (cs < 1 || cs > 8) return 0;
  if (cache[cs] == 0) {
    auto kern = inorm_fwd_fused_kernel<__nv_bfloat16, __nv_bfloat16>;
    if (set_smem(kern) != BF_OK) return 0;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(cs, 64, 1); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = kSmemBytes;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { cudaGetLastError(); n = -1; }
    cache[cs] = n > 0 ? n : -1;
  }
  return cache[cs] > 0 ? cache[cs] : 0;
}

// The fused kernels need the whole image in one cluster (<= 8 slabs), one column chunk, ALL clusters resident at once
// (a second wave of a few clusters doubles the time: measured with 7-block clusters, 37 of 40 fit) and enough blocks to
// fill the machine (otherwise two launches of a full wave are faster).  BF_NORM_FUSED=0 forces the two-launch form.
static bool fused_geom(int I, int P, int C, int sum_es, Geom* g) {
  static int on = -1;
  if (on < 0) { const char* e = getenv("BF_NORM_FUSED"); on = (e != nullptr && e[0] == '0') ? 0 : 1; }
  if (!on || C > kFusedMaxC || C / 8 > kNT) return false;
  int cs = 8;
  while (cs >= 2 && max_active_clusters(cs) < I) --cs;
  if (cs < 2) return false;
  *g = make_geom(I, P, C, sum_es, kFusedRingBytes, cs);
  if (g->chunks != 1 || g->ns < 2) return false;
  return (long)I * g->splits >= num_sms();
}

}  // namespace bf

using namespace bf;

extern "C" int bf_inorm_stats(const void* x, int x_dtype, int I, int P, int C, int64_t ldx, float* stats,
                              void* stream) {
  BF_REQUIRE(x && stats, "bf_inorm_stats: null pointer");
  BF_REQUIRE(x_dtype == BF_BF16 || x_dtype == BF_F16 || x_dtype == BF_F32, "bf_inorm_stats: dtype %d", x_dtype);
  if (int st = check_common("bf_inorm_stats", I, P, C, ldx, x)) return st;
  const Geom g = make_geom(I, P, C, es_of(x_dtype));
  dim3 grid(g.splits, I, g.chunks);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (x_dtype == BF_F32) BF_NORM_LAUNCH(inorm_stats_kernel<float>, grid, s, (const float*)x, (long)ldx, g, stats);
  else if (x_dtype == BF_BF16) BF_NORM_LAUNCH(inorm_stats_kernel<__nv_bfloat16>, grid, s, (const __nv_bfloat16*)x, (long)ldx, g, stats);
  else BF_NORM_LAUNCH(inorm_stats_kernel<__half>, grid, s, (const __half*)x, (long)ldx, g, stats);
  count_launch();
  BF_LAUNCH_CHECK("inorm_stats_kernel");
  return BF_OK;
}

extern "C" int bf_inorm_apply(const bf_inorm_apply_args* a, void* stream) {
  BF_REQUIRE(a && a->x && a->out && a->stats && a->weight && a->bias, "bf_inorm_apply: null pointer");
  if (int st = check_common("bf_inorm_apply", a->I, a->P, a->C, a->ldx, a->x)) return st;
  BF_REQUIRE(a->ldo >= a->C && a->ldo % 8 == 0, "bf_inorm_apply: ldo");
  BF_REQUIRE((reinterpret_cast<uintptr_t>(a->out) & 15) == 0, "bf_inorm_apply: out must be 16-byte aligned");
  BF_REQUIRE((a->film_gamma == nullptr) == (a->film_beta == nullptr), "bf_inorm_apply: film pair");
  BF_REQUIRE(a->film_gamma == nullptr || (a->film_T > 0 && a->I % a->film_T == 0), "bf_inorm_apply: film_T");
  BF_REQUIRE(a->resid_in == nullptr || (a->out_dtype == BF_F32 && a->col_gamma), "bf_inorm_apply: residual needs f32 out");
  BF_REQUIRE(!(a->resid_in != nullptr && a->gelu), "bf_inorm_apply: gelu and the residual epilogue are exclusive");
  if (a->compute_stats) {
    // `stats` is an output here: one fused launch when the shape allows it, otherwise statistics pass + this call again
    const bool plain = !a->gelu && a->film_gamma == nullptr && a->resid_in == nullptr && a->stats_out == nullptr;
    Geom fg;
    const int xi = a->x_dtype, xo = a->out_dtype;
    const bool types = (xi == BF_BF16 && xo == BF_BF16) || (xi == BF_F32 && xo == BF_BF16) || (xi == BF_F32 && xo == BF_F32) ||
                       (xi == BF_F16 && xo == BF_F16);
    if (plain && types && fused_geom(a->I, a->P, a->C, es_of(xi), &fg)) {
      ApplyParams p{};
      p.g = fg;
      p.wide = wide_ok(a->out, a->ldo);
      p.ldx = a->ldx; p.ldo = a->ldo;
      p.weight = a->weight; p.bias = a->bias;
      dim3 grid(fg.splits, a->I, 1);
      cudaStream_t s = static_cast<cudaStream_t>(stream);
      float* sw = const_cast<float*>(a->stats);
      if (xi == BF_BF16) BF_NORM_LAUNCH_CLUSTER((inorm_fwd_fused_kernel<__nv_bfloat16, __nv_bfloat16>), grid, fg.splits, s, (const __nv_bfloat16*)a->x, (__nv_bfloat16*)a->out, p, sw);
      else if (xi == BF_F16) BF_NORM_LAUNCH_CLUSTER((inorm_fwd_fused_kernel<__half, __half>), grid, fg.splits, s, (const __half*)a->x, (__half*)a->out, p, sw);
      else if (xo == BF_BF16) BF_NORM_LAUNCH_CLUSTER((inorm_fwd_fused_kernel<float, __nv_bfloat16>), grid, fg.splits, s, (const float*)a->x, (__nv_bfloat16*)a->out, p, sw);
      else BF_NORM_LAUNCH_CLUSTER((inorm_fwd_fused_kernel<float, float>), grid, fg.splits, s, (const float*)a->x, (float*)a->out, p, sw);
      count_launch();
      BF_LAUNCH_CHECK("inorm_fwd_fused_kernel");
      return BF_OK;
    }
    if (int st = check_cuda(cudaMemsetAsync(const_cast<float*>(a->stats), 0, (size_t)a->I * a->C * 2 * sizeof(float),
                                            static_cast<cudaStream_t>(stream)), "cudaMemsetAsync(stats)")) return st;
    if (int st = bf_inorm_stats(a->x, a->x_dtype, a->I, a->P, a->C, a->ldx, const_cast<float*>(a->stats), stream)) return st;
    bf_inorm_apply_args b = *a;
    b.compute_stats = 0;
    return bf_inorm_apply(&b, stream);
  }
  ApplyParams p{};
  p.g = make_geom(a->I, a->P, a->C, es_of(a->x_dtype) + (a->resid_in != nullptr ? 4 : 0));
  p.wide = wide_ok(a->out, a->ldo);
  p.ldx = a->ldx; p.ldo = a->ldo;
  p.stats = a->stats; p.weight = a->weight; p.bias = a->bias; p.gelu = a->gelu ? (gelu_exact() ? 2 : 1) : 0;
  p.film_gamma = a->film_gamma; p.film_beta = a->film_beta; p.film_T = a->film_T > 0 ? a->film_T : 1;
  p.film_ld = a->film_ld > 0 ? a->film_ld : a->C;
  BF_REQUIRE(p.film_ld >= a->C, "bf_inorm_apply: film_ld < C");
  p.resid_in = a->resid_in; p.row_scale = a->row_scale; p.col_gamma = a->col_gamma;
  p.stats_out = a->stats_out;
  dim3 grid(p.g.splits, a->I, p.g.chunks);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool resid = a->resid_in != nullptr;
#define BF_APPLY(TI, TO)                                                                                       \
  do {                                                                                                         \
    if (resid) BF_NORM_LAUNCH((inorm_apply_kernel<TI, TO, true>), grid, s, (const TI*)a->x, (TO*)a->out, p);   \
    else BF_NORM_LAUNCH((inorm_apply_kernel<TI, TO, false>), grid, s, (const TI*)a->x, (TO*)a->out, p);        \
  } while (0)
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
  if (int st = check_common("bf_inorm_bwd", a->I, a->P, a->C, a->ldx, a->x)) return st;
  BF_REQUIRE(a->ldg >= a->C && a->ldg % 8 == 0, "bf_inorm_bwd: ldg");
  BF_REQUIRE((reinterpret_cast<uintptr_t>(a->gin) & 15) == 0, "bf_inorm_bwd: gin must be 16-byte aligned");
  BF_REQUIRE(a->phase == 1 || a->phase == 2 || a->phase == 3, "bf_inorm_bwd: phase %d", a->phase);
  if (a->phase == 3) {
    // both phases; `red` is an output (no zeroing by the caller).  One fused launch when the shape allows it.
    BF_REQUIRE(a->out, "bf_inorm_bwd: out required in phase 3");
    BF_REQUIRE(a->ldo >= a->C && a->ldo % 8 == 0, "bf_inorm_bwd: ldo");
    BF_REQUIRE((reinterpret_cast<uintptr_t>(a->out) & 15) == 0, "bf_inorm_bwd: out must be 16-byte aligned");
    const int gd = a->g_dtype, xd = a->x_dtype, od = a->out_dtype;
    const bool add = a->add32 != nullptr;
    const bool plain = !a->gelu && a->film_gamma == nullptr && a->dfilm_gamma == nullptr;
    const int combo = (gd == BF_BF16 && xd == BF_BF16 && od == BF_BF16 && !add) ? 1
                    : (gd == BF_BF16 && xd == BF_F32 && od == BF_F32) ? 2
                    : (gd == BF_F32 && xd == BF_BF16 && od == BF_BF16 && !add) ? 3
                    : (gd == BF_F32 && xd == BF_F32 && od == BF_F32) ? 4 : 0;
    Geom fg;
    if (plain && combo != 0 && fused_geom(a->I, a->P, a->C, es_of(gd) + es_of(xd) + (add ? 4 : 0), &fg)) {
      BwdParams p{};
      p.g = fg;
      p.wide = wide_ok(a->out, a->ldo);
      p.ldg = a->ldg; p.ldx = a->ldx; p.ldo = a->ldo;
      p.stats = a->stats; p.weight = a->weight; p.bias = a->bias; p.red = a->red;
      p.row_scale = a->row_scale; p.col_scale = a->col_scale; p.add32 = a->add32;
      p.film_T = 1; p.film_ld = a->C;
      p.dweight = a->dweight; p.dbias = a->dbias; p.dcol_scale = a->dcol_scale;
      BF_REQUIRE((a->dweight == nullptr) == (a->dbias == nullptr), "bf_inorm_bwd: dweight / dbias pair");
      BF_REQUIRE(a->dweight != nullptr || a->dcol_scale == nullptr, "bf_inorm_bwd: dcol_scale needs dweight / dbias as well");
      dim3 grid(fg.splits, a->I, 1);
      cudaStream_t s = static_cast<cudaStream_t>(stream);
      typedef __nv_bfloat16 bf16;
#define BF_FUSED(TG, TX, TO, ADD_) BF_NORM_LAUNCH_CLUSTER((inorm_bwd_fused_kernel<TG, TX, TO, ADD_>), grid, fg.splits, s, (const TG*)a->gin, (const TX*)a->x, (TO*)a->out, p)
      if (combo == 1) BF_FUSED(bf16, bf16, bf16, false);
      else if (combo == 2) { if (add) BF_FUSED(bf16, float, float, true); else BF_FUSED(bf16, float, float, false); }
      else if (combo == 3) BF_FUSED(float, bf16, bf16, false);
      else { if (add) BF_FUSED(float, float, float, true); else BF_FUSED(float, float, float, false); }
#undef BF_FUSED
      count_launch();
      BF_LAUNCH_CHECK("inorm_bwd_fused_kernel");
      return BF_OK;
    }
    if (int st = check_cuda(cudaMemsetAsync(a->red, 0, (size_t)a->I * a->C * 2 * sizeof(float), static_cast<cudaStream_t>(stream)),
                            "cudaMemsetAsync(red)")) return st;
    bf_inorm_bwd_args b = *a;
    b.phase = 1;
    if (int st = bf_inorm_bwd(&b, stream)) return st;
    b.phase = 2;
    return bf_inorm_bwd(&b, stream);
  }
  BwdParams p{};
  p.g = make_geom(a->I, a->P, a->C, es_of(a->g_dtype) + es_of(a->x_dtype) + ((a->phase == 2 && a->add32 != nullptr) ? 4 : 0));
  p.wide = (a->phase == 2 && a->out != nullptr) ? wide_ok(a->out, a->ldo) : 0;
  p.ldg = a->ldg; p.ldx = a->ldx; p.ldo = a->ldo;
  p.stats = a->stats; p.weight = a->weight; p.bias = a->bias; p.gelu = a->gelu ? (gelu_exact() ? 2 : 1) : 0; p.red = a->red;
  p.row_scale = a->row_scale; p.col_scale = a->col_scale; p.film_gamma = a->film_gamma;
  p.film_T = a->film_T > 0 ? a->film_T : 1; p.add32 = a->add32;
  p.film_ld = a->film_ld > 0 ? a->film_ld : a->C;
  BF_REQUIRE(p.film_ld >= a->C, "bf_inorm_bwd: film_ld < C");
  p.dweight = a->dweight; p.dbias = a->dbias; p.dcol_scale = a->dcol_scale;
  p.dfilm_gamma = a->dfilm_gamma; p.dfilm_beta = a->dfilm_beta;
  BF_REQUIRE((a->dweight == nullptr) == (a->dbias == nullptr), "bf_inorm_bwd: dweight / dbias pair");
  BF_REQUIRE((a->dfilm_gamma == nullptr) == (a->dfilm_beta == nullptr), "bf_inorm_bwd: film gradient pair");
  BF_REQUIRE(a->dweight != nullptr || (a->dcol_scale == nullptr && a->dfilm_gamma == nullptr),
             "bf_inorm_bwd: dcol_scale / dfilm gradients need dweight / dbias as well");
  dim3 grid(p.g.splits, a->I, p.g.chunks);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int gd = a->g_dtype, xd = a->x_dtype, od = a->out_dtype;
  const bool gelu = a->gelu != 0;
  if (a->phase == 1) {
#define BF_RED(TG, TX)                                                                                               \
  do {                                                                                                               \
    if (gelu) BF_NORM_LAUNCH((inorm_bwd_reduce_kernel<TG, TX, true>), grid, s, (const TG*)a->gin, (const TX*)a->x, p);  \
    else BF_NORM_LAUNCH((inorm_bwd_reduce_kernel<TG, TX, false>), grid, s, (const TG*)a->gin, (const TX*)a->x, p);      \
  } while (0)
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
  BF_REQUIRE((reinterpret_cast<uintptr_t>(a->out) & 15) == 0, "bf_inorm_bwd: out must be 16-byte aligned");
  BF_REQUIRE(a->add32 == nullptr || od == BF_F32, "bf_inorm_bwd: add32 needs f32 out");
  const bool add = a->add32 != nullptr;
#define BF_APP(TG, TX, TO, ADD_)                                                                                     \
  do {                                                                                                               \
    if (gelu) BF_NORM_LAUNCH((inorm_bwd_apply_kernel<TG, TX, TO, true, ADD_>), grid, s, (const TG*)a->gin, (const TX*)a->x, (TO*)a->out, p);  \
    else BF_NORM_LAUNCH((inorm_bwd_apply_kernel<TG, TX, TO, false, ADD_>), grid, s, (const TG*)a->gin, (const TX*)a->x, (TO*)a->out, p);      \
  } while (0)
  if (gd == BF_F32 && xd == BF_F32 && od == BF_F32) { if (add) BF_APP(float, float, float, true); else BF_APP(float, float, float, false); }
  else if (gd == BF_BF16 && xd == BF_F32 && od == BF_F32) { if (add) BF_APP(__nv_bfloat16, float, float, true); else BF_APP(__nv_bfloat16, float, float, false); }
  else if (add) BF_REQUIRE(false, "bf_inorm_bwd: add32 is supported for (f32|bf16, f32) -> f32 only (g=%d x=%d)", gd, xd);
  else if (gd == BF_F32 && xd == BF_BF16 && od == BF_BF16) BF_APP(float, __nv_bfloat16, __nv_bfloat16, false);
  else if (gd == BF_F32 && xd == BF_F16 && od == BF_F16) BF_APP(float, __half, __half, false);
  else if (gd == BF_BF16 && xd == BF_BF16 && od == BF_BF16) BF_APP(__nv_bfloat16, __nv_bfloat16, __nv_bfloat16, false);
  else if (gd == BF_F16 && xd == BF_F16 && od == BF_F16) BF_APP(__half, __half, __half, false);
  else if (gd == BF_BF16 && xd == BF_F16 && od == BF_BF16) BF_APP(__nv_bfloat16, __half, __nv_bfloat16, false);
  else if (gd == BF_F32 && xd == BF_F16 && od == BF_BF16) BF_APP(float, __half, __nv_bfloat16, false);
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
  launch_k(inorm_bwd_params_kernel, dim3((a->C + 31) / 32), dim3(256), (size_t)(0), static_cast<cudaStream_t>(stream), k);
  count_launch();
  BF_LAUNCH_CHECK("inorm_bwd_params_kernel");
  return BF_OK;
}

extern "C" int bf_resid_bwd(const float* dx, int64_t lddx, const void* z16, void* dz16, int64_t ldz, int dtype,
                            int I, int P, int C, const float* row_scale, const float* coef, float* S0, float* S1,
                            void* stream) {
  BF_REQUIRE(dx && coef && S0, "bf_resid_bwd: null pointer");
  BF_REQUIRE(z16 == nullptr || S1 != nullptr, "bf_resid_bwd: S1 required with z16");
  BF_REQUIRE(dtype == BF_BF16 || dtype == BF_F16 || dtype == BF_F32, "bf_resid_bwd: dtype");
  if (int st = check_common("bf_resid_bwd", I, P, C, lddx, dx)) return st;
  BF_REQUIRE((z16 == nullptr && dz16 == nullptr) || (ldz >= C && ldz % 8 == 0), "bf_resid_bwd: ldz");
  ResidBwdParams p{};
  p.g = make_geom(I, P, C, 4 + (z16 != nullptr ? es_of(dtype) : 0));
  p.wide = dz16 != nullptr ? wide_ok(dz16, ldz) : 0;
  p.lddx = lddx; p.ldz = ldz;
  p.dx = dx; p.z16 = z16; p.row_scale = row_scale; p.coef = coef; p.dz16 = dz16; p.S0 = S0; p.S1 = S1;
  dim3 grid(p.g.splits, I, p.g.chunks);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dtype == BF_BF16) {
    if (z16) BF_NORM_LAUNCH((resid_bwd_kernel<__nv_bfloat16, true>), grid, s, p);
    else BF_NORM_LAUNCH((resid_bwd_kernel<__nv_bfloat16, false>), grid, s, p);
  } else if (dtype == BF_F32) {            // fp32 validation backend
    if (z16) BF_NORM_LAUNCH((resid_bwd_kernel<float, true>), grid, s, p);
    else BF_NORM_LAUNCH((resid_bwd_kernel<float, false>), grid, s, p);
  } else {
    if (z16) BF_NORM_LAUNCH((resid_bwd_kernel<__half, true>), grid, s, p);
    else BF_NORM_LAUNCH((resid_bwd_kernel<__half, false>), grid, s, p);
  }
  count_launch();
  BF_LAUNCH_CHECK("resid_bwd_kernel");
  return BF_OK;
}

extern "C" int bf_colsum16(const void* x, int dtype, int64_t rows, int C, int64_t ldx, float* out, void* stream) {
  BF_REQUIRE(x && out, "bf_colsum16: null pointer");
  BF_REQUIRE(dtype == BF_BF16 || dtype == BF_F16 || dtype == BF_F32, "bf_colsum16: dtype");
  BF_REQUIRE(rows > 0 && rows < (1ll << 31), "bf_colsum16: rows");
  if (int st = check_common("bf_colsum16", 1, (int)rows, C, ldx, x)) return st;
  const Geom g = make_geom(1, (int)rows, C, es_of(dtype));
  dim3 grid(g.splits, 1, g.chunks);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dtype == BF_BF16) BF_NORM_LAUNCH(colsum16_kernel<__nv_bfloat16>, grid, s, (const __nv_bfloat16*)x, (long)ldx, g, out);
  else if (dtype == BF_F32) BF_NORM_LAUNCH(colsum16_kernel<float>, grid, s, (const float*)x, (long)ldx, g, out);
  else BF_NORM_LAUNCH(colsum16_kernel<__half>, grid, s, (const __half*)x, (long)ldx, g, out);
  count_launch();
  BF_LAUNCH_CHECK("colsum16_kernel");
  return BF_OK;
}
