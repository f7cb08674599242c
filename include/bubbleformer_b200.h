/*
 * bubbleformer_b200 -- C ABI of the B200-native FiLMAViT hot path.
 *
 * The upstream project (HPCForge/Bubbleformer) is pure Python on top of PyTorch: it has no FFI of its
 * own.  The "binding" this library replaces is therefore the set of torch library calls on the hot path
 * of FiLMConditionedAViT.forward / backward (upstream bubbleformer/models/axial_vit.py:217-242 and the
 * layers it composes).  Each entry point below names the upstream lines whose device work it performs.
 *
 * Conventions
 *   - plain C: raw DEVICE pointers, explicit sizes / leading dimensions (in elements), a cudaStream_t
 *     passed as void*.  No allocation, no implicit synchronisation, no global mutable state.
 *   - every function returns 0 on success, a BF_ERR_* code otherwise; bf_last_error() returns a
 *     thread-local message for the last failure on the calling thread.
 *   - activations are token-major / channels-last: one *image* is a (b, t) frame of P = h*w tokens,
 *     a token tensor is a row-major (I*P, C) matrix.  16-bit storage is bf16 (blocks) or fp16 (stem/head),
 *     selected by `dtype`; statistics, residual stream and gradients are fp32.
 *   - all kernels are compiled for sm_100a only; there is no CPU or other-architecture fallback.
 */
#ifndef BUBBLEFORMER_B200_H_
#define BUBBLEFORMER_B200_H_

#include <stdint.h>

#if defined(BF_BUILDING)
#define BF_API __attribute__((visibility("default")))
#else
#define BF_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define BF_OK 0
#define BF_ERR_INVALID 1  /* bad argument (shape / alignment / unsupported combination) */
#define BF_ERR_CUDA 2     /* a CUDA runtime or driver call failed */

#define BF_BF16 0
#define BF_F16 1
#define BF_F32 2

/* ---- library ------------------------------------------------------------------------------- */
BF_API const char* bf_last_error(void);
BF_API int bf_version(void);
/* number of kernel launches issued through this library by the calling process (bench.py's gpu_launches) */
BF_API int64_t bf_launch_count(void);

/* GELU.  Upstream uses nn.GELU() = exact erf (layers/linear_layers.py:16, layers/patching.py:47,103).  The default
 * here is the tanh form  0.5 x (1 + tanh(sqrt(2/pi) (x + 0.044715 x^3)))  with one MUFU.TANH: |tanh form - erf form|
 * <= 4.8e-4 absolute, 2e-4 rel-L2 on N(0,1) pre-activations, an eighth of the rounding of the 16-bit value it is stored
 * as; the erf form costs 2-3x the epilogue ALU work of the MLP GEMMs.  bf_set_gelu_mode(1) (or BF_GELU_ERF=1 in the
 * environment) switches EVERY GELU and GELU' of the library (bf_gemm BF_EPI_GELU / BF_EPI_DGELU, bf_inorm_apply,
 * bf_inorm_bwd) to the exact erf form: the validation configuration.  Process-wide, read at launch time.           */
BF_API int bf_set_gelu_mode(int exact_erf);
BF_API int bf_get_gelu_mode(void);

/* SMs set aside for other work that runs NEXT TO these kernels: the NCCL all-reduce kernels of data-parallel training
 * (upstream: Lightning strategy="ddp", scripts/train.py:163).  Every kernel here sizes its grid as one resident wave /
 * one persistent CTA per SM with a static tile order; a CTA that cannot become resident because an NCCL CTA holds its SM
 * only starts when another CTA has finished, i.e. the whole launch takes twice as long.  With n SMs reserved the grids
 * are sized for (SM count - n), so they stay resident beside n communication CTAs (NCCL_MAX_CTAS=n).  Process-wide,
 * read at launch time (set it before capturing a CUDA graph); 0 <= n <= 64, default 0.                           */
BF_API int bf_set_reserved_sms(int n);

/* fp32 validation.  The north star asks for parity within rel-L2 1e-4 "for the fp32 path" (upstream runs fp32 / TF32,
 * scripts/train.py:72).  The production kernels store operands in 16 bits and cannot reach that by construction, so every
 * entry point that moves 16-bit activations also accepts dtype = BF_F32: bf_gemm, bf_attention_fwd/bwd (args->dtype),
 * bf_patch_in / bf_patch_out / bf_patch_wgrad / bf_s2d_gather, bf_resid_bwd, bf_colsum16 (the InstanceNorm, loss and
 * optimiser passes are dtype-generic already).  Those forms are plain fp32 SIMT kernels (csrc/exact.cu): a checking
 * configuration for parity at 1e-4 together with bf_set_gelu_mode(1), not a performance path.                      */

/* ---- tcgen05 GEMM ---------------------------------------------------------------------------
 * D[M,N] = sum_k A[m,k] * B[n,k], fp32 accumulation in TMEM, operands staged by TMA.
 * Replaces: nn.Conv2d 1x1 input_head/output_head (upstream layers/attention.py:47-48,78,121,170-171,210,299),
 * nn.Linear fc1/fc2 (layers/linear_layers.py:14-15,25), Conv2d k2s2 stages 2.. of HMLPEmbed
 * (layers/patching.py:37-44) as implicit GEMM, ConvTranspose2d k2s2 stages of HMLPDebed
 * (layers/patching.py:93-99), and the autograd backward (dgrad / wgrad) of all of them.
 */
enum bf_a_mode {
  BF_A_ROWMAJOR = 0, /* A is (M, K) row-major, leading dimension lda                                 */
  BF_A_S2D = 1,      /* implicit 2x2/stride-2 patch gather from a channels-last image tensor
                        (images, hin, win, cin): M = images*hin/2*win/2, K = 4*cin ordered (ky, kx, ci) */
  BF_A_KM = 2        /* A is stored (K, M) row-major (contraction index outermost): wgrad, A = dY      */
};
enum bf_b_mode {
  BF_B_NK = 0, /* B is (N, K) row-major: a weight matrix as PyTorch stores it                         */
  BF_B_KN = 1, /* B is (K, N) row-major: dgrad reads the same weight without a transposed copy; wgrad  */
  BF_B_KN_S2D = 2 /* wgrad of a 2x2/stride-2 conv stage without a gathered copy: B[k, n] is the implicit patch gather of
                     a channels-last image tensor (s2d_* geometry), k = output pixel (img, yo, xo) -- the contraction
                     index, K = images*hin/2*win/2 -- and n = (ky, kx, ci), N = 4*cin.  Needs a_mode BF_A_KM,
                     (win/2) % 64 == 0 and (2*cin) % 64 == 0; A and B have the same 16-bit type (tcgen05 kind::f16 with an
                     fp16 operand against a bf16 one is an illegal instruction on sm_100a: measured).
                     layers/patching.py:37-44, 93-99 (autograd wgrad)                                            */
};
enum bf_epilogue {
  BF_EPI_STORE16 = 0, /* out16 = acc + bias                                                            */
  BF_EPI_GELU = 1,    /* pre = acc + bias; out16 = gelu(pre) (see GELU above); out16b = pre (if non-null) */
  BF_EPI_RESID = 2,   /* z = acc + bias; out16b = z (if non-null); v = z*col_scale + col_shift;
                         out32 = in32 + row_scale[m / rows_per_group] * col_gamma * v;
                         out16 = (16-bit) out32 (if non-null)                                          */
  BF_EPI_DGELU = 3,   /* out16 = acc * gelu'(aux16)                                                    */
  BF_EPI_ACC32 = 4,   /* out32 = in32 + acc                                                            */
  BF_EPI_ATOMIC32 = 5,/* out32 += acc (TMA reduce-add): split-K wgrad accumulating straight into the fp32 grad */
  BF_EPI_D2S = 6,     /* out16 scattered depth-to-space: m = (img, y, x), n = (ky, kx, co) ->
                         out16[((img*2h + 2y+ky)*2w + 2x+kx)*cout + co]  (ConvTranspose2d k2 s2)        */
  BF_EPI_STORE32 = 7, /* out32 = acc + bias                                                            */
  BF_EPI_GELU_D = 9,  /* pre = acc + bias; out16 = gelu(pre); out16b = gelu'(pre): the forward of fc1 as the training path
                         runs it -- the tanh / erf is evaluated once and the backward (BF_EPI_DMUL) is a multiply     */
  BF_EPI_DMUL = 10,   /* out16 = acc * aux16  (aux16 = gelu'(pre) saved by BF_EPI_GELU_D); colsum_out like BF_EPI_DGELU */
  BF_EPI_QKV_LN = 8   /* QKV projection with the per-head LayerNorm of q and k (upstream layers/attention.py:82,
                         214) fused in: columns are head*3d + [q | k | v], d = ln_head_dim = 64.  out16 receives
                         xhat_q = (q - mean)*rstd, xhat_k (no affine part: bf_attention applies it) and v;
                         ln_rstd[m][head][0..1] = rstd of the raw q / k rows (needed by the backward)          */
};

typedef struct bf_gemm_args {
  int32_t M, N, K;
  int32_t dtype;  /* BF_BF16 | BF_F16: storage type of A, B and of the 16-bit tensors (aux16, out16, out16b).
                     BF_F32: fp32 validation backend -- those tensors are float, the contraction is plain FFMA, every
                     epilogue except BF_EPI_QKV_LN is available, split_k is ignored (see "fp32 validation" below). */
  int32_t a_mode; /* enum bf_a_mode */
  int32_t b_mode; /* enum bf_b_mode */
  int32_t epilogue;
  int32_t split_k; /* >= 1; > 1 only with BF_EPI_ATOMIC32 */
  int32_t bn;      /* 0 = choose the N tile automatically; 64/128/192/256 force it (tuning) */
  int32_t ln_head_dim; /* BF_EPI_QKV_LN: head dimension (64) */
  const void* A;
  const void* B;
  int64_t lda, ldb;
  /* BF_A_S2D geometry of the *input* image tensor */
  int32_t s2d_images, s2d_hin, s2d_win, s2d_cin;
  /* BF_EPI_D2S geometry of the *input* token grid */
  int32_t d2s_h, d2s_w, d2s_cout, rows_per_group;
  /* epilogue operands (unused ones NULL) */
  const float* bias;      /* [N] */
  const float* col_scale; /* [N] */
  const float* col_shift; /* [N] */
  const float* col_gamma; /* [N] */
  const float* row_scale; /* [ceil(M / rows_per_group)] */
  const float* in32;      /* (M, N) ld = ld32 */
  const void* aux16;      /* (M, N) ld = ldo  */
  void* out16;            /* (M, N) ld = ldo  */
  void* out16b;           /* (M, N) ld = ldo  */
  float* out32;           /* (M, N) ld = ld32 */
  int64_t ldo, ld32;
  float* stats_out;       /* BF_EPI_RESID / BF_EPI_STORE16, may be NULL: stats_out[m / rows_per_group][n] += (sum, sum^2)
                             of out32 (RESID) or of the stored out16 (STORE16) -- the raw InstanceNorm statistics of
                             the tensor, so the norm that follows needs no separate pass (rows_per_group = tokens per
                             image, multiple of 32)                                                             */
  float* ln_rstd;         /* BF_EPI_QKV_LN only: (M, N / (3*ln_head_dim), 2) fp32                                */
  float* colsum_out;      /* BF_EPI_DGELU / BF_EPI_DMUL only, may be NULL: [N] += column sums of out16 (gradient of the fc1 bias) */
} bf_gemm_args;

BF_API int bf_gemm(const bf_gemm_args* args, void* stream);

/* ---- InstanceNorm over the token grid (per image, per channel) -------------------------------
 * Token-major tensors (I*P, C).  Statistics are raw sums: stats[img][c] = (sum x, sum x^2), fp32,
 * ACCUMULATED with atomics -- the caller zeroes the buffer first (cudaMemsetAsync).
 * Replaces nn.InstanceNorm2d(affine=True) (upstream layers/attention.py:39-40,153-154,197 and
 * layers/patching.py:45,102; eps 1e-5, biased variance) and the elementwise ops fused around it.
 */
BF_API int bf_inorm_stats(const void* x, int x_dtype, int I, int P, int C, int64_t ldx, float* stats, void* stream);

typedef struct bf_inorm_apply_args {
  const void* x;  int32_t x_dtype;  int32_t out_dtype;   /* BF_BF16 | BF_F16 | BF_F32 */
  int64_t ldx, ldo;
  int32_t I, P, C;
  int32_t gelu;               /* 1: y = gelu(y)  (layers/patching.py:47,103; see "GELU" below)            */
  const float* stats;         /* [I][C][2] */
  const float* weight;        /* [C] */
  const float* bias;          /* [C] */
  const float* film_gamma;    /* [I/film_T][film_ld] or NULL: y = film_gamma*y + film_beta (linear_layers.py:77) */
  const float* film_beta;
  int32_t film_T;  int32_t film_ld;   /* film_ld: row pitch of film_gamma / film_beta; 0 = C.  2C lets both point
                                         into the (B, 2C) output of bf_film_fwd (gamma = gb, beta = gb + C).       */
  const float* resid_in;      /* fp32 (I*P, C) ld = ldo or NULL: out = resid_in + row_scale[img]*col_gamma[c]*y
                                 (layer scale * drop-path + residual, layers/attention.py:317)           */
  const float* row_scale;     /* [I] or NULL */
  const float* col_gamma;     /* [C] (with resid_in) */
  void* out;
  float* stats_out;           /* [I][C][2] or NULL: += (sum, sum^2) of the values written to `out`, i.e. the raw
                                 statistics the NEXT InstanceNorm needs (saves its separate pass)               */
  int32_t compute_stats;      /* 1: `stats` is an OUTPUT (no zeroing needed): the call first computes the statistics of
                                 x and then applies them.  For plain calls (no gelu / film / residual) on block-sized
                                 tensors this is ONE launch: the slabs of an image form a thread-block cluster, exchange
                                 their partial sums through distributed shared memory and re-read x from L2.        */
  int32_t reserved_;
} bf_inorm_apply_args;
BF_API int bf_inorm_apply(const bf_inorm_apply_args* args, void* stream);

/* Backward of y = [gelu](IN(x)) given gin = dL/dy (times an optional per-(image, channel) scale
 * cs = row_scale[img]*col_scale[c]*film_gamma[img/film_T][c] that sat between y and the consumer).
 * phase 1: red[img][c] += (sum g, sum g*xhat), g = gin [* gelu'(.)]         (red zeroed by the caller)
 * phase 2: out = rstd*w*cs*(g - R1/P - xhat*R2/P) [+ add32]
 * phase 3: both; `red` is an output and need not be zeroed.  Without gelu / film on block-sized tensors this is ONE
 *          launch (one thread-block cluster per image, partial sums exchanged through distributed shared memory,
 *          second read of gin / x out of L2); otherwise it runs phase 1 and phase 2 back to back.             */
typedef struct bf_inorm_bwd_args {
  int32_t phase;  int32_t gelu;
  const void* gin;  int32_t g_dtype;  int32_t x_dtype;
  const void* x;
  int64_t ldg, ldx, ldo;
  int32_t I, P, C;  int32_t out_dtype;
  const float* stats;  const float* weight;  const float* bias;
  float* red;                 /* [I][C][2] */
  const float* row_scale;  const float* col_scale;  const float* film_gamma;
  int32_t film_T;  int32_t film_ld;   /* row pitch of film_gamma / dfilm_gamma / dfilm_beta; 0 = C */
  const float* add32;         /* fp32 (I*P, C) ld = ldo or NULL */
  void* out;
  /* phase 2, optional (all NULL = skip): the parameter gradients of bf_inorm_bwd_params accumulated by the same
   * launch with fp32 atomics.  dfilm_gamma / dfilm_beta are ADDED to here (the caller zeroes them).            */
  float* dweight;  float* dbias;  float* dcol_scale;  float* dfilm_gamma;  float* dfilm_beta;
} bf_inorm_bwd_args;
BF_API int bf_inorm_bwd(const bf_inorm_bwd_args* args, void* stream);

/* Parameter gradients from red (tiny): dweight[c] += sum_img cs*R2, dbias[c] += sum_img cs*R1,
 * dcol_scale[c] += sum_img row_scale[img]*(w*R2 + b*R1), dfilm_gamma[b][c] = sum_t (w*R2 + b*R1),
 * dfilm_beta[b][c] = sum_t R1.  NULL outputs are skipped.                                              */
typedef struct bf_inorm_bwd_params_args {
  const float* red;
  int32_t I, P, C, film_T;
  const float* row_scale;  const float* col_scale;  const float* film_gamma;
  const float* weight;  const float* bias;
  float* dweight;  float* dbias;  float* dcol_scale;  float* dfilm_gamma;  float* dfilm_beta;
} bf_inorm_bwd_params_args;
BF_API int bf_inorm_bwd_params(const bf_inorm_bwd_params_args* args, void* stream);

/* Residual-branch backward (layers/attention.py:123,309 reversed): one pass over the fp32 gradient stream
 *   dz16[m,c] = row_scale[img]*coef[c]*dx[m,c];  S0[img][c] += sum_{m in img} rs*dx;  S1[img][c] += sum rs*dx*z16[m,c]
 * S0 / S1 are PER IMAGE, (I, C) fp32 (the caller sums over images): per-channel totals would funnel every block's
 * atomics into C addresses.                                                                               */
BF_API int bf_resid_bwd(const float* dx, int64_t lddx, const void* z16, void* dz16, int64_t ldz, int dtype,
                        int I, int P, int C, const float* row_scale, const float* coef, float* S0, float* S1,
                        void* stream);

/* Feature-scaling constants of the axial block (upstream layers/attention.py:302-307).  The per-image mean of
 * z = IN(o) W^T + b is exactly c = W b_norm2 + b_out, so  z + mean(z)*low + (z - mean(z))*high == z*c1 + c0  with
 * c1 = 1 + high, c0 = c*(low - high).  W: output_head.weight (E, E) fp32; all outputs [E].  gamma / coef (both or
 * neither, may be NULL): coef = gamma*c1, the factor between dX_out and dZ that the backward pass needs.        */
BF_API int bf_feat_consts(const float* W, const float* norm2_bias, const float* out_bias, const float* low,
                          const float* high, const float* gamma, int E, float* c, float* c1, float* c0, float* coef,
                          void* stream);

/* Parameter gradients of one residual branch  X_out = X + mask*gamma*(Z*c1 + c0)  from the per-image sums of
 * bf_resid_bwd (S01 = (2, I, E): S0 = sum mask*dX_out, S1 = sum mask*dX_out*Z), accumulated in place:
 *   without feature scaling (c == NULL):  d_gamma += S1;  d_out_bias += gamma*S0
 *   with:  d_gamma += c1*S1 + c0*S0;  d_high += gamma*(S1 - c*S0);  d_low += gamma*c*S0;  dc = gamma*(low-high)*S0;
 *          d_out_bias += gamma*c1*S0 + dc;  d_W[j,:] += dc[j]*norm2_bias;  d_norm2_bias += W^T dc
 * (layers/attention.py:123 and :302-309 reversed).                                                            */
typedef struct {
  const float* S01;  int32_t I, E;
  const float* gamma;
  const float* c;  const float* c1;  const float* c0;  const float* low;  const float* high;
  const float* W;  const float* norm2_bias;
  float* d_gamma;  float* d_out_bias;  float* d_low;  float* d_high;  float* d_W;  float* d_norm2_bias;
} bf_branch_grad_args;
BF_API int bf_branch_param_grads(const bf_branch_grad_args* args, void* stream);

/* FiLM conditioning vector (upstream layers/linear_layers.py:58-61, 71-72: film_net = LayerNorm(F) -> Linear(F, 2E)):
 *   gb[b, :] = W * LayerNorm_F(cond[b, :]; ln_w, ln_b) + bias       cond (B, F) fp32, W (2E, F), gb (B, 2E)
 * gamma = gb[:, :E], beta = gb[:, E:] are applied by bf_inorm_apply (film_gamma / film_beta).  F <= 32.
 * bf_film_bwd ACCUMULATES the parameter gradients from dgb (B, 2E) (the caller zeroes them); cond gets no gradient
 * (the fluid parameters are data, dataset.py:170-178).  dc_scratch: (B, F) fp32 workspace, ZEROED by the caller
 * (receives d LayerNorm-output, reduced over the 2E outputs with one atomic per block).                           */
BF_API int bf_film_fwd(const float* cond, int B, int F, const float* ln_w, const float* ln_b, const float* W,
                       const float* bias, int E2, float* gb, void* stream);
BF_API int bf_film_bwd(const float* dgb, const float* cond, int B, int F, const float* ln_w, const float* ln_b,
                       const float* W, int E2, float* d_ln_w, float* d_ln_b, float* d_W, float* d_bias,
                       float* dc_scratch, void* stream);

/* Input pipeline (upstream data/dataset.py:120-186, BubbleForecast.__getitem__): the trajectories are resident in HBM
 * as frames (F, C_src, H*W) fp32; one launch cuts B windows of T frames, selects and normalises the fields and writes
 * out (B, T, C_out, H*W):  out[b,t,j,:] = (frames[first_frame[b] + t_off + t, channels[j], :] - diff[c]) * inv_div[c]
 * (t_off = 0 for the input window, T for the target window).  first_frame / channels / diff / inv_div are device arrays. */
BF_API int bf_window_gather(const float* frames, const int64_t* first_frame, const int32_t* channels, const float* diff,
                            const float* inv_div, float* out, int B, int T, int C_src, int C_out, int64_t HW, int t_off,
                            void* stream);

/* Rollout metrics (upstream utils/losses.py:5-15 eikonal_loss, utils/heatflux.py:3-38).
 * bf_eikonal_sums : sums[s] += sum_pixels (|grad phi| - 1)^2 for `slabs` contiguous (H, W) fp32 fields, torch.gradient
 *                   stencil (central inside, first-order one-sided at the edges), spacing dx (upstream: 1/32).
 * bf_heatflux_rows: flux[t] = mean over the W cells of the wall row (y = 0) of frame t of
 *                   [-5 <= x_centre <= 5 and dfun < 0] * (heater_temp - temp) * 0.054 / (dx * lc); frames are
 *                   `frame_stride` floats apart; x_centre = x_min + (i + 0.5) * dx (upstream: x_min -8, lc 0.0007).  */
BF_API int bf_eikonal_sums(const float* phi, float* sums, int64_t slabs, int H, int W, float dx, void* stream);
BF_API int bf_heatflux_rows(const float* dfun, const float* temp, float* flux, int64_t frames, int64_t frame_stride,
                            int W, float heater_temp, float x_min, float dx, float lc, void* stream);

/* Optimiser step over flat fp32 buffers (parameters, gradients, moments; n a multiple of 4), optionally refreshing
 * the bf16 operand mirror p16 in the same pass.  Upstream: bubbleformer/modules.py:132-142 (torch.optim.AdamW / Adam,
 * lion_pytorch.Lion with config/optim_cfg/{adamw,adam,lion}.yaml).  `step` counts from 1 (Adam bias correction);
 * v is unused (may be NULL) for BF_OPT_LION.                                                                    */
enum { BF_OPT_LION = 0, BF_OPT_ADAMW = 1, BF_OPT_ADAM = 2 };
BF_API int bf_optim_step(int kind, float* p, const float* g, float* m, float* v, void* p16, int64_t n, float lr,
                         float beta1, float beta2, float eps, float weight_decay, int64_t step, void* stream);

/* out[c] += sum_rows x[r, c] for a 16-bit matrix (bias gradients of the 1x1 convs / linears) */
BF_API int bf_colsum16(const void* x, int dtype, int64_t rows, int C, int64_t ldx, float* out, void* stream);

/* ---- fused 1-D attention along one axis of the token grid -----------------------------------
 * Replaces upstream layers/attention.py:80-101 (temporal) and :212-238 / :258-277 (axial x / y):
 * head split of the (tokens, 3E) QKV matrix (column = head*3d + [q | k | v]), LayerNorm(d) on q and k,
 * T5 relative-position bias (layers/positional_encoding.py:134-162), softmax(q k^T d^-1/2 + bias),
 * attn = 1/L + (softmax - 1/L)*scale_factor[head], attn @ v, written back token-major (tokens, E).
 * Sequence `s` (0 <= s < n_seq) consists of the L tokens
 *     base(s) + i*tok_stride,  base(s) = (s / inner)*outer_stride + (s % inner)*inner_stride
 * so that time / row / column attention need no permuted copies:
 *     temporal (B,T,P): n_seq = B*P, inner = P, outer_stride = T*P, inner_stride = 1, tok_stride = P, L = T
 *     x axis  (I,h,w):  n_seq = I*h, inner = h, outer_stride = P,   inner_stride = w, tok_stride = 1, L = w
 *     y axis  (I,h,w):  n_seq = I*w, inner = w, outer_stride = P,   inner_stride = 1, tok_stride = w, L = h
 * bucket[r], r = (j - i) + L - 1, is the T5 bucket index of key j relative to query i (host computed).
 * forward : out(tokens, E)    = out_scale * attention            (+= if accumulate)
 * backward: out(tokens, 3E)   = d qkv given dout(tokens, E)*out_scale (+= if accumulate); parameter
 *           gradients are atomically ACCUMULATED into d_* (fp32).  bf16 only, L <= 64, d in {32,48,64,96,128}.
 */
typedef struct bf_attn_args {
  const void* qkv;  int64_t ld_qkv;
  void* out;        int64_t ld_out;
  const void* dout; int64_t ld_dout;
  int32_t heads, head_dim, L, accumulate;
  int64_t n_seq, inner, outer_stride, inner_stride, tok_stride;
  const float* qn_w;  const float* qn_b;  const float* kn_w;  const float* kn_b;  /* [d] */
  const float* bias_emb;       /* [32][heads] */
  const int32_t* bucket;       /* [2L-1] */
  const float* scale_factor;   /* [heads] or NULL (attn_scale=False) */
  float out_scale;
  int32_t prenorm;             /* 1: qkv holds xhat_q | xhat_k | v as written by bf_gemm(BF_EPI_QKV_LN): the kernel
                                  applies the LayerNorm affine itself and the backward returns gradients w.r.t. the
                                  RAW q / k using `rstd`.  Fast paths: head_dim 64, L <= 32 (32-row tiles, short
                                  sequences packed) and 32 < L <= 64 (64-row tiles).                               */
  float* d_qn_w;  float* d_qn_b;  float* d_kn_w;  float* d_kn_b;
  float* d_bias_emb;  float* d_scale_factor;
  const float* rstd;           /* prenorm backward: (tokens, heads, 2) rstd of the raw q / k rows                  */
  float* d_qkv_bias;           /* prenorm backward, may be NULL: [3E] += column sums of the d qkv this launch writes
                                  (gradient of the input_head bias; replaces a separate bf_colsum16 pass)           */
  int32_t dtype;               /* storage type of qkv / out / dout: BF_BF16 (= 0, the production kernels) or BF_F32
                                  (fp32 validation backend, prenorm must be 0)                                      */
  int32_t reserved0;
} bf_attn_args;
BF_API int bf_attention_fwd(const bf_attn_args* args, void* stream);
BF_API int bf_attention_bwd(const bf_attn_args* args, void* stream);

/* ---- patch boundary: fp32 NCHW fields <-> 16-bit channels-last tokens -------------------------
 * bf_patch_in : out(I, H/2, W/2, N) = 2x2/stride-2 conv of x(I, F, H, W) with Wkn[(f,ky,kx)][n] (fp32, K x N);
 *               optional stats[img][n] += (sum, sum^2) of the stored values (feeds the following IN).
 *               First Conv2d of HMLPEmbed (upstream layers/patching.py:37-44); also d(input) of the last
 *               ConvTranspose2d of HMLPDebed.
 * bf_patch_out: out(I, F, 2h, 2w) fp32 = 2x2/stride-2 conv-transpose of a(I, h, w, C) with Wck[c][(f,ky,kx)].
 *               Last ConvTranspose2d of HMLPDebed (layers/patching.py:93-99); also d(input) of the first Conv2d.
 * bf_patch_wgrad: dW[n][(f,ky,kx)] += sum_pix a[pix][n] * x[img, f, 2y+ky, 2x+kx]  (weight gradient of both).
 * bf_s2d_gather: explicit im2col (I, Hin, Win, C) -> (I*Hin/2*Win/2, 4C), K order (ky, kx, ci), with optional
 *               fp16 <-> bf16 conversion on the way.
 * bf_cast16   : flat fp32 -> bf16/fp16 (operand copies of the fp32 master weights).
 */
BF_API int bf_patch_in(const float* x, const float* Wkn, void* out, int dtype, float* stats, int I, int F, int H,
                       int W, int N, void* stream);
BF_API int bf_patch_out(const void* a, int dtype, const float* Wck, float* out, int I, int F, int h, int w, int C,
                        void* stream);
BF_API int bf_patch_wgrad(const void* a, int dtype, const float* x, float* dW, int I, int F, int H, int W, int N,
                          void* stream);
BF_API int bf_s2d_gather(const void* in, int in_dtype, void* out, int out_dtype, int I, int Hin, int Win, int C,
                         void* stream);
/* fp16 <-> bf16 (the fp16 stem/head activations become bf16 operands of the backward GEMMs) */
BF_API int bf_convert16(const void* in, int in_dtype, void* out, int out_dtype, int64_t n, void* stream);
BF_API int bf_cast16(const float* in, void* out, int dtype, int64_t n, void* stream);

/* ---- relative-L2 training loss (the step right after the path) ------------------------
 * upstream utils/losses.py:67-94 with the configuration of modules.py:50 (d=2, p=2, mean b, mean t, sum c).
 * A slab is one (b, t, c) field of n_per_slab = H*W contiguous fp32 values.
 * bf_lploss_sums: sums[slab] += (sum (pred-tgt)^2, sum tgt^2)          (caller zeroes `sums`, finishes the scalar)
 * bf_lploss_bwd : dpred[slab, :] = coef[slab] * (pred - tgt)                                            */
BF_API int bf_lploss_sums(const float* pred, const float* tgt, float* sums, int64_t slabs, int64_t n_per_slab,
                          void* stream);
BF_API int bf_lploss_bwd(const float* pred, const float* tgt, const float* coef, float* dpred, int64_t slabs,
                         int64_t n_per_slab, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BUBBLEFORMER_B200_H_ */
