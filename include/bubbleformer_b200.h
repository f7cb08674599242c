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

/* ---- library ------------------------------------------------------------------------------- */
BF_API const char* bf_last_error(void);
BF_API int bf_version(void);
/* number of kernel launches issued through this library by the calling process (bench.py's gpu_launches) */
BF_API int64_t bf_launch_count(void);

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
  BF_B_KN = 1  /* B is (K, N) row-major: dgrad reads the same weight without a transposed copy; wgrad  */
};
enum bf_epilogue {
  BF_EPI_STORE16 = 0, /* out16 = acc + bias                                                            */
  BF_EPI_GELU = 1,    /* pre = acc + bias; out16 = gelu_erf(pre); out16b = pre (if non-null)           */
  BF_EPI_RESID = 2,   /* z = acc + bias; out16b = z (if non-null); v = z*col_scale + col_shift;
                         out32 = in32 + row_scale[m / rows_per_group] * col_gamma * v;
                         out16 = (16-bit) out32 (if non-null)                                          */
  BF_EPI_DGELU = 3,   /* out16 = acc * gelu_erf'(aux16)                                                */
  BF_EPI_ACC32 = 4,   /* out32 = in32 + acc                                                            */
  BF_EPI_ATOMIC32 = 5,/* atomicAdd(out32, acc): split-K wgrad accumulating straight into the fp32 grad */
  BF_EPI_D2S = 6,     /* out16 scattered depth-to-space: m = (img, y, x), n = (ky, kx, co) ->
                         out16[((img*2h + 2y+ky)*2w + 2x+kx)*cout + co]  (ConvTranspose2d k2 s2)        */
  BF_EPI_STORE32 = 7  /* out32 = acc + bias                                                            */
};

typedef struct bf_gemm_args {
  int32_t M, N, K;
  int32_t dtype;  /* BF_BF16 | BF_F16: storage type of A, B and of 16-bit outputs */
  int32_t a_mode; /* enum bf_a_mode */
  int32_t b_mode; /* enum bf_b_mode */
  int32_t epilogue;
  int32_t split_k; /* >= 1; > 1 only with BF_EPI_ATOMIC32 */
  int32_t bn;      /* 0 = choose the N tile automatically; 64/128/192/256 force it (tuning) */
  int32_t reserved0;
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
} bf_gemm_args;

BF_API int bf_gemm(const bf_gemm_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BUBBLEFORMER_B200_H_ */
