"""ctypes binding of libbubbleformer_b200.so (the C ABI declared in include/bubbleformer_b200.h).

There is no fallback: if the shared library is missing the import raises, and every wrapper refuses
tensors that are not on a CUDA device.  Build with `python bubbleformer_b200/build.py`.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbubbleformer_b200.so")

BF_BF16, BF_F16 = 0, 1
A_ROWMAJOR, A_S2D, A_KM = 0, 1, 2
B_NK, B_KN, B_KN_S2D = 0, 1, 2
EPI_STORE16, EPI_GELU, EPI_RESID, EPI_DGELU, EPI_ACC32, EPI_ATOMIC32, EPI_D2S, EPI_STORE32, EPI_QKV_LN = range(9)
EPI_GELU_D, EPI_DMUL = 9, 10


class BubbleformerB200Error(RuntimeError):
    pass


class GemmArgs(C.Structure):
    _fields_ = [
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("dtype", C.c_int32), ("a_mode", C.c_int32), ("b_mode", C.c_int32),
        ("epilogue", C.c_int32), ("split_k", C.c_int32), ("bn", C.c_int32), ("ln_head_dim", C.c_int32),
        ("A", C.c_void_p), ("B", C.c_void_p),
        ("lda", C.c_int64), ("ldb", C.c_int64),
        ("s2d_images", C.c_int32), ("s2d_hin", C.c_int32), ("s2d_win", C.c_int32), ("s2d_cin", C.c_int32),
        ("d2s_h", C.c_int32), ("d2s_w", C.c_int32), ("d2s_cout", C.c_int32), ("rows_per_group", C.c_int32),
        ("bias", C.c_void_p), ("col_scale", C.c_void_p), ("col_shift", C.c_void_p), ("col_gamma", C.c_void_p),
        ("row_scale", C.c_void_p), ("in32", C.c_void_p), ("aux16", C.c_void_p),
        ("out16", C.c_void_p), ("out16b", C.c_void_p), ("out32", C.c_void_p),
        ("ldo", C.c_int64), ("ld32", C.c_int64),
        ("stats_out", C.c_void_p), ("ln_rstd", C.c_void_p), ("colsum_out", C.c_void_p),
    ]


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise BubbleformerB200Error(
            f"{LIB_PATH} not found: the CUDA extension is not built "
            "(run `python bubbleformer_b200/build.py`); there is no CPU or PyTorch fallback")
    lib = C.CDLL(LIB_PATH)
    lib.bf_last_error.restype = C.c_char_p
    lib.bf_version.restype = C.c_int
    lib.bf_launch_count.restype = C.c_int64
    return lib


lib = _load()


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = lib.bf_last_error().decode(errors="replace")
        raise BubbleformerB200Error(f"{what} failed (status {status}): {msg}")


def launch_count() -> int:
    return int(lib.bf_launch_count())


BF_F32 = 2


class InormApplyArgs(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("x_dtype", C.c_int32), ("out_dtype", C.c_int32),
        ("ldx", C.c_int64), ("ldo", C.c_int64),
        ("I", C.c_int32), ("P", C.c_int32), ("C", C.c_int32), ("gelu", C.c_int32),
        ("stats", C.c_void_p), ("weight", C.c_void_p), ("bias", C.c_void_p),
        ("film_gamma", C.c_void_p), ("film_beta", C.c_void_p),
        ("film_T", C.c_int32), ("film_ld", C.c_int32),
        ("resid_in", C.c_void_p), ("row_scale", C.c_void_p), ("col_gamma", C.c_void_p),
        ("out", C.c_void_p), ("stats_out", C.c_void_p),
        ("compute_stats", C.c_int32), ("reserved_", C.c_int32),
    ]


class InormBwdArgs(C.Structure):
    _fields_ = [
        ("phase", C.c_int32), ("gelu", C.c_int32),
        ("gin", C.c_void_p), ("g_dtype", C.c_int32), ("x_dtype", C.c_int32),
        ("x", C.c_void_p),
        ("ldg", C.c_int64), ("ldx", C.c_int64), ("ldo", C.c_int64),
        ("I", C.c_int32), ("P", C.c_int32), ("C", C.c_int32), ("out_dtype", C.c_int32),
        ("stats", C.c_void_p), ("weight", C.c_void_p), ("bias", C.c_void_p),
        ("red", C.c_void_p),
        ("row_scale", C.c_void_p), ("col_scale", C.c_void_p), ("film_gamma", C.c_void_p),
        ("film_T", C.c_int32), ("film_ld", C.c_int32),
        ("add32", C.c_void_p), ("out", C.c_void_p),
        ("dweight", C.c_void_p), ("dbias", C.c_void_p), ("dcol_scale", C.c_void_p),
        ("dfilm_gamma", C.c_void_p), ("dfilm_beta", C.c_void_p),
    ]


class InormBwdParamsArgs(C.Structure):
    _fields_ = [
        ("red", C.c_void_p),
        ("I", C.c_int32), ("P", C.c_int32), ("C", C.c_int32), ("film_T", C.c_int32),
        ("row_scale", C.c_void_p), ("col_scale", C.c_void_p), ("film_gamma", C.c_void_p),
        ("weight", C.c_void_p), ("bias", C.c_void_p),
        ("dweight", C.c_void_p), ("dbias", C.c_void_p), ("dcol_scale", C.c_void_p),
        ("dfilm_gamma", C.c_void_p), ("dfilm_beta", C.c_void_p),
    ]


class AttnArgs(C.Structure):
    _fields_ = [
        ("qkv", C.c_void_p), ("ld_qkv", C.c_int64),
        ("out", C.c_void_p), ("ld_out", C.c_int64),
        ("dout", C.c_void_p), ("ld_dout", C.c_int64),
        ("heads", C.c_int32), ("head_dim", C.c_int32), ("L", C.c_int32), ("accumulate", C.c_int32),
        ("n_seq", C.c_int64), ("inner", C.c_int64), ("outer_stride", C.c_int64),
        ("inner_stride", C.c_int64), ("tok_stride", C.c_int64),
        ("qn_w", C.c_void_p), ("qn_b", C.c_void_p), ("kn_w", C.c_void_p), ("kn_b", C.c_void_p),
        ("bias_emb", C.c_void_p), ("bucket", C.c_void_p), ("scale_factor", C.c_void_p),
        ("out_scale", C.c_float), ("prenorm", C.c_int32),
        ("d_qn_w", C.c_void_p), ("d_qn_b", C.c_void_p), ("d_kn_w", C.c_void_p), ("d_kn_b", C.c_void_p),
        ("d_bias_emb", C.c_void_p), ("d_scale_factor", C.c_void_p),
        ("rstd", C.c_void_p), ("d_qkv_bias", C.c_void_p),
        ("dtype", C.c_int32), ("reserved0", C.c_int32),
    ]


class BranchGradArgs(C.Structure):
    _fields_ = [
        ("S01", C.c_void_p), ("I", C.c_int32), ("E", C.c_int32),
        ("gamma", C.c_void_p),
        ("c", C.c_void_p), ("c1", C.c_void_p), ("c0", C.c_void_p), ("low", C.c_void_p), ("high", C.c_void_p),
        ("W", C.c_void_p), ("norm2_bias", C.c_void_p),
        ("d_gamma", C.c_void_p), ("d_out_bias", C.c_void_p), ("d_low", C.c_void_p), ("d_high", C.c_void_p),
        ("d_W", C.c_void_p), ("d_norm2_bias", C.c_void_p),
    ]


EXPORTS = ["bf_set_gelu_mode", "bf_get_gelu_mode", "bf_set_reserved_sms", "bf_film_fwd", "bf_film_bwd", "bf_window_gather", "bf_eikonal_sums", "bf_heatflux_rows", "bf_optim_step", "bf_feat_consts", "bf_branch_param_grads", "bf_last_error", "bf_version", "bf_launch_count", "bf_gemm", "bf_inorm_stats", "bf_inorm_apply",
           "bf_inorm_bwd", "bf_inorm_bwd_params", "bf_resid_bwd", "bf_colsum16", "bf_attention_fwd",
           "bf_attention_bwd", "bf_lploss_sums", "bf_lploss_bwd", "bf_patch_in", "bf_patch_out", "bf_patch_wgrad", "bf_s2d_gather", "bf_cast16", "bf_convert16"]

_vp, _i, _i64 = C.c_void_p, C.c_int, C.c_int64
lib.bf_gemm.argtypes = [C.POINTER(GemmArgs), _vp]
lib.bf_inorm_stats.argtypes = [_vp, _i, _i, _i, _i, _i64, _vp, _vp]
lib.bf_inorm_apply.argtypes = [C.POINTER(InormApplyArgs), _vp]
lib.bf_inorm_bwd.argtypes = [C.POINTER(InormBwdArgs), _vp]
lib.bf_inorm_bwd_params.argtypes = [C.POINTER(InormBwdParamsArgs), _vp]
lib.bf_resid_bwd.argtypes = [_vp, _i64, _vp, _vp, _i64, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]
lib.bf_window_gather.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i64, _i, _vp]
lib.bf_eikonal_sums.argtypes = [_vp, _vp, _i64, _i, _i, C.c_float, _vp]
lib.bf_heatflux_rows.argtypes = [_vp, _vp, _vp, _i64, _i64, _i, C.c_float, C.c_float, C.c_float, C.c_float, _vp]
lib.bf_optim_step.argtypes = [_i, _vp, _vp, _vp, _vp, _vp, _i64, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                              _i64, _vp]
lib.bf_feat_consts.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp]
lib.bf_branch_param_grads.argtypes = [C.POINTER(BranchGradArgs), _vp]
lib.bf_set_gelu_mode.argtypes = [_i]
lib.bf_film_fwd.argtypes = [_vp, _i, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp]
lib.bf_film_bwd.argtypes = [_vp, _vp, _i, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp]
lib.bf_colsum16.argtypes = [_vp, _i, _i64, _i, _i64, _vp, _vp]
lib.bf_attention_fwd.argtypes = [C.POINTER(AttnArgs), _vp]
lib.bf_attention_bwd.argtypes = [C.POINTER(AttnArgs), _vp]
lib.bf_patch_in.argtypes = [_vp, _vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _vp]
lib.bf_patch_out.argtypes = [_vp, _i, _vp, _vp, _i, _i, _i, _i, _i, _vp]
lib.bf_patch_wgrad.argtypes = [_vp, _i, _vp, _vp, _i, _i, _i, _i, _i, _vp]
lib.bf_s2d_gather.argtypes = [_vp, _i, _vp, _i, _i, _i, _i, _i, _vp]
lib.bf_convert16.argtypes = [_vp, _i, _vp, _i, _i64, _vp]
lib.bf_cast16.argtypes = [_vp, _vp, _i, _i64, _vp]
lib.bf_lploss_sums.argtypes = [_vp, _vp, _vp, _i64, _i64, _vp]
lib.bf_lploss_bwd.argtypes = [_vp, _vp, _vp, _vp, _i64, _i64, _vp]
