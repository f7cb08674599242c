"""ctypes binding of libbubbleformer_b200.so (the C ABI declared in include/bubbleformer_b200.h).

There is no fallback: if the shared library is missing the import raises, and every wrapper refuses
tensors that are not on a CUDA device.  Build with `python -m bubbleformer_b200.build`.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbubbleformer_b200.so")

BF_BF16, BF_F16 = 0, 1
A_ROWMAJOR, A_S2D, A_KM = 0, 1, 2
B_NK, B_KN = 0, 1
EPI_STORE16, EPI_GELU, EPI_RESID, EPI_DGELU, EPI_ACC32, EPI_ATOMIC32, EPI_D2S, EPI_STORE32 = range(8)


class BubbleformerB200Error(RuntimeError):
    pass


class GemmArgs(C.Structure):
    _fields_ = [
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("dtype", C.c_int32), ("a_mode", C.c_int32), ("b_mode", C.c_int32),
        ("epilogue", C.c_int32), ("split_k", C.c_int32), ("bn", C.c_int32), ("reserved0", C.c_int32),
        ("A", C.c_void_p), ("B", C.c_void_p),
        ("lda", C.c_int64), ("ldb", C.c_int64),
        ("s2d_images", C.c_int32), ("s2d_hin", C.c_int32), ("s2d_win", C.c_int32), ("s2d_cin", C.c_int32),
        ("d2s_h", C.c_int32), ("d2s_w", C.c_int32), ("d2s_cout", C.c_int32), ("rows_per_group", C.c_int32),
        ("bias", C.c_void_p), ("col_scale", C.c_void_p), ("col_shift", C.c_void_p), ("col_gamma", C.c_void_p),
        ("row_scale", C.c_void_p), ("in32", C.c_void_p), ("aux16", C.c_void_p),
        ("out16", C.c_void_p), ("out16b", C.c_void_p), ("out32", C.c_void_p),
        ("ldo", C.c_int64), ("ld32", C.c_int64),
    ]


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise BubbleformerB200Error(
            f"{LIB_PATH} not found: the CUDA extension is not built "
            "(run `python -m bubbleformer_b200.build`); there is no CPU or PyTorch fallback")
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
