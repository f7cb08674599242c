"""Thin tensor-level wrappers over the C ABI: shape / dtype / device checks, then raw pointers.

PyTorch is used only for device memory and streams.  Every function requires CUDA tensors and raises
otherwise -- there is no eager fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib as L

_DT = {torch.bfloat16: L.BF_BF16, torch.float16: L.BF_F16}


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    if not t.is_cuda:
        raise L.BubbleformerB200Error("bubbleformer_b200 ops need CUDA tensors (no CPU fallback)")
    return t.data_ptr()


def _f32(t: Optional[torch.Tensor], n: int, name: str) -> Optional[int]:
    if t is None:
        return None
    if t.dtype != torch.float32 or not t.is_contiguous() or t.numel() < n:
        raise L.BubbleformerB200Error(f"{name}: expected contiguous float32 with >= {n} elements")
    return _ptr(t)


def gemm(A: torch.Tensor, B: torch.Tensor, M: int, N: int, K: int, *, epilogue: int,
         a_mode: int = L.A_ROWMAJOR, b_mode: int = L.B_NK, split_k: int = 1, bn: int = 0,
         lda: Optional[int] = None, ldb: Optional[int] = None,
         s2d: Optional[tuple] = None, d2s: Optional[tuple] = None, rows_per_group: int = 1,
         bias=None, col_scale=None, col_shift=None, col_gamma=None, row_scale=None,
         in32=None, aux16=None, out16=None, out16b=None, out32=None,
         ldo: Optional[int] = None, ld32: Optional[int] = None) -> None:
    """D[M,N] = sum_k A[m,k] B[n,k] with a fused epilogue; see bf_gemm in include/bubbleformer_b200.h."""
    if A.dtype not in _DT or B.dtype != A.dtype:
        raise L.BubbleformerB200Error(f"gemm: operands must both be bf16 or fp16, got {A.dtype}/{B.dtype}")
    a = L.GemmArgs()
    a.M, a.N, a.K = M, N, K
    a.dtype = _DT[A.dtype]
    a.a_mode, a.b_mode, a.epilogue, a.split_k, a.bn = a_mode, b_mode, epilogue, split_k, bn
    a.A, a.B = _ptr(A), _ptr(B)
    if lda is None:
        lda = A.stride(0) if (A.dim() == 2 and a_mode != L.A_S2D) else 0
    if ldb is None:
        ldb = B.stride(0) if B.dim() == 2 else B.shape[-1] * (B.shape[-2] if B.dim() == 3 else 1)
    a.lda, a.ldb = lda, ldb
    if s2d is not None:
        a.s2d_images, a.s2d_hin, a.s2d_win, a.s2d_cin = s2d
    if d2s is not None:
        a.d2s_h, a.d2s_w, a.d2s_cout = d2s
    a.rows_per_group = rows_per_group
    a.bias = _f32(bias, N, "bias")
    a.col_scale = _f32(col_scale, N, "col_scale")
    a.col_shift = _f32(col_shift, N, "col_shift")
    a.col_gamma = _f32(col_gamma, N, "col_gamma")
    a.row_scale = _f32(row_scale, (M + rows_per_group - 1) // rows_per_group, "row_scale")
    for name, t in (("aux16", aux16), ("out16", out16), ("out16b", out16b)):
        if t is not None and t.dtype != A.dtype:
            raise L.BubbleformerB200Error(f"gemm: {name} must have the operand dtype {A.dtype}")
    for name, t in (("in32", in32), ("out32", out32)):
        if t is not None and t.dtype != torch.float32:
            raise L.BubbleformerB200Error(f"gemm: {name} must be float32")
    a.in32, a.aux16 = _ptr(in32), _ptr(aux16)
    a.out16, a.out16b, a.out32 = _ptr(out16), _ptr(out16b), _ptr(out32)
    if ldo is None:
        ref = out16 if out16 is not None else (out16b if out16b is not None else aux16)
        ldo = ref.stride(0) if (ref is not None and ref.dim() == 2) else N
    if ld32 is None:
        ref = out32 if out32 is not None else in32
        ld32 = ref.stride(0) if (ref is not None and ref.dim() == 2) else N
    a.ldo, a.ld32 = ldo, ld32
    L.check(L.lib.bf_gemm(C.byref(a), _stream()), "bf_gemm")
