"""Tensor-level wrappers over the C ABI: torch tensors in, raw device pointers + strides out.

PyTorch is used for device memory and streams only; every arithmetic step happens in libvitb200.so.
"""
import ctypes

import torch

from . import _lib
from ._lib import VbGemmDesc

EPI_STORE, EPI_GELU, EPI_RESIDUAL, EPI_RELU, EPI_DGELU, EPI_DRELU, EPI_ACCUM = range(7)
BF16, F32 = 0, 1


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dt(t):
    if t.dtype == torch.bfloat16:
        return BF16
    if t.dtype == torch.float32:
        return F32
    raise TypeError(f"unsupported dtype {t.dtype}")


def _mat(t, name):
    """Returns (ptr, ld, batch_stride, batches, rows, cols) of a 2-D or 3-D row-major tensor."""
    if t.dim() == 2:
        assert t.stride(1) == 1, f"{name}: innermost stride must be 1"
        return t.data_ptr(), t.stride(0), 0, 1, t.shape[0], t.shape[1]
    assert t.dim() == 3 and t.stride(2) == 1, f"{name}: expected [batch, rows, cols] with unit inner stride"
    return t.data_ptr(), t.stride(1), t.stride(0), t.shape[0], t.shape[1], t.shape[2]


def gemm(A, B, C, *, a_major=0, b_major=0, epilogue=EPI_STORE, bias=None, aux=None, C2=None, split_k=1,
         c_row_offset=0, aux_broadcast=False, max_ctas=0, direct=False):
    """C = epilogue(A @ B^T). ``A`` is [M,K] (a_major 0) or stored [K,M] (a_major 1); ``B`` is [N,K] or stored [K,N].

    3-D tensors add a leading batch dimension (B may stay 2-D to be shared). ``C``/``C2``/``aux`` are [M(+off),N].
    """
    lib = _lib.load()
    d = VbGemmDesc()
    pa, lda, bsa, nba, ra, ca = _mat(A, "A")
    pb, ldb, bsb, nbb, rb, cb = _mat(B, "B")
    M, K = (ra, ca) if a_major == 0 else (ca, ra)
    N, Kb = (rb, cb) if b_major == 0 else (cb, rb)
    assert K == Kb, f"contraction mismatch: A gives K={K}, B gives K={Kb}"
    assert A.dtype == torch.bfloat16 and B.dtype == torch.bfloat16
    out = C if C is not None else C2
    pc, ldc, bsc, nbc, rc, cc = _mat(out, "C")
    assert cc == N and nbc == nba, (cc, N, nbc, nba)
    d.M, d.N, d.K, d.batches = M, N, K, nba
    d.a_major, d.b_major, d.epilogue, d.c_dtype = a_major, b_major, epilogue, _dt(out)
    d.split_k, d.c_row_offset, d.c_rows = split_k, c_row_offset, rc
    d.aux_batch_broadcast = 1 if aux_broadcast else 0
    d.A, d.lda, d.batch_stride_a = pa, lda, bsa
    d.B, d.ldb, d.batch_stride_b = pb, ldb, (bsb if nbb > 1 else 0)
    if C is not None:
        d.C, d.ldc, d.batch_stride_c = pc, ldc, bsc
    if C2 is not None:
        p2, ld2, bs2, _, _, _ = _mat(C2, "C2")
        d.C2, d.ldc2, d.batch_stride_c2 = p2, ld2, bs2
    if aux is not None:
        assert aux.dtype == out.dtype, "AUX must have the dtype of C"
        px, ldx, bsx, _, _, _ = _mat(aux, "aux")
        d.AUX, d.ldaux, d.batch_stride_aux = px, ldx, bsx
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.is_contiguous() and bias.numel() == N
        d.bias = bias.data_ptr()
    d.max_ctas = max_ctas
    d.debug_direct_store = 1 if direct else 0
    _lib.check(lib.vb_gemm_bf16(ctypes.byref(d), _stream()), "vb_gemm_bf16")
    return out
