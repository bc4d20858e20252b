"""Thin wrapper of ``cherry_gemm_f64_batched`` (the FP64 tensor-core GEMM of the large-S fit),
used by the unit tests and the DMMA micro-benchmark."""
import torch

from .. import _lib


def gemm_f64_batched(A: torch.Tensor, B: torch.Tensor, trans_a=False, trans_b=False, C=None, ksplit=1):
    """``C[b] (+)= op(A[b]) op(B[b])``; A, B: CUDA fp64 ``[batch, n, n]``, n a multiple of 80."""
    lib = _lib.load()
    assert A.dtype == torch.float64 and B.dtype == torch.float64 and A.is_cuda
    A, B = A.contiguous(), B.contiguous()
    batch, n, _ = A.shape
    accumulate = C is not None
    if C is None:
        C = torch.empty_like(A)
    desc = torch.empty(lib.cherry_gemm_desc_bytes(batch), dtype=torch.uint8, device=A.device)
    partial = torch.empty((batch * ksplit, n, n), dtype=torch.float64, device=A.device) if ksplit > 1 else None
    rc = lib.cherry_gemm_f64_batched(_lib.ptr(A), _lib.ptr(B), _lib.ptr(C), n, batch, int(trans_a), int(trans_b),
                                     int(accumulate), ksplit, _lib.ptr(desc), _lib.ptr(partial),
                                     _lib.current_stream_ptr())
    _lib.check(rc, "cherry_gemm_f64_batched")
    return C
