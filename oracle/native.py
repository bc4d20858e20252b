"""ctypes wrapper of oracle/count_encoded.c (CPU ORACLE -- TEST INFRASTRUCTURE).

``count_batch_oracle(batch, grid, S, directed)`` runs the scalar reference restatement on an
encoded batch (any object with the CountBatch fields as numpy arrays).
"""
import ctypes
import os

import numpy as np

from .build_ref import C_LIB, build_c_oracle

_lib = None


def _load():
    global _lib
    if _lib is None:
        src = os.path.join(os.path.dirname(os.path.abspath(__file__)), "count_encoded.c")
        if not os.path.exists(C_LIB) or (
            os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(C_LIB)
        ):
            build_c_oracle()
        _lib = ctypes.CDLL(C_LIB)
        _lib.oracle_quantization_idx.restype = ctypes.c_int
        _lib.oracle_quantization_idx.argtypes = [ctypes.c_double, ctypes.c_void_p, ctypes.c_int]
        _lib.oracle_count_lg.restype = None
        _lib.oracle_count_lg.argtypes = [ctypes.c_void_p] * 6 + [ctypes.c_int64] + [ctypes.c_void_p] * 3 + [
            ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        _lib.oracle_count_co.restype = None
        _lib.oracle_count_co.argtypes = [ctypes.c_void_p] * 6 + [ctypes.c_int64] + [ctypes.c_void_p] * 2 + [
            ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    return _lib


def _p(a):
    return a.ctypes.data


def quantization_idx_c(t: float, grid: np.ndarray) -> int:
    grid = np.ascontiguousarray(grid, dtype=np.float64)
    return int(_load().oracle_quantization_idx(float(t), _p(grid), len(grid)))


def count_batch_oracle(batch, grid, S: int, directed: bool, pair_slice=None) -> np.ndarray:
    """fp64 counts [K,S,S] (lg) or [K,S*S,S*S] (co) of ``batch`` on one CPU thread."""
    lib = _load()
    grid = np.ascontiguousarray(sorted(grid), dtype=np.float64)
    K = len(grid)
    arrs = [np.ascontiguousarray(x) for x in (batch.msa, batch.fams, batch.pair_a, batch.pair_b,
                                              batch.pair_t, batch.pair_fam)]
    if pair_slice is not None:
        for i in (2, 3, 4, 5):
            arrs[i] = np.ascontiguousarray(arrs[i][pair_slice])
    n_pairs = len(arrs[2])
    aux = np.ascontiguousarray(batch.aux)
    if batch.kind == "lg":
        counts = np.zeros((K, S, S), dtype=np.float64)
        rv = np.ascontiguousarray(batch.rate_vals, dtype=np.float64)
        lib.oracle_count_lg(*[_p(a) for a in arrs], n_pairs, _p(rv), _p(aux), _p(grid), K, S,
                            int(directed), _p(counts))
    else:
        n = S * S
        counts = np.zeros((K, n, n), dtype=np.float64)
        lib.oracle_count_co(*[_p(a) for a in arrs], n_pairs, _p(aux), _p(grid), K, S,
                            int(directed), _p(counts))
    return counts
