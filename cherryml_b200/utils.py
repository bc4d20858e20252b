"""Alphabet, quantization grid semantics and family striping.

Host-side mirror of the reference's ``cherryml/utils.py`` (alphabet order :7-28,
``quantization_idx`` :35-56, ``get_process_args`` :59-67, ``get_families`` :80-95).
``quantization_idx`` here is the scalar host definition of the bucket semantics; the
product path evaluates the same fp64 expression on the GPU (csrc/count_kernels.cu,
``quantize_bucket``) and never calls this function for data.
"""
import bisect
import contextlib
import os
from typing import List, Optional, Sequence

amino_acids = list("ARNDCQEGHILKMFPSTWYV")


def get_amino_acids() -> List[str]:
    return amino_acids[:]


def quantization_idx(
    branch_length: float, quantization_points_sorted: Sequence[float]
) -> Optional[int]:
    """Nearest grid point in relative error; ``None`` when outside the grid."""
    q = quantization_points_sorted
    if branch_length < q[0] or branch_length > q[-1]:
        return None
    ub = bisect.bisect_left(q, branch_length)
    if ub == 0:
        return 0
    left, right = float(q[ub - 1]), float(q[ub])
    if branch_length / left - 1 < right / branch_length - 1:
        return ub - 1
    return ub


def quantization_idx_array(branch_lengths, quantization_points_sorted):
    """``quantization_idx`` for an array of branch lengths (same fp64 expressions, evaluated by
    numpy); -1 where the scalar function returns ``None``."""
    import numpy as np

    q = np.asarray(quantization_points_sorted, dtype=np.float64)
    t = np.asarray(branch_lengths, dtype=np.float64)
    ub = np.searchsorted(q, t, side="left")
    inside = (t >= q[0]) & (t <= q[-1])
    ubc = np.clip(ub, 1, len(q) - 1)
    left, right = q[ubc - 1], q[ubc]
    with np.errstate(divide="ignore", invalid="ignore"):
        choose_left = (t / left - 1) < (right / t - 1)
    idx = np.where(ub == 0, 0, np.where(choose_left, ubc - 1, ubc))
    return np.where(inside, idx, -1)


def get_process_args(process_rank: int, num_processes: int, all_args: List) -> List:
    """Rank ``r`` of ``P`` owns items ``r, r+P, r+2P, ...`` (the reference's striping)."""
    if not 0 <= process_rank < num_processes:
        return []  # no index i has i % P == r
    return list(all_args[process_rank::num_processes])


def get_families(msa_dir: str) -> List[str]:
    names = sorted(os.listdir(msa_dir))
    return [x.split(".")[0] for x in names if x.endswith(".txt")]


@contextlib.contextmanager
def pushd(new_dir):
    """Run the body with ``new_dir`` as the working directory (reference utils.py:70-77)."""
    previous = os.getcwd()
    os.chdir(new_dir)
    try:
        yield
    finally:
        os.chdir(previous)


def make_quantization_points(center: float, step: float, num_steps: int) -> List[str]:
    """Geometric grid as the ``"%.8f"`` strings the end-to-end pipeline uses
    (reference ``estimation_end_to_end/_cherry.py:267-272``)."""
    return [
        ("%.8f" % (center * step**i)) for i in range(-num_steps, num_steps + 1, 1)
    ]
