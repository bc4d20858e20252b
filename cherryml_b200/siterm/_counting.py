"""Per-site transition counting for SiteRM: drop-in for the reference's
``_get_raw_count_matrices`` (``cherryml/_siterm/_site_specific_rate_matrix.py:189-261``)."""
from typing import List, Tuple

import numpy as np
import torch

from .. import _lib


def get_raw_count_matrices_device(
    transitions: List[Tuple[str, str, float]],
    quantization_points_sorted: List[float],
    alphabet: List[str],
    include_reverse_transitions: bool = True,
    device="cuda",
) -> torch.Tensor:
    """``[L, B, S, S]`` fp64 CUDA tensor of per-site counts of the cherries
    ``(seq_x, seq_y, t)``; the branch length is NOT scaled by a site rate (as in the reference)."""
    lib = _lib.load()
    dev = torch.device(device)
    S, B = len(alphabet), len(quantization_points_sorted)
    L = len(transitions[0][0])
    lut = np.full(256, S, dtype=np.uint8)
    for i, ch in enumerate(alphabet):
        lut[ord(ch)] = i
    n = len(transitions)
    for x, y, _ in transitions:
        assert len(x) == L and len(y) == L
    xa = lut[np.frombuffer("".join(x for x, _, _ in transitions).encode("latin-1"), dtype=np.uint8)]
    xb = lut[np.frombuffer("".join(y for _, y, _ in transitions).encode("latin-1"), dtype=np.uint8)]
    t = np.array([float(tt) for _, _, tt in transitions], dtype=np.float64)
    grid = np.asarray(quantization_points_sorted, dtype=np.float64)
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    xa_d, xb_d, t_d, g_d = d(xa), d(xb), d(t), d(grid)
    raw = torch.zeros((L, B, S, S), dtype=torch.int64, device=dev)
    out = torch.empty((L, B, S, S), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        stream = _lib.current_stream_ptr()
        _lib.check(lib.cherry_count_per_site(_lib.ptr(xa_d), _lib.ptr(xb_d), _lib.ptr(t_d), n, L, L, _lib.ptr(g_d),
                                             B, S, _lib.ptr(raw), stream), "cherry_count_per_site")
        _lib.check(lib.cherry_symmetrize_lg(_lib.ptr(raw), L * B, S, int(not include_reverse_transitions),
                                            _lib.ptr(out), stream), "cherry_symmetrize_lg")
    return out


def get_raw_count_matrices(transitions, quantization_points_sorted, alphabet,
                           include_reverse_transitions: bool = True) -> np.ndarray:
    return get_raw_count_matrices_device(transitions, quantization_points_sorted, alphabet,
                                         include_reverse_transitions).cpu().numpy()
