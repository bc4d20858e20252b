"""JTT-IPW closed-form estimate of the rate matrix (the optimiser's default initialisation).

Mirrors the reference's ``cherryml/estimation/_jtt_ipw.py:55-117`` on a count tensor that is
already resident on the device (a handful of reductions over ``[K, S, S]``; no kernel of its
own -- this sits between the two hot kernels and is listed as a "next" row in SURVEY.md 8f).
"""
from typing import Optional

import numpy as np
import torch


def jtt_ipw_from_counts(
    times,
    counts: torch.Tensor,
    mask: Optional[np.ndarray] = None,
    use_ipw: bool = True,
    pseudocounts: float = 1e-8,
    symmetrize_count_matrices: bool = True,
    max_time: Optional[float] = None,
) -> np.ndarray:
    t = torch.as_tensor(np.asarray(times, dtype=np.float64), device=counts.device)
    c = counts.to(torch.float64)
    if max_time is not None:
        keep = t <= max_time
        t, c = t[keep], c[keep]
    S = c.shape[-1]
    c = c + pseudocounts
    if symmetrize_count_matrices:
        c = (c + c.transpose(1, 2)) / 2.0
    if mask is not None:
        c = c * torch.as_tensor(np.asarray(mask, dtype=np.float64), device=c.device)
    eye = torch.eye(S, dtype=torch.float64, device=c.device)
    F = c.sum(dim=0)
    F_off = F * (1.0 - eye)
    ctps = F_off / F_off.sum(dim=1, keepdim=True)
    if use_ipw:
        off_rows = (c * (1.0 - eye)).sum(dim=2)  # [K, S]
        M = (off_rows / t[:, None]).sum(dim=0) / F.sum(dim=1)
    else:
        # np.median averages the two middle values of an even-sized grid (torch.median takes the lower one)
        M = 1.0 / float(np.median(t.cpu().numpy())) * F_off.sum(dim=1) / F.sum(dim=1)
    res = M[:, None] * ctps
    res = res - torch.diag(torch.diagonal(res)) - torch.diag(M)
    return res.cpu().numpy()


def _jtt_ipw_stage():
    import logging
    import os
    import time

    from .. import caching
    from ..io import read_count_matrices_array, read_mask_matrix, write_rate_matrix

    logger = logging.getLogger(__name__)

    @caching.cached_computation(output_dirs=["output_rate_matrix_dir"], write_extra_log_files=True)
    def jtt_ipw(
        count_matrices_path: str,
        mask_path: Optional[str],
        use_ipw: bool,
        output_rate_matrix_dir: str,
        normalize: bool = False,
        max_time: Optional[float] = None,
        pseudocounts: float = 1e-8,
        symmetrize_count_matrices: bool = True,
    ) -> None:
        """JTT-IPW estimator; drop-in for the reference's ``cherryml.estimation.jtt_ipw``
        (``estimation/_jtt_ipw.py:28-125``): same arguments, writes ``result.txt`` and
        ``profiling.txt`` into ``output_rate_matrix_dir``."""
        start_time = time.time()
        logger.info("Starting")
        from ..counting import device_result

        resident = device_result(os.path.dirname(count_matrices_path))
        if resident is not None and os.path.basename(count_matrices_path) == "result.txt":
            q, states, counts = resident
        else:
            q, states, counts_np = read_count_matrices_array(count_matrices_path)
            # a closed-form estimate on K x S x S numbers (no kernel on this path): computed where
            # the counts are -- on the device when the counting stage left them there, on the host
            # when they come from a file
            counts = torch.from_numpy(counts_np)
        mask = read_mask_matrix(mask_path).to_numpy() if mask_path is not None else None
        res = jtt_ipw_from_counts(q, counts, mask=mask, use_ipw=use_ipw, pseudocounts=pseudocounts,
                                  symmetrize_count_matrices=symmetrize_count_matrices, max_time=max_time)
        if normalize:
            from ._engine import solve_stationary_dist

            pi = solve_stationary_dist(res)
            res = res / (pi @ -np.diag(res))
        write_rate_matrix(res, states, os.path.join(output_rate_matrix_dir, "result.txt"))
        logger.info("Done!")
        with open(os.path.join(output_rate_matrix_dir, "profiling.txt"), "w") as f:
            f.write(f"Total time: {time.time() - start_time} seconds\n")

    return jtt_ipw


jtt_ipw = _jtt_ipw_stage()
