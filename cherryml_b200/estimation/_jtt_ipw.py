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
        M = 1.0 / torch.median(t) * F_off.sum(dim=1) / F.sum(dim=1)
    res = M[:, None] * ctps
    res = res - torch.diag(torch.diagonal(res)) - torch.diag(M)
    return res.cpu().numpy()
