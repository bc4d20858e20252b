"""Batched per-site fit (SiteRM): drop-in for the reference's
``quantized_transitions_mle_vectorized_over_sites``
(``cherryml/_siterm/_cherryml_vectorized.py:107-402``).

L independent rate matrices (one per site), B time buckets each, N <= 32 states, fitted
simultaneously on the GPU by the small-state fit kernels with ``n_problems = L``.

How the reference's parameterisation maps onto the engine's (``rate.py``-style) one:
the reference trains ``theta[L,N]`` and a FULL matrix ``Theta[L,N,N]`` and uses
``softplus(Theta + Theta^T)`` above the diagonal.  Both halves ``Theta_ij``, ``Theta_ji``
always receive the same gradient, hence the same Adam update, so ``u_ij = Theta_ij + Theta_ji``
evolves exactly like one parameter trained with twice the learning rate (Adam's update does
not depend on the gradient's scale).  The engine therefore runs with ``lr_pi = 0.1``,
``lr_upper = 0.2``, per-problem loss normalisation and per-problem best iterates starting
from ``+inf`` (``_cherryml_vectorized.py:341-372``).
"""
import logging
import time
from typing import Dict, List, Optional

import numpy as np
import torch

from ..estimation._engine import FitEngine

logger = logging.getLogger(__name__)


def solve_stationary_dist_fast(rate_matrices: np.ndarray) -> np.ndarray:
    """Stationary distributions by power iteration, as the reference initialises them
    (``_cherryml_vectorized.py:70-104``): fp32 ``matrix_exp`` of the diagonal-normalised
    matrices, then 100 squarings with row renormalisation.  Host-side, one-off.

    Same arithmetic per matrix, less of it: matrices whose fp32 normalised form is bit-identical
    (``rate_l * Q0`` for every site of a family collapses to a handful) are iterated once, and the
    loop stops when a squaring leaves every matrix bit-identical (the remaining squarings would
    reproduce it)."""
    diag_avg = np.mean(np.diagonal(rate_matrices, axis1=1, axis2=2), axis=1)
    normalized = rate_matrices * (-1.0 / diag_avg)[:, None, None]
    L, N, _ = normalized.shape
    as_f32 = np.ascontiguousarray(normalized.astype(np.float32).reshape(L, N * N))
    _, first, inverse = np.unique(as_f32.view(np.dtype((np.void, N * N * 4))).reshape(L), return_index=True,
                                  return_inverse=True)
    exp_matrices = torch.matrix_exp(torch.from_numpy(as_f32[first].reshape(-1, N, N))).numpy()
    for _ in range(100):
        squared = exp_matrices @ exp_matrices
        squared /= squared.sum(axis=2, keepdims=True)
        done = np.array_equal(squared, exp_matrices)
        exp_matrices = squared
        if done:
            break
    pi = exp_matrices[:, 0, :]
    pi /= pi.sum(axis=1, keepdims=True)
    return pi[inverse.reshape(-1)]


def _theta_from_initialization(initialization: np.ndarray) -> np.ndarray:
    L, N, _ = initialization.shape
    pi_all = solve_stationary_dist_fast(initialization)
    if not (np.allclose(pi_all.sum(axis=1), 1, atol=1e-3) and np.all(pi_all > 1e-8)):
        raise ValueError("At least one stationary distribution is degenerate.")
    sqrt_pi = np.sqrt(pi_all)
    S_all = sqrt_pi[:, :, None] * initialization / sqrt_pi[:, None, :]
    iu = np.triu_indices(N, k=1)
    upper = np.log(np.exp(S_all[:, iu[0], iu[1]]) - 1)  # = Theta_ij + Theta_ji
    return np.concatenate([np.log(pi_all).astype(np.float64), upper], axis=1)


def _random_theta(L: int, N: int, seed: int = 42) -> np.ndarray:
    """The reference's start without an initialisation: fp32 ``0.01*randn`` for ``theta`` then
    ``Theta`` from the CPU generator seeded with 42 (``:173-181, 307-321``)."""
    gen = torch.Generator(device="cpu")
    gen.manual_seed(seed)
    theta = 0.01 * torch.randn(L, N, generator=gen)
    Theta = 0.01 * torch.randn(L, N, N, generator=gen)
    iu = torch.triu_indices(N, N, offset=1)
    upper = (Theta + Theta.transpose(1, 2))[:, iu[0], iu[1]]
    return torch.cat([theta, upper], dim=1).double().numpy()


def quantized_transitions_mle_vectorized_over_sites(
    counts: np.ndarray,
    times: List[float],
    num_epochs: int,
    initialization: Optional[np.ndarray] = None,
    num_cores: int = 1,
    device: str = "cpu",
    process_group=None,
) -> Dict:
    """Estimate site-specific rate matrices from their count matrices.

    ``process_group`` (torch.distributed, one process per GPU): the sites are independent
    problems, so rank r fits the contiguous block ``site_blocks(L, world)[r]`` and the results are
    gathered -- no exchange step during training.  Every rank returns the
    full result.  Without an initialisation the random start of a site depends on its position in
    the batch (as in the reference), so sharded and unsharded runs agree only with one.

    ``counts``: ``L x B x N x N``; ``times``: ``L x B``.  Returns the reference's dictionary:
    ``res`` (``L x N x N``, per-site best iterate), ``loss_per_epoch``,
    ``loss_per_epoch_per_site`` and ``time_*`` entries.  ``device`` / ``num_cores`` are
    accepted for compatibility; the computation runs on a CUDA device."""
    if process_group is not None:
        return _fit_sharded_over_sites(counts, times, num_epochs, initialization, num_cores, device, process_group)
    st = time.time()
    prof = {}
    counts = np.asarray(counts, dtype=np.float64)
    times = np.asarray(times, dtype=np.float64)
    L, B, N, _ = counts.shape
    if N > 32:
        raise NotImplementedError("the batched per-site fit supports at most 32 states")
    logger.info(f"Going to estimate site rate matrices for L={L} sites, over N={N} states. "
                f"Number of time buckets: {B}.")
    prof["time_preamble"] = time.time() - st
    st = time.time()
    theta0 = _theta_from_initialization(np.asarray(initialization, dtype=np.float64)) \
        if initialization is not None else _random_theta(L, N)
    dev = torch.device(device if str(device).startswith("cuda") else "cuda")
    eng = FitEngine(times, counts, theta0, num_epochs=num_epochs, learning_rate=0.1, lr_upper=0.2,
                    do_adam=True, loss_normalization=True, best_mode=1, device=dev)
    if initialization is not None:
        np.testing.assert_almost_equal(eng.Q.cpu().numpy(), initialization, decimal=3)
    prof["time_initialize_model"] = time.time() - st
    st = time.time()
    eng.run()
    res = eng.results()
    prof["time_compute_loss"] = time.time() - st
    for key in ("time_send_counts_to_gpu", "time_initialize_tensors", "time_zero_grad", "time_get_Q",
                "time_cpu_loss_analysis", "time_backwards", "time_optimizer_step"):
        prof.setdefault(key, 0.0)
    per_site = res["loss_per_problem"] if num_epochs > 0 else np.zeros((0, L))
    out = {
        "res": res["Q_best"] if L > 1 else res["Q_best"][None] if res["Q_best"].ndim == 2 else res["Q_best"],
        "loss_per_epoch": per_site.sum(axis=1),
        "loss_per_epoch_per_site": per_site,
    }
    out.update(prof)
    return out


def site_blocks(num_sites: int, world_size: int) -> List[range]:
    """Contiguous, balanced blocks of sites, one per rank (the first ``num_sites % world_size``
    blocks are one site longer)."""
    base, extra = divmod(num_sites, world_size)
    blocks, start = [], 0
    for r in range(world_size):
        n = base + (1 if r < extra else 0)
        blocks.append(range(start, start + n))
        start += n
    return blocks


def _fit_sharded_over_sites(counts, times, num_epochs, initialization, num_cores, device, process_group) -> Dict:
    import torch.distributed as dist

    counts = np.asarray(counts, dtype=np.float64)
    times = np.asarray(times, dtype=np.float64)
    L = counts.shape[0]
    world, rank = dist.get_world_size(process_group), dist.get_rank(process_group)
    mine = site_blocks(L, world)[rank]
    local = None
    if len(mine):
        sl = slice(mine.start, mine.stop)
        local = quantized_transitions_mle_vectorized_over_sites(
            counts[sl], times[sl], num_epochs, None if initialization is None else np.asarray(initialization)[sl],
            num_cores, device)
    pieces = [None] * world
    dist.all_gather_object(pieces, None if local is None else
                           {k: local[k] for k in ("res", "loss_per_epoch_per_site")}, group=process_group)
    pieces = [p for p in pieces if p is not None]
    per_site = np.concatenate([p["loss_per_epoch_per_site"] for p in pieces], axis=1)
    out = {
        "res": np.concatenate([p["res"] for p in pieces], axis=0),
        "loss_per_epoch": per_site.sum(axis=1),
        "loss_per_epoch_per_site": per_site,
    }
    if local is not None:
        out.update({k: v for k, v in local.items() if k.startswith("time_")})
    return out
