"""CPU ORACLE for the quantized-transitions MLE fit -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import
this module.  The product (``cherryml_b200``) never does.

A restatement, on CPU tensors, of the reference's fit (songlab-cal/CherryML v0.2.0):

* reversible parameterisation  Q(theta)    estimation/_ratelearn/rate.py:167-188
* inversion of an initial Q into theta     estimation/_ratelearn/rate.py:61-91
* tensors / dtypes / seed / optimiser      estimation/_ratelearn/ratelearner.py:66-152
* the epoch loop (expm, loss, best iterate, power-of-two snapshots, backward, step)
                                           estimation/_ratelearn/trainer.py:118-243

The matrix exponential, its derivative, ``log`` and Adam are third-party code in the
reference too (``torch``, unpinned in its requirements.txt; 2.11.0 here): this oracle calls
the same library ops in the same order.  ``dtype=torch.float32`` reproduces the reference
as shipped (fp32 parameters and expm, fp64 loss); it is PINNED against goldens produced by
running the unmodified reference (tests/golden/make_golden_fit.py ->
tests/golden/fit/*/reference_run.npz, tests/test_oracle_fit.py).  ``dtype=torch.float64``
is the same computation with every tensor in fp64 -- what the reference would compute
without its two hard-coded float casts (ratelearner.py:98,107) -- and is the ground truth
for the fp64 CUDA path (tolerance 1e-6 relative, BASELINE.json north_star).
"""
import math
from typing import Dict, Optional

import numpy as np
import torch


def solve_stationary_dist(rate_matrix: np.ndarray) -> np.ndarray:
    """rate.py:9-17 (eigenvector of Q^T for the eigenvalue of smallest modulus)."""
    eigvals, eigvecs = np.linalg.eig(rate_matrix.transpose())
    index = np.argmin(np.abs(eigvals.real))
    pi = eigvecs.real[:, index]
    return pi / sum(pi)


def theta_from_initialization(init: np.ndarray, mask: np.ndarray):
    """rate.py:61-88: (log pi, softplus^-1 of the upper triangle of D^1/2 Q D^-1/2)."""
    S = init.shape[0]
    pi = solve_stationary_dist(init)
    if np.any(np.abs(pi) < 1e-8):
        raise ValueError("Stationary distribution of initialization is degenerate.")
    if np.any(np.abs(mask * init - init) > 1e-8):
        raise ValueError("initialization not compatible with mask")
    sym = np.diag(np.sqrt(pi)) @ init @ np.diag(1.0 / np.sqrt(pi))
    with np.errstate(divide="ignore"):
        vals = [np.log(np.exp(sym[i, j]) - 1) for i in range(S) for j in range(i + 1, S)]
    return np.log(pi), np.array(vals)


def rate_matrix(upper_diag: torch.Tensor, pi_logits: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """rate.py:167-188 (mode "pande_reversible")."""
    S = pi_logits.shape[0]
    rmat_off = torch.zeros(S, S, dtype=upper_diag.dtype)
    iu = torch.triu_indices(row=S, col=S, offset=1)
    rmat_off[iu[0], iu[1]] = torch.nn.functional.softplus(upper_diag)
    rmat_off = rmat_off + rmat_off.T
    rmat_off = rmat_off * mask
    pi = torch.softmax(pi_logits, dim=-1)
    pi_mat = torch.diag(pi.sqrt())
    pi_inv_mat = torch.diag(pi.sqrt() ** (-1))
    mat = (pi_inv_mat @ rmat_off) @ pi_mat
    mat = mat - torch.diag(mat.sum(1))
    return mat


def fit_oracle(
    times,
    counts: np.ndarray,
    mask: Optional[np.ndarray] = None,
    initialization: Optional[np.ndarray] = None,
    learning_rate: float = 0.1,
    num_epochs: int = 100,
    do_adam: bool = True,
    loss_normalization: bool = True,
    return_best_iter: bool = True,
    dtype=torch.float32,
) -> Dict[str, np.ndarray]:
    """Returns dict(loss[num_epochs], Q_1, Q_2, Q_4, ..., Q_best, Q_last, result)."""
    torch.manual_seed(0)  # ratelearner.py:77
    counts = np.asarray(counts, dtype=np.float64)
    S = counts.shape[1]
    qtimes = torch.tensor([float(t) for t in times], dtype=dtype)  # ratelearner.py:149 (fp32 there)
    cmats = torch.tensor(counts)  # fp64, ratelearner.py:150
    mask_np = np.ones((S, S)) if mask is None else np.asarray(mask, dtype=np.float64)
    mask_t = torch.tensor(mask_np, dtype=dtype)
    pi0 = torch.tensor(np.ones(S) / S).to(dtype)
    pi_logits = torch.log(pi0).clone().requires_grad_(True)  # rate.py:44-47, pi_requires_grad=True
    nparams_half = int(0.5 * S * (S - 1))
    upper_diag = (0.01 * torch.randn(nparams_half, dtype=torch.float32)).to(dtype).requires_grad_(True)
    if initialization is not None:
        log_pi, vals = theta_from_initialization(np.asarray(initialization, dtype=np.float64), mask_np)
        with torch.no_grad():
            pi_logits.copy_(torch.tensor(log_pi))
            upper_diag.copy_(torch.tensor(vals))
    # parameter order as registered by the module: _pi first, then upper_diag (rate.py:44-53)
    params = [pi_logits, upper_diag]
    opt = (torch.optim.Adam(params, lr=learning_rate) if do_adam
           else torch.optim.SGD(params, lr=learning_rate))
    out: Dict[str, np.ndarray] = {}
    losses = []
    best_loss, q_best = None, None
    Q = None
    for epoch in range(num_epochs):  # trainer.py:156-218
        opt.zero_grad()
        Q = rate_matrix(upper_diag, pi_logits, mask_t)
        mats = torch.log(torch.matrix_exp(qtimes[:, None, None] * Q))
        mats = mats * cmats
        loss = 0.0 + -1 / 1.0 * mats.sum()
        sample_size = 0.0 + cmats.sum()
        if loss_normalization:
            loss = loss / sample_size
        if best_loss is None or loss < best_loss:
            best_loss = loss
            q_best = Q.detach().numpy().copy()
        if (epoch & (epoch + 1)) == 0:
            out[f"Q_{epoch + 1}"] = Q.detach().numpy().copy()
        loss.backward()
        opt.step()
        losses.append(loss.item())
    out["loss"] = np.array(losses)
    if num_epochs > 0:
        out["Q_best"] = q_best.copy()
        out["Q_last"] = Q.detach().numpy().copy()
        out["result"] = q_best.copy() if return_best_iter else out["Q_last"].copy()
    return out


def loss_and_grad_oracle(Q: np.ndarray, times, counts: np.ndarray, dtype=torch.float64):
    """-sum C.log expm(tQ) / sum C and its gradient with respect to Q (autograd)."""
    Qt = torch.tensor(np.asarray(Q), dtype=dtype, requires_grad=True)
    t = torch.tensor([float(x) for x in times], dtype=dtype)
    C = torch.tensor(np.asarray(counts, dtype=np.float64))
    loss = -(torch.log(torch.matrix_exp(t[:, None, None] * Qt)) * C).sum() / C.sum()
    loss.backward()
    return float(loss.item()), Qt.grad.numpy().copy()
