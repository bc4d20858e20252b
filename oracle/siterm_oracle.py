"""CPU restatement of the reference's batched per-site fit (TEST INFRASTRUCTURE).

Follows ``cherryml/_siterm/_cherryml_vectorized.py``: stationary distributions by fp32
``matrix_exp`` + 100 squarings (:70-104), inversion of the parameterisation (:190-236), the
rate matrices ``Q_l = D^-1/2 (softplus(Theta + Theta^T) off-diagonal, symmetrised) D^1/2`` with
``pi = softmax(theta)`` (:238-257), the loss ``sum_l -sum_b <C_lb, log expm(t_lb Q_l)> / sum C_l``
(:259-287), Adam(lr 0.1) and the per-site best iterate (:300-372).  torch on the CPU supplies
``matrix_exp``, autograd and Adam, as in the reference.
PINNED against tests/golden/siterm/aa_init.npz and dna_noinit.npz (outputs of the UNMODIFIED
reference function, tests/golden/make_golden_siterm.py) in tests/test_oracle_siterm.py.
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
from typing import Dict, Optional

import numpy as np
import torch


def stationary_distributions(rate_matrices: np.ndarray) -> np.ndarray:
    diag_mean = np.mean(np.diagonal(rate_matrices, axis1=1, axis2=2), axis=1)
    scaled = rate_matrices * (-1 / diag_mean)[:, None, None]
    P = torch.matrix_exp(torch.tensor(scaled, dtype=torch.float32)).numpy()
    for _ in range(100):
        P = P @ P
        P /= P.sum(axis=2, keepdims=True)
    pi = P[:, 0, :]
    return pi / pi.sum(axis=1, keepdims=True)


def fit_sites(counts: np.ndarray, times: np.ndarray, num_epochs: int, initialization: Optional[np.ndarray] = None,
              num_threads: Optional[int] = None) -> Dict:
    if num_threads:
        torch.set_num_threads(num_threads)
    counts_t = torch.tensor(np.asarray(counts))
    times_t = torch.tensor(np.asarray(times))
    L, B, N, _ = counts_t.shape
    torch.manual_seed(42)
    theta = 0.01 * torch.randn(L, N)
    Theta = 0.01 * torch.randn(L, N, N)
    if initialization is not None:
        pi = stationary_distributions(initialization)
        if not (np.allclose(pi.sum(axis=1), 1, atol=1e-3) and np.all(pi > 1e-8)):
            raise ValueError("At least one stationary distribution is degenerate.")
        sq = np.sqrt(pi)[:, :, None]
        S = (sq * np.eye(N)[None]) @ initialization @ ((1.0 / sq) * np.eye(N)[None])
        iu = np.triu_indices(N, k=1)
        Th = np.zeros_like(S)
        with np.errstate(divide="ignore"):
            Th[:, iu[0], iu[1]] = np.log(np.exp(S[:, iu[0], iu[1]]) - 1)
        Th = (Th + Th.transpose(0, 2, 1)) / 2.0
        theta = torch.tensor(np.log(pi), dtype=torch.float64)
        Theta = torch.tensor(Th, dtype=torch.float64)
    theta = theta.clone().requires_grad_(True)
    Theta = Theta.clone().requires_grad_(True)
    upper = torch.triu(torch.ones(N, N), diagonal=1)

    def rate_matrices():
        pi = torch.nn.functional.softmax(theta, dim=1)
        S = torch.nn.functional.softplus(Theta + Theta.transpose(1, 2)) * upper
        S = S + S.transpose(1, 2)
        off = torch.diag_embed(1.0 / pi.sqrt()) @ S @ torch.diag_embed(pi.sqrt())
        return off - torch.diag_embed(off.sum(dim=2))

    opt = torch.optim.Adam([theta, Theta], lr=0.1)
    loss_best = torch.full((L,), float("inf"))
    Q_best = rate_matrices().detach()
    per_epoch_site = np.zeros((num_epochs, L))
    per_epoch = np.zeros(num_epochs)
    total = counts_t.sum(dim=(1, 2, 3))
    for epoch in range(num_epochs):
        opt.zero_grad()
        Q = rate_matrices()
        logp = torch.log(torch.matrix_exp(times_t.view(L, B, 1, 1) * Q.unsqueeze(1)))
        per_site = -(counts_t * logp).sum(dim=(2, 3)).sum(dim=1) / total
        loss = per_site.sum()
        better = per_site < loss_best
        loss_best = torch.where(better, per_site, loss_best)
        Q_best = torch.where(better.view(-1, 1, 1), Q.detach().to(Q_best.dtype), Q_best)
        per_epoch_site[epoch] = per_site.detach().numpy()
        per_epoch[epoch] = float(loss.detach())
        loss.backward()
        opt.step()
    return {"res": Q_best.numpy(), "loss_per_epoch": per_epoch, "loss_per_epoch_per_site": per_epoch_site}
