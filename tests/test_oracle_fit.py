"""The fit oracle is pinned here (CPU): in fp32 mode it must reproduce what the UNMODIFIED
reference produced (tests/golden/fit/*/reference_run.npz, written by
tests/golden/make_golden_fit.py) -- learned matrices bit-for-bit, losses to an ulp."""
import os

import numpy as np
import pytest
import torch

from cherryml_b200.io import read_count_matrices_array, read_mask_matrix, read_rate_matrix
from oracle.fit_oracle import fit_oracle, loss_and_grad_oracle, rate_matrix, theta_from_initialization
from tests.conftest import GOLDEN

FIT = os.path.join(GOLDEN, "fit")
INP = os.path.join(FIT, "inputs")
LG_COUNTS = os.path.join(GOLDEN, "counting/medium3/refcpp_count_matrices_dir_cherries_plus_plus/result.txt")
TOY = os.path.join(INP, "matrices_toy.txt")

# name -> (count matrices, initialization, mask)
SMALL_CASES = {
    "toy3_init": (TOY, "3x3_pande_reversible_initialization.txt", None),
    "toy3_init_mask": (TOY, "3x3_pande_reversible_initialization_mask.txt", "3x3_mask.txt"),
    "toy3_noinit": (TOY, None, None),
    "toy3_sgd": (TOY, "3x3_pande_reversible_initialization.txt", None),
    "lg20_init_equ": (LG_COUNTS, "equ.txt", None),
    "lg20_init_lg": (LG_COUNTS, "lg.txt", None),
    "lg20_noinit_mask": (LG_COUNTS, None, "20x20_random_mask.txt"),
}


def load_case(name):
    counts_path, init, mask = SMALL_CASES[name]
    q, states, counts = read_count_matrices_array(counts_path)
    init_a = read_rate_matrix(os.path.join(INP, init)).to_numpy() if init else None
    mask_a = read_mask_matrix(os.path.join(INP, mask)).to_numpy().astype(np.float64) if mask else None
    golden = np.load(os.path.join(FIT, name, "reference_run.npz"))
    return q, states, counts, init_a, mask_a, golden


@pytest.mark.parametrize("name", sorted(SMALL_CASES))
def test_fp32_oracle_reproduces_the_reference_run(name):
    q, _, counts, init, mask, g = load_case(name)
    out = fit_oracle(q, counts, mask, init, float(g["lr"]), int(g["num_epochs"]),
                     do_adam=("sgd" not in name), dtype=torch.float32)
    assert np.allclose(out["loss"], g["loss"], rtol=1e-14, atol=0)
    for key in g.files:
        if key.startswith("Q_") or key == "result":
            assert np.array_equal(out[key].astype(np.float32), g[key]), key


def test_fp64_oracle_is_within_fp32_tolerance_of_the_reference_run():
    q, _, counts, init, mask, g = load_case("lg20_init_equ")
    out = fit_oracle(q, counts, mask, init, float(g["lr"]), int(g["num_epochs"]), dtype=torch.float64)
    assert np.max(np.abs(out["loss"] - g["loss"]) / np.abs(g["loss"])) < 1e-5
    assert np.max(np.abs(out["result"] - g["result"])) < 1e-4 * np.max(np.abs(g["result"]))


def test_parameterisation_inverts_the_initialisation():
    lg = read_rate_matrix(os.path.join(INP, "lg.txt")).to_numpy()
    log_pi, upper = theta_from_initialization(lg, np.ones((20, 20)))
    Q = rate_matrix(torch.tensor(upper), torch.tensor(log_pi), torch.ones(20, 20, dtype=torch.float64)).numpy()
    assert np.allclose(Q, lg, atol=1e-6)
    assert np.allclose(Q.sum(axis=1), 0, atol=1e-12)


def test_loss_gradient_oracle_matches_finite_differences():
    rng = np.random.default_rng(0)
    S = 4
    Q = rng.random((S, S))
    np.fill_diagonal(Q, 0)
    np.fill_diagonal(Q, -Q.sum(axis=1))
    t = [0.05, 0.7, 3.0]
    C = rng.integers(0, 50, size=(3, S, S)).astype(float)
    loss, grad = loss_and_grad_oracle(Q, t, C)
    h = 1e-6
    for (i, j) in [(0, 1), (2, 2), (3, 0)]:
        Qp, Qm = Q.copy(), Q.copy()
        Qp[i, j] += h
        Qm[i, j] -= h
        fd = (loss_and_grad_oracle(Qp, t, C)[0] - loss_and_grad_oracle(Qm, t, C)[0]) / (2 * h)
        assert abs(fd - grad[i, j]) < 1e-6 * max(1.0, abs(grad[i, j]))
