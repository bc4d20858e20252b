"""``matrix_exponential`` (both reference back ends) against torch.matrix_exp in fp64 and the
identity the reference tests directly (tests/evaluation_tests/likelihood_test.py:1264-1278)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from cherryml_b200.io import read_rate_matrix
from cherryml_b200.markov_chain import (FactorizedReversibleModel, chain_product, matrix_exponential,
                                        matrix_exponential_reversible)
from tests.conftest import GOLDEN

LG = os.path.join(GOLDEN, "fit", "inputs", "lg.txt")


def _ref(exponents, Q):
    t = torch.tensor(np.asarray(exponents, dtype=np.float64))
    return torch.matrix_exp(t[:, None, None] * torch.tensor(Q)).numpy()


@pytest.mark.parametrize("reversible", [False, True])
def test_matrix_exponential_lg(reversible):
    Q = read_rate_matrix(LG).to_numpy()
    exponents = [0.0, 1e-6, 0.003, 0.1, 1.0, 7.5, 40.0] + list(np.exp(np.linspace(-9, 3, 150)))
    fact = FactorizedReversibleModel(Q) if reversible else None
    got = matrix_exponential(np.array(exponents), Q, fact, reversible, "cuda")
    exp = _ref(exponents, Q)
    assert got.shape == (len(exponents), 20, 20)
    assert np.max(np.abs(got - exp)) < 1e-12
    pos = exp > 0  # t = 0 gives exact zeros off the diagonal
    assert np.array_equal(got[~pos], exp[~pos])
    assert np.max(np.abs(got[pos] - exp[pos]) / exp[pos]) < 1e-9  # small probabilities keep relative accuracy
    assert np.allclose(got.sum(axis=2), 1.0, atol=1e-12)


def test_matrix_exponential_400_states_and_product_identity():
    Q = read_rate_matrix(LG).to_numpy()
    QQ = chain_product(Q, Q)
    exponents = [0.01, 0.5, 2.0, 11.0]
    got = matrix_exponential(np.array(exponents), QQ, None, False, "cuda")
    single = _ref(exponents, Q)
    # expm(t (Q (+) Q)) = expm(tQ) (x) expm(tQ)
    for k in range(len(exponents)):
        assert np.max(np.abs(got[k] - np.kron(single[k], single[k]))) < 1e-12
        assert abs(got[k][0, 0] - single[k][0, 0] ** 2) < 1e-13
    rev = matrix_exponential_reversible(exponents, FactorizedReversibleModel(QQ), "cuda")
    assert np.max(np.abs(rev - got)) < 1e-10


def test_matrix_exponential_empty_and_out_of_range():
    from cherryml_b200 import _lib

    Q = read_rate_matrix(LG).to_numpy()
    assert matrix_exponential(np.array([]), Q, None, False, "cuda").shape == (0, 20, 20)
    with pytest.raises(_lib.CherryError):
        matrix_exponential(np.array([1e6]), chain_product(Q, Q), None, False, "cuda")
