"""The likelihood oracle against the FastTree-verified constants of the reference's own tests and
against outputs of the unmodified reference function (tests/golden/likelihood)."""
import numpy as np
import pytest

from oracle.likelihood_oracle import log_likelihood
from tests._ll_cases import AA, fasttree_kats, golden_cases


@pytest.mark.parametrize("kat", fasttree_kats(), ids=lambda k: k[0])
def test_fasttree_verified_constants(kat):
    name, tree, msa, cmap, rates, pi1, Q1, pi2, Q2, ll_exp, lls_exp, dec = kat
    if Q2 is None:
        cmap = np.eye(len(rates))
    ll, lls = log_likelihood(tree, msa, cmap, rates, AA, pi1, Q1, pi2, Q2)
    np.testing.assert_almost_equal(ll, ll_exp, decimal=dec)
    if lls_exp is not None:
        np.testing.assert_almost_equal(lls, lls_exp, decimal=dec)


@pytest.mark.parametrize("i", range(len(golden_cases())))
def test_matches_reference_function(i):
    c = golden_cases()[i]
    ll, lls = log_likelihood(c["tree"], c["msa"], c["contact_map"], c["site_rates"], AA, c["pi1"], c["Q1"], c["pi2"],
                             c["Q2"])
    # the reference's reversible back end (eigendecomposition) and scipy's Pade expm agree to ~1e-9
    # relative on 400 x 400 models; the north star's tolerance for fp64 log-likelihoods is 1e-6
    assert abs(ll - c["ll"]) <= 1e-7 * abs(c["ll"])
    np.testing.assert_allclose(lls, c["lls"], rtol=1e-6, atol=1e-9)
