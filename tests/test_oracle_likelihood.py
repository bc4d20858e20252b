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


def test_pair_site_model_on_real_data_fasttree_constant():
    """The reference's Test_real_data_pair_site_medium (likelihood_test.py:996-1068) at 4 rate
    categories: sites with the median rate coupled in pairs under WAG x WAG reproduce FastTree's
    single-site log-likelihood."""
    import os

    from cherryml_b200.io import Tree, read_msa, read_site_rates, read_tree
    from cherryml_b200.markov_chain import chain_product, compute_stationary_distribution
    from tests._ll_cases import LL_DIR, rate_matrix

    d = os.path.join(LL_DIR, "1a92")
    tree = read_tree(os.path.join(d, "tree_4_cat.txt"))
    msa = read_msa(os.path.join(d, "msa.txt"))
    site_rates = read_site_rates(os.path.join(d, "site_rates_4_cat.txt"))
    median = np.median(site_rates)
    places = [i for i, r in enumerate(site_rates) if r == median]
    np.random.seed(1)
    np.random.shuffle(places)
    cmap = np.eye(len(site_rates))
    for i in range(len(places) // 4):
        j, k = places[2 * i], places[2 * i + 1]
        cmap[j, k] = cmap[k, j] = 1
    assert cmap.sum() > len(site_rates)  # some sites are coupled
    scaled = Tree()
    scaled.add_nodes(tree.nodes())
    for u, v, length in tree.edges():
        scaled.add_edge(u, v, length * median)
    wag = rate_matrix("wag")
    wag2 = chain_product(wag, wag)
    ll, _ = log_likelihood(scaled, msa, cmap, [r / median for r in site_rates], AA,
                           compute_stationary_distribution(wag), wag, compute_stationary_distribution(wag2), wag2)
    np.testing.assert_almost_equal(ll, -4337.8688, decimal=4)
