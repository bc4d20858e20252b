"""Batched per-site fit against goldens produced by the UNMODIFIED reference function
(tests/golden/make_golden_siterm.py) and against the reference's own recovery KATs
(_siterm/_site_specific_rate_matrix.py:1905-1920, 1984-1998)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from cherryml_b200.siterm import quantized_transitions_mle_vectorized_over_sites
from tests.conftest import GOLDEN

G = os.path.join(GOLDEN, "siterm")


def test_aa_with_initialisation_fp64_path():
    """With an initialisation the reference runs in fp64: tolerance 1e-6 relative."""
    g = np.load(os.path.join(G, "aa_init.npz"))
    r = quantized_transitions_mle_vectorized_over_sites(g["counts"], g["times"], num_epochs=30, initialization=g["init"])
    assert r["res"].shape == g["res"].shape
    ref_l, got_l = g["loss_per_epoch_per_site"], r["loss_per_epoch_per_site"]
    assert np.max(np.abs(got_l - ref_l) / np.abs(ref_l)) < 1e-6
    assert np.max(np.abs(r["loss_per_epoch"] - g["loss_per_epoch"]) / np.abs(g["loss_per_epoch"])) < 1e-6
    assert np.max(np.abs(r["res"] - g["res"])) < 1e-6 * np.max(np.abs(g["res"]))


def test_dna_without_initialisation_fp32_reference():
    """Without an initialisation the reference's parameters are fp32: tolerance 1e-4."""
    g = np.load(os.path.join(G, "dna_noinit.npz"))
    r = quantized_transitions_mle_vectorized_over_sites(g["counts"], g["times"], num_epochs=40, initialization=None)
    ref_l, got_l = g["loss_per_epoch_per_site"], r["loss_per_epoch_per_site"]
    assert np.max(np.abs(got_l - ref_l) / np.abs(ref_l)) < 1e-4
    assert np.max(np.abs(r["res"] - g["res"])) < 1e-3 * np.max(np.abs(g["res"]))


def test_zero_epochs_returns_the_initialisation():
    g = np.load(os.path.join(G, "dna_init_0epochs.npz"))
    r = quantized_transitions_mle_vectorized_over_sites(g["counts"], g["times"], num_epochs=0, initialization=g["init"])
    assert np.max(np.abs(r["res"] - g["res"])) < 1e-6
    assert np.mean((r["res"] - g["init"]) ** 2) < 1e-6  # the reference's own assertion


def test_recovers_the_true_rate_matrices():
    """The reference's KAT: counts := expm(t Q_true); 100 epochs from a random start recover
    Q_true with mean squared error < 1e-3."""
    g = np.load(os.path.join(G, "dna_noinit.npz"))
    r = quantized_transitions_mle_vectorized_over_sites(g["counts"], g["times"], num_epochs=100, initialization=None)
    assert np.mean((g["Q_true"] - r["res"]) ** 2) < 1e-3


def _ref_raw_counts(transitions, q, alphabet, include_reverse=True):
    """Literal restatement of the reference loop (_site_specific_rate_matrix.py:189-261) for the test."""
    from oracle.counting_oracle import quantization_idx

    a2i = {c: i for i, c in enumerate(alphabet)}
    L, B, S = len(transitions[0][0]), len(q), len(alphabet)
    out = np.zeros((L, B, S, S))
    qa = np.array(q)
    for x, y, t in transitions:
        b = quantization_idx(t, qa)
        if b is None:
            continue
        for l in range(L):
            xl, yl = a2i.get(x[l], -1), a2i.get(y[l], -1)
            if xl >= 0 and yl >= 0:
                out[l, b, xl, yl] += 1.0
    return (out + out.transpose(0, 1, 3, 2)) / 2.0 if include_reverse else out


def test_per_site_counts_reference_example():
    """The reference's in-module test (test_get_raw_count_matrices, :264-322)."""
    from cherryml_b200.siterm import get_raw_count_matrices

    transitions = [("AG", "BH", 0.35 + 0.36), ("EG", "FH", 0.49 + 0.410), ("CG", "DG", 0.17 + 0.28 + 0.01 + 0.02)]
    alphabet = ["-", "A", "B", "C", "D", "E", "F", "G", "H", "I", "J", "K", "L", "M"]
    q = [0.40, 0.80, 2.0]
    for rev in (True, False):
        got = get_raw_count_matrices(transitions, q, alphabet, include_reverse_transitions=rev)
        assert got.shape == (2, 3, 14, 14)
        assert np.array_equal(got, _ref_raw_counts(transitions, q, alphabet, rev))
    a2i = {c: i for i, c in enumerate(alphabet)}
    got = get_raw_count_matrices(transitions, q, alphabet)
    assert got[0, 0, a2i["C"], a2i["D"]] == 0.5 and got[1, 0, a2i["G"], a2i["G"]] == 1.0
    assert got[1, 1, a2i["G"], a2i["H"]] == 1.0 and got.sum() == 6.0


def test_per_site_counts_random():
    from cherryml_b200.siterm import get_raw_count_matrices

    rng = np.random.default_rng(1)
    alphabet = list("ARNDCQEGHILKMFPSTWYV")
    letters = np.array(alphabet + ["-", "X"])
    L, n = 331, 19
    transitions = [("".join(rng.choice(letters, L)), "".join(rng.choice(letters, L)), float(t))
                   for t in np.exp(rng.uniform(np.log(1e-3), np.log(30.0), n))]
    q = [0.03 * 1.1 ** (8 * i) for i in range(-8, 9)]
    got = get_raw_count_matrices(transitions, q, alphabet)
    assert np.array_equal(got, _ref_raw_counts(transitions, q, alphabet))


@pytest.mark.parametrize("name", ["aa", "aa_gap_state"])
def test_estimate_site_specific_rate_matrices_matches_reference(name):
    """The whole SiteRM stage (cherry++ transitions -> per-site counts -> pseudocounts ->
    compaction -> batched fit) against the UNMODIFIED reference function
    (tests/golden/make_golden_siterm_estimate.py); the reference fits in fp64 here."""
    import json

    from cherryml_b200.io import Tree
    from cherryml_b200.siterm import estimate_site_specific_rate_matrices_given_tree_and_site_rates

    g = np.load(os.path.join(G, f"estimate_{name}.npz"))
    m = json.loads(str(g["meta"]))
    tree = Tree()
    tree.add_nodes(m["names"])
    for i in range(1, len(m["names"])):
        tree.add_edge(m["names"][m["parent"][i]], m["names"][i], m["length"][i])
    r = estimate_site_specific_rate_matrices_given_tree_and_site_rates(
        tree=tree, site_rates=m["rates"], msa=m["msa"], alphabet=m["alphabet"], regularization_strength=m["lam"],
        regularization_rate_matrix=g["Q0"], quantization_points=m["grid"], optimization_num_epochs=m["epochs"],
        use_vectorized_cherryml_implementation=True,
    )
    assert r["res"].shape == g["res"].shape
    assert np.max(np.abs(r["res"] - g["res"])) < 1e-6 * np.max(np.abs(g["res"]))
    assert "time_get_raw_count_matrices" in r and "time_get_pseudocount_matrices" in r


def _dna_tree():
    """(((leaf_1:1,leaf_2:1):1):1,((leaf_3:1,leaf_4:1):1):1); -- the tree of the reference's SiteRM KATs."""
    from cherryml_b200.io import Tree

    t = Tree()
    t.add_nodes(["r", "a", "b", "c", "d", "leaf_1", "leaf_2", "leaf_3", "leaf_4"])
    t.add_edges([("r", "a", 1.0), ("a", "b", 1.0), ("b", "leaf_1", 1.0), ("b", "leaf_2", 1.0),
                 ("r", "c", 1.0), ("c", "d", 1.0), ("d", "leaf_3", 1.0), ("d", "leaf_4", 1.0)])
    return t


def _jc(states, off):
    import pandas as pd

    n = len(states)
    m = np.full((n, n), off)
    np.fill_diagonal(m, -1.0)
    return pd.DataFrame(m, index=states, columns=states)


def test_site_rate_estimation_reference_kats():
    """Known answers of the reference's test_learn_site_rate_matrix_with_site_rate_prior
    (_learn_site_rate_matrix.py:1020-1050) and ..._and_gaps (:1052-1106)."""
    from cherryml_b200.siterm import (estimate_site_rates, get_standard_site_rate_grid, get_standard_site_rate_prior,
                                      learn_site_rate_matrices)

    dna = ["A", "C", "G", "T"]
    grid, prior = get_standard_site_rate_grid(), get_standard_site_rate_prior()
    for leaves, expected in [("AACG", 0.62312361621777), ("ACGT", 0.8541314966877565), ("AAAA", 0.17651113509036334)]:
        states = {f"leaf_{i + 1}": ch for i, ch in enumerate(leaves)}
        (rate,) = estimate_site_rates(_dna_tree(), states, grid, prior, _jc(dna, 1.0 / 3.0))
        np.testing.assert_almost_equal(rate, expected)
    # flat prior on the power-of-two grid: degenerate optima, as the reference documents
    g2 = [2.0 ** i for i in range(-10, 10)]
    (rate,) = estimate_site_rates(_dna_tree(), {"leaf_1": "A", "leaf_2": "A", "leaf_3": "C", "leaf_4": "G"}, g2,
                                  [1.0] * 20, _jc(dna, 1.0 / 3.0))
    np.testing.assert_almost_equal(rate, 0.5)
    r = learn_site_rate_matrices(
        tree=_dna_tree(), leaf_states={"leaf_1": "A", "leaf_2": "-", "leaf_3": "A", "leaf_4": "A"},
        alphabet=dna + ["-"], regularization_rate_matrix=_jc(dna + ["-"], 0.25), regularization_strength=0.5,
        site_rate_grid=grid, site_rate_prior=prior, alphabet_for_site_rate_estimation=dna,
        rate_matrix_for_site_rate_estimation=_jc(dna, 1.0 / 3.0), use_vectorized_implementation=True,
    )
    np.testing.assert_almost_equal(r["learnt_site_rates"][0], 0.33164477502323253)
    expected_T = np.array([
        [-0.5652167201042175, 0.0038684408646076918, 0.003868441330268979, 0.0038684408646076918, 0.5536113381385803],
        [0.018508626148104668, -0.31188488006591797, 0.08713673055171967, 0.08713670074939728, 0.11910282075405121],
        [0.01850862428545952, 0.08713671565055847, -0.31188488006591797, 0.08713671565055847, 0.11910280585289001],
        [0.018508626148104668, 0.08713670074939728, 0.08713673055171967, -0.3118848204612732, 0.11910276859998703],
        [1.1817187070846558, 0.05313650146126747, 0.05313650518655777, 0.053136471658945084, -1.3411281108856201],
    ])
    np.testing.assert_array_almost_equal(r["learnt_rate_matrices"][0].T, expected_T.T, decimal=1)  # the reference's decimal


def test_public_api_with_tree_matches_reference():
    import json

    from cherryml_b200.io import Tree, read_rate_matrix
    from cherryml_b200.markov_chain import get_lg_path
    from cherryml_b200.siterm import learn_site_specific_rate_matrices

    g = np.load(os.path.join(G, "public_api_tree_given.npz"))
    m = json.loads(str(g["meta"]))
    tree = Tree()
    tree.add_nodes(m["names"])
    for i in range(1, len(m["names"])):
        tree.add_edge(m["names"][m["parent"][i]], m["names"][i], m["length"][i])
    r = learn_site_specific_rate_matrices(
        tree=tree, msa=m["msa"], alphabet=list("ARNDCQEGHILKMFPSTWYV"), regularization_rate_matrix=read_rate_matrix(
            get_lg_path()), regularization_strength=0.5, num_epochs=20, quantization_grid_num_steps=8)
    assert r["learnt_site_rates"] == g["site_rates"].tolist()
    assert np.max(np.abs(r["learnt_rate_matrices"] - g["res"])) < 1e-6 * np.max(np.abs(g["res"]))
    assert r["learnt_tree"] is tree


def test_public_api_without_tree_uses_fast_cherries(tmp_path):
    """tree=None: FastCherries supplies tree and site rates; the result equals running the two
    stages by hand, and just_run_fast_cherries returns the star-of-cherries tree only."""
    from cherryml_b200.io import read_rate_matrix, read_site_rates, read_tree, write_msa
    from cherryml_b200.markov_chain import get_lg_path
    from cherryml_b200.phylogeny_estimation import fast_cherries
    from cherryml_b200.siterm import (estimate_site_specific_rate_matrices_given_tree_and_site_rates,
                                      learn_site_specific_rate_matrices)
    from tests._fc_cases import load_cases, parse_msa

    case = next(c for c in load_cases() if c["name"] == "synthetic_n33_L100_R20")
    names, seqs = parse_msa(case["msa_text"])
    msa = dict(zip(names, seqs))
    aa = list("ARNDCQEGHILKMFPSTWYV")
    lg = read_rate_matrix(get_lg_path())
    only = learn_site_specific_rate_matrices(tree=None, msa=msa, alphabet=aa, regularization_rate_matrix=lg,
                                             just_run_fast_cherries=True)
    assert only["learnt_rate_matrices"] is None and len(only["learnt_tree"].children("root")) == 17
    r = learn_site_specific_rate_matrices(tree=None, msa=msa, alphabet=aa, regularization_rate_matrix=lg,
                                          num_epochs=15, quantization_grid_num_steps=8)
    # the site rates are FastCherries' on the MSA as write_msa renders it (the reference's hand-off)
    (tmp_path / "msas").mkdir()
    write_msa(msa, str(tmp_path / "msas" / "f.txt"))
    fast_cherries(msa_dir=str(tmp_path / "msas"), families=["f"], rate_matrix_path=get_lg_path(),
                  num_rate_categories=20, max_iters=50, num_processes=1, verbose=False,
                  output_tree_dir=str(tmp_path / "t"), output_site_rates_dir=str(tmp_path / "s"),
                  output_likelihood_dir=str(tmp_path / "l"))
    assert r["learnt_site_rates"] == read_site_rates(str(tmp_path / "s" / "f.txt"))
    assert r["learnt_tree"].edges() == read_tree(str(tmp_path / "t" / "f.txt")).edges()
    step = 1.1 ** (64 / 8)
    by_hand = estimate_site_specific_rate_matrices_given_tree_and_site_rates(
        tree=r["learnt_tree"], site_rates=r["learnt_site_rates"], msa=msa, alphabet=aa, regularization_strength=0.5,
        regularization_rate_matrix=lg.to_numpy(), quantization_points=[0.03 * step ** i for i in range(-8, 9)],
        optimization_num_epochs=15)
    assert np.array_equal(r["learnt_rate_matrices"], by_hand["res"])
    assert r["learnt_rate_matrices"].shape == (100, 20, 20)


def test_sharded_over_sites_equals_plain_run():
    """One-rank NCCL group through the site-sharded path (blocks + gather): same result as the plain call."""
    import torch
    import torch.distributed as dist

    g = np.load(os.path.join(G, "aa_init.npz"))
    plain = quantized_transitions_mle_vectorized_over_sites(g["counts"], g["times"], num_epochs=10,
                                                            initialization=g["init"])
    created = not dist.is_initialized()
    if created:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29534")
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    try:
        sharded = quantized_transitions_mle_vectorized_over_sites(g["counts"], g["times"], num_epochs=10,
                                                                  initialization=g["init"],
                                                                  process_group=dist.group.WORLD)
    finally:
        if created:
            dist.destroy_process_group()
    assert np.array_equal(sharded["res"], plain["res"])
    assert np.array_equal(sharded["loss_per_epoch_per_site"], plain["loss_per_epoch_per_site"])
