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
