"""The FastCherries oracle (oracle/fast_cherries_oracle.py) against (1) the known answers of the
reference's own C++ unit tests and (2) outputs of the unmodified reference program
(tests/golden/fast_cherries, see make_golden_fast_cherries*.py); plus host-side pieces of the
product that need no GPU (grid, categories, weights, MSA encoding, tree layout)."""
import math

import numpy as np
import pytest

from oracle import fast_cherries_oracle as fo
from tests._fc_cases import expected_outputs, load_cases, load_kats, msa_text, parse_msa, parse_rate_matrix

CASES = load_cases()
_TABLES = {}


def table_for(Q, q, cats):
    key = (Q.tobytes(), q.tobytes(), cats.tobytes())
    if key not in _TABLES:
        _TABLES[key] = fo.log_table_scipy(Q, q, cats)
    return _TABLES[key]


def test_reference_kats_branch_lengths():
    k = load_kats()
    Q = np.array(k["rate_matrix"])
    bl = k["branch_lengths"]
    q, cats = np.array(bl["grid"]), np.array(bl["rate_categories"])
    T = table_for(Q, q, cats)
    sym = T + np.swapaxes(T, 2, 3)
    for case in bl["cases"]:
        xa = np.array([c[0] for c in case["cherries"]])
        xb = np.array([c[1] for c in case["cherries"]])
        got = fo.branch_length_indices(xa, xb, sym, np.array(bl["site_to_rate"]))
        assert got.tolist() == case["expected"]


def test_reference_kats_site_rates():
    k = load_kats()
    Q = np.array(k["rate_matrix"])
    sr = k["site_rates"]
    q, cats = np.array(sr["grid"]), np.array(sr["rate_categories"])
    T = table_for(Q, q, cats)
    sym = T + np.swapaxes(T, 2, 3)
    priors = np.array([2 * math.log(r) - 3 * r for r in cats])
    for case in sr["cases"]:
        xa = np.array([c[0] for c in case["cherries"]])
        xb = np.array([c[1] for c in case["cherries"]])
        got = fo.site_rate_indices(xa, xb, sym, np.array(sr["lengths_index"])[: len(xa)], priors)
        assert got.tolist() == case["expected"]


# the big demo families take a couple of seconds each in numpy: keep three of them on the CPU suite
_CPU_CASES = [c for c in CASES if c["demo_family"] is None] + [c for c in CASES if c["demo_family"]][:3]


@pytest.mark.parametrize("case", _CPU_CASES, ids=[c["name"] for c in _CPU_CASES])
def test_oracle_matches_reference_program(case):
    alphabet, Q = parse_rate_matrix(case["rate_matrix_text"])
    names, seqs = parse_msa(msa_text(case))
    enc = fo.encode(seqs, alphabet)
    q = fo.quantization_points(0.03, 1.1, case["num_steps"])
    cats = fo.rate_categories(case["num_rate_categories"])
    cherries, lengths, rates, _, _ = fo.fast_cherries_oracle(enc, table_for(Q, q, cats), q, cats, case["seed"],
                                                             case["max_iters"])
    exp_cherries, exp_dist, exp_rates = expected_outputs(case)
    assert [(names[a], names[b]) for a, b in cherries] == exp_cherries
    assert ["%.17f" % x for x in lengths] == exp_dist
    assert ["%.17f" % x for x in rates] == exp_rates


def test_mt19937_known_answer():
    # the 10000th output of std::mt19937 seeded with 5489 is 4123659995 (C++ standard, [rand.predef])
    rng = fo.MT19937(5489)
    for _ in range(9999):
        rng()
    assert rng() == 4123659995


def test_product_setup_scalars_match_oracle():
    from cherryml_b200.phylogeny_estimation import _fast_cherries as fc

    for steps in (8, 64):
        assert np.array_equal(fc.quantization_grid(0.03, 1.1, steps), fo.quantization_points(0.03, 1.1, steps))
    for R in (1, 2, 4, 20):
        cats = fc.ble_rate_categories(R)
        assert np.array_equal(cats, fo.rate_categories(R))
        assert np.array_equal(fc.initial_rate_weights(cats), fo.initial_site_rate_weights(cats))
    # the gamma CDF itself, against scipy, in both branches of AS 32
    from scipy.special import gammainc

    for x in (0.01, 0.5, 0.99, 2.9, 3.0, 3.1, 10.0, 50.0):
        assert abs(fc._gamma_cdf(x, 3.0) - gammainc(3.0, x)) < 1e-7


def test_encode_families_and_tree_layout(tmp_path):
    from cherryml_b200.phylogeny_estimation import _fast_cherries as fc

    p = tmp_path / "fam.txt"
    p.write_text(">a\nAR-X\n>b\nNDCQ\n>c\nARND\n")
    names, buf, fams = fc.encode_families([str(p)], list("ARNDCQ"))
    assert names == [["a", "b", "c"]]
    assert fams["n_seqs"][0] == 3 and fams["n_sites"][0] == 4 and fams["row_stride"][0] == 16
    rows = buf.reshape(3, 16)
    assert rows[0, :4].tolist() == [0, 1, 6, 6] and rows[1, :4].tolist() == [2, 3, 4, 5]
    assert (rows[:, 4:] == 6).all()
    tree = fc.cherries_tree(["a", "b", "c"], [(2, 0)], [0.25], 1)
    assert tree.nodes() == ["root", "internal-0", "c", "a", "b"]
    assert tree.edges() == [("root", "internal-0", 1.0), ("internal-0", "c", 0.125), ("internal-0", "a", 0.125),
                            ("root", "b", 1.0)]
