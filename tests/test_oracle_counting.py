"""The counting oracle is pinned here (CPU, no GPU needed).

1. against the reference's own golden count matrices for its tiny fixtures (every mode the
   reference tests: tests/counting_tests/counting_test.py:91-413);
2. against goldens produced by RUNNING the unmodified reference (C++ binaries and Python
   implementation) on three families of its ``medium`` fixture (tests/golden/make_golden.py);
3. in the build container only: against the reference's ``medium`` goldens on all 32 families
   (counting_test.py:416-590) read straight from the reference checkout.
The C restatement on encoded arrays is pinned to the same goldens.
"""
import os

import numpy as np
import pytest

from cherryml_b200.counting._ingest import build_co_batch, build_lg_batch
from cherryml_b200.utils import amino_acids
from oracle.counting_oracle import (
    count_co_transitions_oracle,
    count_transitions_oracle,
    quantization_idx,
    quantization_idx_vec,
    read_count_matrices_text,
)
from oracle.native import count_batch_oracle, quantization_idx_c

FAMS3 = ["fam1", "fam2", "fam3"]
ILST = ["I", "L", "S", "T"]
GRID7 = [1.0, 2.0, 3.0, 4.0, 5.0, 6.0, 7.0]
GRID_LG = [0.06 * 1.1**i for i in range(-51, 51, 1)]
GRID_CO = [0.06 * 2.0**i for i in range(-5, 5, 1)]
MEDIUM3 = ["1a92_1_A", "1a4p_1_A", "1a64_1_A"]

# (dataset, families, alphabet, grid, mode, golden dir) -- the reference's tiny LG cases
LG_CASES = [
    ("tiny", FAMS3, ILST, [1.99, 5.01], "edge", "count_matrices_dir_edges"),
    ("tiny", FAMS3, ILST, [1.99, 10.01], "cherry", "count_matrices_dir_cherries"),
    ("tiny", FAMS3, ILST, [1.99, 10.01], "cherry++", "count_matrices_dir_cherries"),
    ("tiny_2", FAMS3, ILST, [1.99, 10.01], "cherry++", "count_matrices_dir_cherries_plus_plus"),
    ("tiny_3", ["fam1"], list("ABCDEF"), GRID7, "cherry++", "count_matrices_dir_cherries_plus_plus"),
]
CO_CASES = [
    ("tiny", FAMS3, ILST, [1.99, 5.01], "edge", "count_co_matrices_dir_edges"),
    ("tiny", FAMS3, ILST, [1.99, 10.01], "cherry", "count_co_matrices_dir_cherries"),
    ("tiny", FAMS3, ILST, [1.99, 10.01], "cherry++", "count_co_matrices_dir_cherries"),
    ("tiny_2", FAMS3, ILST, [1.99, 10.01], "cherry++", "count_co_matrices_dir_cherries_plus_plus"),
    ("tiny_4", ["fam1"], list("ABC"), GRID7, "cherry++", "count_co_matrices_dir_cherries_plus_plus"),
]


def _golden(root, ds, d):
    return read_count_matrices_text(os.path.join(root, ds, d, "result.txt"))


@pytest.mark.parametrize("case", LG_CASES, ids=lambda c: f"{c[0]}-{c[4]}")
@pytest.mark.parametrize("f32", [True, False])
def test_lg_oracle_matches_reference_goldens(golden_counting, case, f32):
    ds, fams, aa, grid, mode, gdir = case
    root = os.path.join(golden_counting, ds)
    q, counts = count_transitions_oracle(
        f"{root}/tree_dir", f"{root}/msa_dir", f"{root}/site_rates_dir", fams, aa, grid, mode, f32)
    gq, _, gcounts = _golden(golden_counting, ds, gdir)
    assert np.allclose(q, gq)
    assert np.array_equal(counts, gcounts)
    # C restatement on the encoded layout
    batch = build_lg_batch(f"{root}/tree_dir", f"{root}/msa_dir", f"{root}/site_rates_dir", fams, aa, mode, f32)
    assert np.array_equal(count_batch_oracle(batch, grid, len(aa), mode == "edge"), gcounts)


@pytest.mark.parametrize("case", CO_CASES, ids=lambda c: f"{c[0]}-{c[4]}")
def test_co_oracle_matches_reference_goldens(golden_counting, case):
    ds, fams, aa, grid, mode, gdir = case
    root = os.path.join(golden_counting, ds)
    q, counts = count_co_transitions_oracle(
        f"{root}/tree_dir", f"{root}/msa_dir", f"{root}/contact_map_dir", fams, aa, grid, mode, 2, True)
    gq, _, gcounts = _golden(golden_counting, ds, gdir)
    assert np.allclose(q, gq)
    assert np.array_equal(counts, gcounts)
    batch = build_co_batch(f"{root}/tree_dir", f"{root}/msa_dir", f"{root}/contact_map_dir", fams, aa, mode, 2, True)
    assert np.array_equal(count_batch_oracle(batch, grid, len(aa), mode == "edge"), gcounts)


MODES = [("cherry++", "cherries_plus_plus", "msa_dir"), ("cherry", "cherries", "msa_dir"),
         ("edge", "edges", "msa_with_anc_dir")]


@pytest.mark.parametrize("mode,tag,msa_sub", MODES)
@pytest.mark.parametrize("personality", ["cpp", "py"])
def test_lg_oracle_matches_reference_run_medium3(golden_counting, mode, tag, msa_sub, personality):
    m3 = os.path.join(golden_counting, "medium3")
    f32 = personality == "cpp"
    _, _, gcounts = read_count_matrices_text(f"{m3}/ref{personality}_count_matrices_dir_{tag}/result.txt")
    _, counts = count_transitions_oracle(
        f"{m3}/tree_dir", f"{m3}/{msa_sub}", f"{m3}/site_rates_dir", MEDIUM3, amino_acids, GRID_LG, mode, f32)
    assert np.array_equal(counts, gcounts)
    batch = build_lg_batch(f"{m3}/tree_dir", f"{m3}/{msa_sub}", f"{m3}/site_rates_dir", MEDIUM3, amino_acids, mode, f32)
    assert np.array_equal(count_batch_oracle(batch, GRID_LG, 20, mode == "edge"), gcounts)


@pytest.mark.parametrize("mode,tag,msa_sub", MODES)
def test_co_oracle_matches_reference_run_medium3(golden_counting, mode, tag, msa_sub):
    m3 = os.path.join(golden_counting, "medium3")
    gcounts = np.load(f"{m3}/refcpp_count_co_matrices_dir_{tag}/result.npz")["counts"]
    _, counts = count_co_transitions_oracle(
        f"{m3}/tree_dir", f"{m3}/{msa_sub}", f"{m3}/contact_map_dir", MEDIUM3, amino_acids, GRID_CO, mode, 7, True)
    assert np.array_equal(counts, gcounts)
    batch = build_co_batch(f"{m3}/tree_dir", f"{m3}/{msa_sub}", f"{m3}/contact_map_dir", MEDIUM3, amino_acids, mode, 7, True)
    assert np.array_equal(count_batch_oracle(batch, GRID_CO, 20, mode == "edge"), gcounts)
    if mode == "cherry++":
        py = np.load(f"{m3}/refpy_count_co_matrices_dir_cherries_plus_plus/result.npz")["counts"]
        _, counts = count_co_transitions_oracle(
            f"{m3}/tree_dir", f"{m3}/msa_dir", f"{m3}/contact_map_dir", MEDIUM3, amino_acids, GRID_CO, mode, 7, False)
        assert np.array_equal(counts, py)


def test_lg_oracle_matches_reference_medium_all32(reference_dir):
    """Build container only: the reference's own medium goldens, all 32 families."""
    d = os.path.join(reference_dir, "tests/counting_tests/test_input_data/medium")
    fams = sorted(x[:-4] for x in os.listdir(f"{d}/msa_dir"))
    _, counts = count_transitions_oracle(
        f"{d}/tree_dir", f"{d}/msa_dir", f"{d}/site_rates_dir", fams, amino_acids, GRID_LG, "cherry++", True)
    _, _, g = read_count_matrices_text(f"{d}/count_matrices_dir_cherries_plus_plus/result.txt")
    assert np.array_equal(counts, g)
    batch = build_lg_batch(f"{d}/tree_dir", f"{d}/msa_dir", f"{d}/site_rates_dir", fams, amino_acids, "cherry++", True)
    assert np.array_equal(count_batch_oracle(batch, GRID_LG, 20, False), g)


def test_quantization_idx_definitions_agree():
    rng = np.random.default_rng(0)
    grid = np.array(sorted(GRID_LG))
    # random points, exact grid points, midpoints, and values just around grid points
    t = np.concatenate([
        np.exp(rng.uniform(np.log(grid[0] / 3), np.log(grid[-1] * 3), 20000)),
        grid, np.sqrt(grid[:-1] * grid[1:]), np.nextafter(grid, 0), np.nextafter(grid, np.inf),
        0.5 * (grid[:-1] + grid[1:]),
    ])
    vec = quantization_idx_vec(t, grid)
    from cherryml_b200.utils import quantization_idx as product_qidx

    for ti, vi in zip(t, vec):
        ref = quantization_idx(float(ti), grid)
        assert (-1 if ref is None else ref) == vi
        assert quantization_idx_c(float(ti), grid) == vi
        p = product_qidx(float(ti), grid)
        assert (-1 if p is None else p) == vi
    assert (vec == -1).any() and (vec == 0).any() and (vec == len(grid) - 1).any()
