"""Host-side ingest (pairing, encoding, column sorting, tiles) checked on the CPU.

tests/_emulate.py replays what the CUDA kernels do with an encoded batch in numpy; the
result must equal the oracle, so a GPU-side mismatch can only come from the kernels."""
import os
import tempfile

import numpy as np
import pytest

from cherryml_b200.counting._ingest import (build_co_batch, build_lg_batch, contacting_pairs,
                                             extract_pairs, lg_column_layout)
from cherryml_b200.io import read_tree
from cherryml_b200.synthetic import (as_count_batch, quantization_grid, synthetic_co, synthetic_lg,
                                     write_text_rendering)
from cherryml_b200.utils import amino_acids
from oracle.counting_oracle import count_co_transitions_oracle, count_transitions_oracle, leaf_pairs, parse_tree
from oracle.native import count_batch_oracle
from tests._emulate import emulate_count, emulate_symmetrize
from tests.test_oracle_counting import CO_CASES, GRID_CO, GRID_LG, LG_CASES, MEDIUM3, MODES


@pytest.mark.parametrize("case", LG_CASES, ids=lambda c: f"{c[0]}-{c[4]}")
def test_lg_ingest_tiny(golden_counting, case):
    ds, fams, aa, grid, mode, _ = case
    root = os.path.join(golden_counting, ds)
    for f32 in (True, False):
        batch = build_lg_batch(f"{root}/tree_dir", f"{root}/msa_dir", f"{root}/site_rates_dir", fams, aa, mode, f32)
        got = emulate_symmetrize(emulate_count(batch, sorted(grid), len(aa)), "lg", len(aa), mode == "edge")
        _, exp = count_transitions_oracle(f"{root}/tree_dir", f"{root}/msa_dir", f"{root}/site_rates_dir",
                                          fams, aa, grid, mode, f32)
        assert np.array_equal(got, exp)


@pytest.mark.parametrize("case", CO_CASES, ids=lambda c: f"{c[0]}-{c[4]}")
def test_co_ingest_tiny(golden_counting, case):
    ds, fams, aa, grid, mode, _ = case
    root = os.path.join(golden_counting, ds)
    batch = build_co_batch(f"{root}/tree_dir", f"{root}/msa_dir", f"{root}/contact_map_dir", fams, aa, mode, 2, True)
    got = emulate_symmetrize(emulate_count(batch, sorted(grid), len(aa)), "co", len(aa), mode == "edge")
    _, exp = count_co_transitions_oracle(f"{root}/tree_dir", f"{root}/msa_dir", f"{root}/contact_map_dir",
                                         fams, aa, grid, mode, 2, True)
    assert np.array_equal(got, exp)


@pytest.mark.parametrize("mode,tag,msa_sub", MODES)
def test_ingest_medium3(golden_counting, mode, tag, msa_sub):
    m3 = os.path.join(golden_counting, "medium3")
    batch = build_lg_batch(f"{m3}/tree_dir", f"{m3}/{msa_sub}", f"{m3}/site_rates_dir", MEDIUM3, amino_acids, mode, True)
    got = emulate_symmetrize(emulate_count(batch, sorted(GRID_LG), 20), "lg", 20, mode == "edge")
    assert np.array_equal(got, count_batch_oracle(batch, GRID_LG, 20, mode == "edge"))
    cb = build_co_batch(f"{m3}/tree_dir", f"{m3}/{msa_sub}", f"{m3}/contact_map_dir", MEDIUM3[:2], amino_acids, mode, 7, True)
    got = emulate_symmetrize(emulate_count(cb, sorted(GRID_CO), 20), "co", 20, mode == "edge")
    assert np.array_equal(got, count_batch_oracle(cb, GRID_CO, 20, mode == "edge"))


def test_pairing_matches_oracle_pairing(golden_counting):
    m3 = os.path.join(golden_counting, "medium3")
    for fam in MEDIUM3:
        for mode in ("cherry++", "cherry", "edge"):
            for f32 in (True, False):
                mine = extract_pairs(read_tree(f"{m3}/tree_dir/{fam}.txt"), mode, f32)
                children, root = parse_tree(f"{m3}/tree_dir/{fam}.txt", f32)
                assert mine == leaf_pairs(children, root, mode)


def test_column_layout_properties():
    rng = np.random.default_rng(3)
    for L in (1, 3, 4, 17, 210, 300, 511):
        rates = rng.choice([0.1, 0.5, 1.0, 2.5, 7.0], size=L)
        vals, dest, group_cat, stride = lg_column_layout(rates)
        assert stride % 16 == 0 and len(group_cat) == stride // 4
        assert len(set(dest.tolist())) == L and dest.max() < stride
        # every site lands in a group of its own category
        assert np.array_equal(vals[group_cat[dest // 4]], rates)
    # continuous rates: every site its own category, still valid
    rates = rng.random(37)
    vals, dest, group_cat, stride = lg_column_layout(rates)
    assert len(vals) == 37 and np.array_equal(vals[group_cat[dest // 4]], rates)


def test_contacting_pairs():
    cmap = np.zeros((12, 12), dtype=int)
    for i, j in [(0, 7), (7, 0), (1, 3), (2, 11), (5, 5), (4, 10)]:
        cmap[i, j] = 1
    assert contacting_pairs(cmap, 7).tolist() == [[0, 7], [2, 11]]
    assert contacting_pairs(cmap, 2).tolist() == [[0, 7], [1, 3], [2, 11], [4, 10]]


def test_empty_and_ragged_inputs(tmp_path):
    """Families with a single leaf (no pairs), an all-gap sequence and an empty family list."""
    d = tmp_path
    for sub in ("tree_dir", "msa_dir", "site_rates_dir"):
        os.makedirs(d / sub)
    (d / "tree_dir" / "one.txt").write_text("2 nodes\nr\na\n1 edges\nr a 0.5\n")
    (d / "msa_dir" / "one.txt").write_text(">a\nARND\n")
    (d / "site_rates_dir" / "one.txt").write_text("4 sites\n1.0 1.0 2.0 2.0")
    (d / "tree_dir" / "two.txt").write_text("3 nodes\nr\na\nb\n2 edges\nr a 0.5\nr b 0.25\n")
    (d / "msa_dir" / "two.txt").write_text(">a\n----A\n>b\nAR-XA\n")
    (d / "site_rates_dir" / "two.txt").write_text("5 sites\n1.0 0.5 1.0 3.0 1.0")
    grid = [0.1, 0.75, 2.0]
    args = (str(d / "tree_dir"), str(d / "msa_dir"), str(d / "site_rates_dir"))
    for fams in ([], ["one"], ["one", "two"]):
        batch = build_lg_batch(*args, fams, amino_acids, "cherry++", False)
        got = emulate_symmetrize(emulate_count(batch, grid, 20), "lg", 20, False)
        _, exp = count_transitions_oracle(*args, fams, amino_acids, grid, "cherry++", False)
        assert np.array_equal(got, exp)
    assert exp.sum() == 1.0  # only the last site of "two" is a valid pair


def test_synthetic_batches_and_text_rendering():
    grid = quantization_grid()
    syn = synthetic_lg(5, 48, 123, 4, seed=7)
    batch = as_count_batch(syn)
    c = count_batch_oracle(batch, grid, 20, False)
    got = emulate_symmetrize(emulate_count(batch, grid, 20), "lg", 20, False)
    assert np.array_equal(got, c)
    with tempfile.TemporaryDirectory() as d:
        names = write_text_rendering(syn, d)
        _, oc = count_transitions_oracle(d + "/tree_dir", d + "/msa_dir", d + "/site_rates_dir", names,
                                         amino_acids, grid, "cherry++", False)
        assert np.array_equal(oc, c)
        again = build_lg_batch(d + "/tree_dir", d + "/msa_dir", d + "/site_rates_dir", names, amino_acids,
                               "cherry++", False)
        assert np.array_equal(again.msa, batch.msa) and np.array_equal(again.pair_t, batch.pair_t)
    syn = synthetic_co(3, 24, 90, seed=8)
    batch = as_count_batch(syn)
    c = count_batch_oracle(batch, grid, 20, False)
    got = emulate_symmetrize(emulate_count(batch, grid, 20), "co", 20, False)
    assert np.array_equal(got, c)
    with tempfile.TemporaryDirectory() as d:
        names = write_text_rendering(syn, d)
        _, oc = count_co_transitions_oracle(d + "/tree_dir", d + "/msa_dir", d + "/contact_map_dir", names,
                                            amino_acids, grid, "cherry++", 7, False)
        assert np.array_equal(oc, c)


# ------------------------------------------------------------ native (C++) ingest == Python ingest
_BATCH_FIELDS = ("msa", "fams", "pair_a", "pair_b", "pair_t", "pair_fam", "rate_vals", "aux", "tiles")


def _assert_same_batch(a, b):
    for name in _BATCH_FIELDS:
        x, y = getattr(a, name), getattr(b, name)
        assert x.dtype == y.dtype and x.shape == y.shape, name
        assert np.array_equal(x, y), name
    assert a.r_pad == b.r_pad and a.n_sites_examined == b.n_sites_examined
    assert a.family_names == b.family_names


@pytest.mark.parametrize("case", LG_CASES + CO_CASES, ids=lambda c: f"{c[0]}-{c[4]}-{c[5][:9]}")
def test_native_ingest_equals_python_ingest_tiny(golden_counting, case):
    from cherryml_b200.counting._ingest import build_co_batch_native, build_lg_batch_native

    ds, fams, aa, _, mode, gdir = case
    root = os.path.join(golden_counting, ds)
    for f32 in (True, False):
        for nt in (1, 3):
            if "co_matrices" in gdir:
                py = build_co_batch(f"{root}/tree_dir", f"{root}/msa_dir", f"{root}/contact_map_dir", fams, aa, mode, 2, f32)
                nat = build_co_batch_native(f"{root}/tree_dir", f"{root}/msa_dir", f"{root}/contact_map_dir", fams, aa,
                                            mode, 2, f32, n_threads=nt)
            else:
                py = build_lg_batch(f"{root}/tree_dir", f"{root}/msa_dir", f"{root}/site_rates_dir", fams, aa, mode, f32)
                nat = build_lg_batch_native(f"{root}/tree_dir", f"{root}/msa_dir", f"{root}/site_rates_dir", fams, aa,
                                            mode, f32, n_threads=nt)
            _assert_same_batch(py, nat)


@pytest.mark.parametrize("mode,tag,msa_sub", MODES)
def test_native_ingest_equals_python_ingest_medium3(golden_counting, mode, tag, msa_sub):
    from cherryml_b200.counting._ingest import build_co_batch_native, build_lg_batch_native

    m3 = os.path.join(golden_counting, "medium3")
    for f32 in (True, False):
        _assert_same_batch(
            build_lg_batch(f"{m3}/tree_dir", f"{m3}/{msa_sub}", f"{m3}/site_rates_dir", MEDIUM3, amino_acids, mode, f32),
            build_lg_batch_native(f"{m3}/tree_dir", f"{m3}/{msa_sub}", f"{m3}/site_rates_dir", MEDIUM3, amino_acids,
                                  mode, f32))
        _assert_same_batch(
            build_co_batch(f"{m3}/tree_dir", f"{m3}/{msa_sub}", f"{m3}/contact_map_dir", MEDIUM3, amino_acids, mode, 7, f32),
            build_co_batch_native(f"{m3}/tree_dir", f"{m3}/{msa_sub}", f"{m3}/contact_map_dir", MEDIUM3, amino_acids,
                                  mode, 7, f32))


def test_native_ingest_synthetic_rendering_and_empty_family_list(tmp_path):
    from cherryml_b200.counting._ingest import build_co_batch_native, build_lg_batch_native

    syn = synthetic_lg(5, 16, 70, 4, seed=3)
    names = write_text_rendering(syn, str(tmp_path / "lg"))
    d = str(tmp_path / "lg")
    _assert_same_batch(
        build_lg_batch(d + "/tree_dir", d + "/msa_dir", d + "/site_rates_dir", names, amino_acids, "cherry++", False),
        build_lg_batch_native(d + "/tree_dir", d + "/msa_dir", d + "/site_rates_dir", names, amino_acids, "cherry++",
                              False, n_threads=4))
    syn = synthetic_co(4, 12, 60, seed=5)
    names = write_text_rendering(syn, str(tmp_path / "co"))
    d = str(tmp_path / "co")
    _assert_same_batch(
        build_co_batch(d + "/tree_dir", d + "/msa_dir", d + "/contact_map_dir", names, amino_acids, "cherry++", 7, True),
        build_co_batch_native(d + "/tree_dir", d + "/msa_dir", d + "/contact_map_dir", names, amino_acids, "cherry++",
                              7, True, n_threads=2))
    _assert_same_batch(
        build_lg_batch(d + "/tree_dir", d + "/msa_dir", d + "/site_rates_dir", [], amino_acids, "cherry++", True),
        build_lg_batch_native(d + "/tree_dir", d + "/msa_dir", d + "/site_rates_dir", [], amino_acids, "cherry++", True))


def test_native_ingest_errors(tmp_path):
    """Malformed inputs are rejected with the readers' messages (no partial batch)."""
    from cherryml_b200 import _lib
    from cherryml_b200.counting._ingest import build_lg_batch_native

    syn = synthetic_lg(2, 8, 20, 2, seed=1)
    d = str(tmp_path)
    names = write_text_rendering(syn, d)
    args = (d + "/tree_dir", d + "/msa_dir", d + "/site_rates_dir")
    with pytest.raises(_lib.CherryError, match="cannot open"):
        build_lg_batch_native(*args, names + ["missing"], amino_acids, "cherry++", True)
    with pytest.raises(_lib.CherryError, match="Unknown edge_or_cherry"):
        build_lg_batch_native(*args, names, amino_acids, "twig", True)
    tree = open(f"{d}/tree_dir/{names[0]}.txt").read()
    open(f"{d}/tree_dir/{names[0]}.txt", "w").write(tree.replace(" nodes", " knots", 1))
    with pytest.raises(_lib.CherryError, match="should start with"):
        build_lg_batch_native(*args, names, amino_acids, "cherry++", True)
    open(f"{d}/tree_dir/{names[0]}.txt", "w").write(tree)
    msa = open(f"{d}/msa_dir/{names[1]}.txt").read()
    open(f"{d}/msa_dir/{names[1]}.txt", "w").write(msa.replace(">seq3\n", ">other\n", 1))
    with pytest.raises(_lib.CherryError, match="not in the MSA"):
        build_lg_batch_native(*args, names, amino_acids, "cherry++", True)
    open(f"{d}/msa_dir/{names[1]}.txt", "w").write(msa)
    rates = open(f"{d}/site_rates_dir/{names[0]}.txt").read()
    open(f"{d}/site_rates_dir/{names[0]}.txt", "w").write("5 sites\n" + " ".join(rates.split("\n")[1].split(" ")[:5]))
    with pytest.raises(_lib.CherryError, match="only 5 site rates"):
        build_lg_batch_native(*args, names, amino_acids, "cherry++", True)
