"""Host-side set-up code that was rewritten for speed must keep producing exactly what the
straightforward statement of the same computation produces (CPU only)."""
import numpy as np
import torch

from cherryml_b200.evaluation._likelihood import _encode_leaves, _tree_arrays
from cherryml_b200.io import Tree
from cherryml_b200.siterm._vectorized import solve_stationary_dist_fast
from cherryml_b200.utils import amino_acids


def _plain_power_iteration(rate_matrices):
    """The reference's loop as written (_cherryml_vectorized.py:70-104): every matrix, 100 squarings."""
    diag_avg = np.mean(np.diagonal(rate_matrices, axis1=1, axis2=2), axis=1)
    normalized = rate_matrices * (-1.0 / diag_avg)[:, None, None]
    e = torch.matrix_exp(torch.tensor(normalized, dtype=torch.float32)).numpy()
    for _ in range(100):
        e = e @ e
        e /= e.sum(axis=2, keepdims=True)
    pi = e[:, 0, :]
    pi /= pi.sum(axis=1, keepdims=True)
    return pi


def _random_rate_matrices(rng, n, S):
    m = rng.uniform(0.01, 1.0, (n, S, S))
    for i in range(n):
        np.fill_diagonal(m[i], 0.0)
        np.fill_diagonal(m[i], -m[i].sum(axis=1))
    return m


def test_stationary_distributions_bit_identical_to_the_plain_loop():
    rng = np.random.default_rng(3)
    base = _random_rate_matrices(rng, 1, 20)[0]
    cases = [
        base[None] * rng.choice(np.linspace(0.05, 4.0, 20), 150)[:, None, None],  # a family: rate_l * Q0
        base[None] * rng.uniform(0.01, 5.0, 40)[:, None, None],
        _random_rate_matrices(rng, 30, 20),
        _random_rate_matrices(rng, 12, 4),
        _random_rate_matrices(rng, 1, 20),
    ]
    for m in cases:
        want = _plain_power_iteration(m.copy())
        got = solve_stationary_dist_fast(m.copy())
        assert got.dtype == want.dtype and got.shape == want.shape
        assert got.tobytes() == want.tobytes()


def _random_tree(rng, n_leaves, max_children=2):
    tree = Tree()
    roots = []
    for i in range(n_leaves):
        tree.add_node(f"l{i}")
        roots.append(f"l{i}")
    edges, k = [], 0
    while len(roots) > 1:
        take = min(len(roots), int(rng.integers(2, max_children + 1)))
        idx = sorted(rng.choice(len(roots), take, replace=False).tolist(), reverse=True)
        kids = [roots.pop(i) for i in idx]
        parent = f"i{k}"
        k += 1
        tree.add_node(parent)
        edges.extend((parent, c, float(rng.uniform(0.0, 1.0))) for c in kids)
        roots.append(parent)
    for e in reversed(edges):
        tree.add_edge(*e)
    return tree


def test_node_table_follows_the_post_order_traversal():
    rng = np.random.default_rng(5)
    trees = [_random_tree(rng, n, mc) for n, mc in ((1, 2), (2, 2), (9, 2), (40, 4), (300, 3))]
    for tree in trees:
        nodes, lengths, leaves, max_depth = _tree_arrays(tree)
        order = tree.postorder_traversal()
        assert len(nodes) == len(order) == tree.num_nodes()
        depth = {tree.root(): 0}
        for v in tree.preorder_traversal():
            for c, _ in tree.children(v):
                depth[c] = depth[v] + 1
        assert max_depth == max(depth.values())
        assert leaves == [v for v in order if tree.is_leaf(v)]
        for i, v in enumerate(order):
            assert nodes["depth"][i] == depth[v]
            assert bool(nodes["flags"][i] & 1) == tree.is_leaf(v)
            if tree.is_leaf(v):
                assert leaves[nodes["obs_row"][i]] == v
            else:
                assert nodes["obs_row"][i] == -1
            if tree.is_root(v):
                assert lengths[i] == 0.0 and not nodes["flags"][i] & 2
            else:
                parent, length = tree.parent(v)
                assert lengths[i] == length
                assert bool(nodes["flags"][i] & 2) == (tree.children(parent)[0][0] == v)


def test_leaf_encoding():
    import pytest

    msa = {"a": "ACDXY-", "b": "acd.WV", "c": "YYYYYY"}
    enc = _encode_leaves(msa, ["c", "a", "b"], amino_acids)
    assert enc.dtype == np.uint8 and enc.shape == (3, 6)
    aa = {ch: i for i, ch in enumerate(amino_acids)}
    for row, name in zip(enc, ["c", "a", "b"]):
        assert row.tolist() == [aa.get(ch, 20) for ch in msa[name]]
    with pytest.raises(ValueError, match="different lengths"):
        _encode_leaves({"a": "AC", "b": "A"}, ["a", "b"], amino_acids)


def test_likelihood_stage_files_with_the_device_function_replaced(tmp_path, monkeypatch):
    """The stage's own work (validation, per-family files, caching layout) around a stand-in for
    dp_likelihood_computation."""
    import os

    import pytest

    from cherryml_b200 import io
    from cherryml_b200.evaluation import _likelihood as ev
    from cherryml_b200.markov_chain import compute_stationary_distribution, get_lg_path

    families = ["f1", "f0"]
    for f in families:
        tree = io.Tree()
        tree.add_nodes(["r", "a", "b"])
        tree.add_edge("r", "a", 0.1)
        tree.add_edge("r", "b", 0.2)
        io.write_tree(tree, str(tmp_path / "tree" / (f + ".txt")))
        io.write_msa({"a": "ACD", "b": "ACE"}, str(tmp_path / "msa" / (f + ".txt")))
        io.write_site_rates([1.0, 0.5, 2.0], str(tmp_path / "rates" / (f + ".txt")))
    Q = io.read_rate_matrix(get_lg_path()).to_numpy(dtype=np.float64)
    io.write_probability_distribution(compute_stationary_distribution(Q), amino_acids, str(tmp_path / "pi.txt"))
    calls = []

    def fake_dp(tree, msa, contact_map, site_rates, output_profiling_path=None, **kw):
        calls.append((sorted(msa), contact_map, list(site_rates), kw["device_1"]))
        return -6.0, [-1.0, -2.0, -3.0]

    monkeypatch.setattr(ev, "dp_likelihood_computation", fake_dp)
    kw = dict(tree_dir=str(tmp_path / "tree"), msa_dir=str(tmp_path / "msa"), site_rates_dir=str(tmp_path / "rates"),
              contact_map_dir=None, families=families, amino_acids=amino_acids, pi_1_path=str(tmp_path / "pi.txt"),
              Q_1_path=get_lg_path(), reversible_1=True, device_1="cuda", pi_2_path=None, Q_2_path=None,
              reversible_2=None, device_2=None, num_processes=1)
    out = str(tmp_path / "ll")
    ev.compute_log_likelihoods(output_likelihood_dir=out, **kw)
    assert len(calls) == 2 and calls[0] == (["a", "b"], None, [1.0, 0.5, 2.0], "cuda")
    for f in families:
        assert io.read_log_likelihood(os.path.join(out, f + ".txt")) == (-6.0, [-1.0, -2.0, -3.0])
    assert os.path.exists(os.path.join(out, "profiling.txt"))
    # a model whose states are not the requested alphabet is rejected before any family is touched
    with pytest.raises(Exception, match="expected amino acids"):
        ev.compute_log_likelihoods(output_likelihood_dir=str(tmp_path / "ll2"), **{**kw, "amino_acids": amino_acids[::-1]})
    with pytest.raises(NotImplementedError):
        ev.compute_log_likelihoods(output_likelihood_dir=str(tmp_path / "ll3"), use_cpp_implementation=True, **kw)
