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
