"""CPU ORACLE for transition counting -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import
this module.  The product (``cherryml_b200``) never does.

A numpy restatement of the reference's counting algorithm (songlab-cal/CherryML v0.2.0),
reading the reference's own text formats directly (its parsers are restated here on
purpose: the oracle shares no code with ``cherryml_b200``):

* ``quantization_idx``                cherryml/utils.py:35-56 == counting/_count_transitions.cpp:295-307
* LG ``cherry++`` / ``cherry`` / ``edge``   counting/_count_transitions.py:37-198 (== .cpp:316-522)
* co-transitions                      counting/_count_co_transitions.py:38-224 (== .cpp:306-549)
* C++ personality: branch lengths parsed by ``std::stof`` (float32), .cpp:247.

Pinned (tests/test_oracle_counting.py) against the reference's golden count matrices
``tests/counting_tests/test_input_data/{tiny,tiny_2,tiny_3,tiny_4}/count_*`` (copied to
``tests/golden/counting``) and, in the build container, against ``medium/`` goldens and the
reference binaries compiled into ``oracle/_ref``.
"""
import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np


# ------------------------------------------------------------------ text formats
def parse_tree(path: str, float32_lengths: bool):
    """Returns (children: {node: [(child, length)]}, root).  io/_tree.py:214-265."""
    with open(path) as f:
        lines = f.read().strip().split("\n")
    n = int(lines[0].split(" ")[0])
    children: Dict[str, List[Tuple[str, float]]] = {lines[i]: [] for i in range(1, n + 1)}
    m = int(lines[n + 1].split(" ")[0])
    has_parent = set()
    for i in range(n + 2, n + 2 + m):
        u, v, length = lines[i].split(" ")
        length = float(length)
        if float32_lengths:
            length = float(np.float32(length))  # std::stof then widened to double
        children[u].append((v, length))
        has_parent.add(v)
    roots = [u for u in children if u not in has_parent]
    assert len(roots) == 1, roots
    return children, roots[0]


def parse_msa(path: str) -> Dict[str, str]:
    with open(path) as f:
        lines = f.read().strip().split("\n")
    return {lines[2 * i][1:]: lines[2 * i + 1] for i in range(len(lines) // 2)}


def parse_site_rates(path: str) -> np.ndarray:
    lines = open(path).read().strip().split("\n")
    return np.array([float(x) for x in lines[1].split(" ")], dtype=np.float64)


def parse_contact_map(path: str) -> np.ndarray:
    lines = open(path).read().strip().split("\n")
    L = int(lines[0].split(" ")[0])
    return np.array([[int(c) for c in lines[i + 1]] for i in range(L)], dtype=np.int64)


# ------------------------------------------------------------------ quantisation
def quantization_idx(t: float, q: np.ndarray) -> Optional[int]:
    """Scalar definition, cherryml/utils.py:35-56."""
    if t < q[0] or t > q[-1]:
        return None
    ub = int(np.searchsorted(q, t))
    if ub == 0:
        return 0
    if t / q[ub - 1] - 1 < q[ub] / t - 1:
        return ub - 1
    return ub


def quantization_idx_vec(t: np.ndarray, q: np.ndarray) -> np.ndarray:
    """Vectorised, elementwise identical to the scalar definition; -1 = outside the grid."""
    t = np.asarray(t, dtype=np.float64)
    out = np.full(t.shape, -1, dtype=np.int64)
    ok = ~((t < q[0]) | (t > q[-1]))
    tt = t[ok]
    ub = np.searchsorted(q, tt)  # side="left": first index with q >= t
    res = ub.copy()
    nz = ub > 0
    left = q[ub[nz] - 1]
    right = q[ub[nz]]
    pick_left = (tt[nz] / left - 1) < (right / tt[nz] - 1)
    tmp = ub[nz]
    tmp[pick_left] -= 1
    res[nz] = tmp
    out[ok] = res
    return out


# ---------------------------------------------------------------------- pairing
def leaf_pairs(children, root, mode: str) -> List[Tuple[str, str, float]]:
    """(a, b, t) per counted pair.  cherry++: _count_transitions.py:65-126."""
    pairs: List[Tuple[str, str, float]] = []
    if mode == "cherry++":
        import sys

        sys.setrecursionlimit(max(sys.getrecursionlimit(), 100000))

        def dfs(node):
            if not children[node]:
                return node, 0.0
            leaves, dists = [], []
            for child, length in children[node]:
                leaf, d = dfs(child)
                if leaf is not None:
                    leaves.append(leaf)
                    dists.append(d + length)
            i = 0
            while i + 1 <= len(leaves) - 1:
                pairs.append((leaves[i], leaves[i + 1], dists[i] + dists[i + 1]))
                i += 2
            if len(leaves) % 2 == 0:
                return None, None
            return leaves[-1], dists[-1]

        dfs(root)
        n_leaves = sum(1 for v in children if not children[v])
        assert len(pairs) == n_leaves // 2
    elif mode == "cherry":
        for node, ch in children.items():
            if len(ch) == 2 and all(not children[c] for c, _ in ch):
                pairs.append((ch[0][0], ch[1][0], ch[0][1] + ch[1][1]))
    elif mode == "edge":
        for node, ch in children.items():
            for child, length in ch:
                pairs.append((node, child, length))
    else:
        raise ValueError(mode)
    return pairs


def _encode(seq: str, lut: np.ndarray) -> np.ndarray:
    return lut[np.frombuffer(seq.encode("latin-1"), dtype=np.uint8)]


def _lut(states: Sequence[str]) -> np.ndarray:
    lut = np.full(256, -1, dtype=np.int64)
    for i, s in enumerate(states):
        lut[ord(s)] = i
    return lut


# --------------------------------------------------------------------- counting
def count_transitions_oracle(
    tree_dir: str,
    msa_dir: str,
    site_rates_dir: str,
    families: Sequence[str],
    amino_acids: Sequence[str],
    quantization_points: Sequence[float],
    edge_or_cherry: str,
    float32_branch_lengths: bool,
) -> Tuple[np.ndarray, np.ndarray]:
    """Returns (sorted grid [K], counts fp64 [K,S,S])."""
    if edge_or_cherry.startswith("cherry++__"):
        edge_or_cherry = "cherry++"
    q = np.array(sorted(float(x) for x in quantization_points), dtype=np.float64)
    S = len(amino_acids)
    lut = _lut(amino_acids)
    counts = np.zeros((len(q), S, S), dtype=np.float64)
    for fam in families:
        children, root = parse_tree(os.path.join(tree_dir, fam + ".txt"), float32_branch_lengths)
        msa = parse_msa(os.path.join(msa_dir, fam + ".txt"))
        rates = parse_site_rates(os.path.join(site_rates_dir, fam + ".txt"))
        for a, b, t in leaf_pairs(children, root, edge_or_cherry):
            xa, xb = _encode(msa[a], lut), _encode(msa[b], lut)
            bucket = quantization_idx_vec(t * rates[: len(xa)], q)
            ok = (bucket >= 0) & (xa >= 0) & (xb >= 0)
            if edge_or_cherry == "edge":
                np.add.at(counts, (bucket[ok], xa[ok], xb[ok]), 1.0)
            else:
                np.add.at(counts, (bucket[ok], xa[ok], xb[ok]), 0.5)
                np.add.at(counts, (bucket[ok], xb[ok], xa[ok]), 0.5)
    return q, counts


def count_co_transitions_oracle(
    tree_dir: str,
    msa_dir: str,
    contact_map_dir: str,
    families: Sequence[str],
    amino_acids: Sequence[str],
    quantization_points: Sequence[float],
    edge_or_cherry: str,
    minimum_distance_for_nontrivial_contact: int,
    float32_branch_lengths: bool,
) -> Tuple[np.ndarray, np.ndarray]:
    """Returns (sorted grid [K], counts fp64 [K,S^2,S^2])."""
    if edge_or_cherry.startswith("cherry++__"):
        edge_or_cherry = "cherry++"
    q = np.array(sorted(float(x) for x in quantization_points), dtype=np.float64)
    S = len(amino_acids)
    lut = _lut(amino_acids)
    counts = np.zeros((len(q), S * S, S * S), dtype=np.float64)
    for fam in families:
        children, root = parse_tree(os.path.join(tree_dir, fam + ".txt"), float32_branch_lengths)
        msa = parse_msa(os.path.join(msa_dir, fam + ".txt"))
        cmap = parse_contact_map(os.path.join(contact_map_dir, fam + ".txt"))
        ii, jj = np.where(cmap == 1)
        keep = (np.abs(ii - jj) >= minimum_distance_for_nontrivial_contact) & (ii < jj)
        ii, jj = ii[keep], jj[keep]
        for a, b, t in leaf_pairs(children, root, edge_or_cherry):
            k = quantization_idx(t, q)
            if k is None:
                continue
            xa, xb = _encode(msa[a], lut), _encode(msa[b], lut)
            ok = (xa[ii] >= 0) & (xa[jj] >= 0) & (xb[ii] >= 0) & (xb[jj] >= 0)
            s = xa[ii[ok]] * S + xa[jj[ok]]
            e = xb[ii[ok]] * S + xb[jj[ok]]
            sr = xa[jj[ok]] * S + xa[ii[ok]]
            er = xb[jj[ok]] * S + xb[ii[ok]]
            if edge_or_cherry == "edge":
                np.add.at(counts[k], (s, e), 0.5)
                np.add.at(counts[k], (sr, er), 0.5)
            else:
                for u, v in ((s, e), (sr, er), (e, s), (er, sr)):
                    np.add.at(counts[k], (u, v), 0.25)
    return q, counts


def read_count_matrices_text(path: str) -> Tuple[np.ndarray, List[str], np.ndarray]:
    """Minimal result.txt reader (io/_count_matrices.py:8-63) for comparing with goldens."""
    lines = open(path).read().strip().split("\n")
    K = int(lines[0].split(" ")[0])
    S = int(lines[1].split(" ")[0])
    q = np.zeros(K)
    counts = np.zeros((K, S, S))
    pos, states = 2, []
    for k in range(K):
        q[k] = float(lines[pos])
        states = lines[pos + 1].strip().split()
        for i in range(S):
            counts[k, i] = [float(x) for x in lines[pos + 2 + i].strip().split()[1:]]
        pos += 2 + S
    return q, states, counts
