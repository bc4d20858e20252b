"""Text formats either side of the hot path (host side, Python).

These mirror the on-disk formats of the reference's ``cherryml/io`` package so that the
directories a CherryML user already has can be consumed and produced unchanged:

* MSA ``family.txt``          -- ``io/_msa.py:51-73``
* tree ("N nodes / M edges")  -- ``io/_tree.py:214-265`` (children keep edge-line order)
* site rates                  -- ``io/_site_rates.py:5-26``
* contact map                 -- ``io/_contact_map.py:6-28``
* count matrices result.txt   -- ``io/_count_matrices.py:8-81`` and the C++ writer
                                  ``counting/_count_transitions.cpp:524-548``
* rate / mask matrices        -- ``io/_rate_matrix.py:37-76``
"""
import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np


# --------------------------------------------------------------------------- trees
class Tree:
    """Rooted tree with ordered children (order = order of the edge lines)."""

    def __init__(self) -> None:
        self._children: Dict[str, List[Tuple[str, float]]] = {}
        self._parent: Dict[str, Tuple[str, float]] = {}
        self._edges: List[Tuple[str, str, float]] = []

    def add_node(self, v: str) -> None:
        self._children[v] = []

    def add_nodes(self, nodes: Sequence[str]) -> None:
        for v in nodes:
            self.add_node(v)

    def add_edge(self, u: str, v: str, length: float) -> None:
        if v in self._parent:
            raise Exception(
                f"Node {v} already has a parent ({self._parent[v][0]}), cannot "
                f"also have parent {u} - graph is not a tree."
            )
        self._children[u].append((v, length))
        self._parent[v] = (u, length)
        self._edges.append((u, v, length))

    def add_edges(self, edges) -> None:
        for u, v, length in edges:
            self.add_edge(u, v, length)

    def edges(self) -> List[Tuple[str, str, float]]:
        return self._edges[:]

    def is_node(self, v: str) -> bool:
        return v in self._children

    def nodes(self) -> List[str]:
        return list(self._children.keys())

    def root(self) -> str:
        roots = [u for u in self._children if u not in self._parent]
        if len(roots) != 1:
            raise Exception(f"Tree should have one root, but found: {roots}")
        return roots[0]

    def children(self, u: str) -> List[Tuple[str, float]]:
        return list(self._children[u])

    def is_leaf(self, u: str) -> bool:
        return len(self._children[u]) == 0

    def is_root(self, u: str) -> bool:
        return u not in self._parent

    def num_nodes(self) -> int:
        return len(self._children)

    def num_edges(self) -> int:
        return len(self._edges)

    def parent(self, u: str) -> Tuple[str, float]:
        return self._parent[u]

    def leaves(self) -> List[str]:
        return [u for u in self._children if not self._children[u]]

    def internal_nodes(self) -> List[str]:
        return [u for u in self._children if self._children[u]]

    def scaled(self, scaling_factor: float, node_name_prefix: str = "") -> "Tree":
        """A copy with every branch length multiplied by ``scaling_factor`` and every node name
        prefixed (reference io/_tree.py:115-129, which goes through a temporary tree file: node
        and edge order are kept, lengths are ``d * scaling_factor``)."""
        res = Tree()
        res.add_nodes([node_name_prefix + v for v in self.nodes()])
        for u, v, d in self._edges:
            res.add_edge(node_name_prefix + u, node_name_prefix + v, float(repr(d * scaling_factor)))
        return res

    def postorder_traversal(self) -> List[str]:
        """Children (in edge order) before their parent; iterative, so deep trees are fine."""
        res, stack = [], [(self.root(), 0)]
        while stack:
            v, k = stack.pop()
            kids = self._children[v]
            if k < len(kids):
                stack.append((v, k + 1))
                stack.append((kids[k][0], 0))
            else:
                res.append(v)
        return res

    def preorder_traversal(self) -> List[str]:
        res, stack = [], [self.root()]
        while stack:
            v = stack.pop()
            res.append(v)
            stack.extend(c for c, _ in reversed(self._children[v]))
        return res


def read_tree(tree_path: str) -> Tree:
    with open(tree_path, "r") as f:
        lines = f.read().strip().split("\n")
    try:
        n, s = lines[0].split(" ")
        if s != "nodes":
            raise Exception
        n = int(n)
    except Exception:
        raise Exception(
            f"Tree file: {tree_path} should start with '[num_nodes] nodes'. "
            f"It started with: '{lines[0]}'"
        )
    tree = Tree()
    for i in range(1, n + 1):
        tree.add_node(lines[i])
    try:
        m, s = lines[n + 1].split(" ")
        if s != "edges":
            raise Exception
        m = int(m)
    except Exception:
        raise Exception(
            f"Tree file: {tree_path} should have line '[num_edges] edges' at "
            f"position {n + 1}, but it had line: '{lines[n + 1]}'"
        )
    if len(lines) != n + 1 + m + 1:
        raise Exception(
            f"Tree file: {tree_path} should have {m} edges, but it has "
            f"{len(lines) - n - 2} edges instead."
        )
    for i in range(n + 2, n + 2 + m):
        try:
            u, v, length = lines[i].split(" ")
            length = float(length)
        except Exception:
            raise Exception(
                f"Tree file: {tree_path} should have line '[u] [v] [length]' at"
                f" position {i}, but it had line: '{lines[i]}'"
            )
        if not tree.is_node(u) or not tree.is_node(v):
            raise Exception(
                f"In Tree file {tree_path}: {u} and {v} should be nodes in the"
                f" tree, but the nodes are: {tree.nodes()}"
            )
        tree.add_edge(u, v, length)
    return tree


def write_tree(tree: Tree, tree_path: str, scaling_factor: float = 1.0, node_name_prefix: str = "") -> None:
    """``scaling_factor`` multiplies every branch length, ``node_name_prefix`` is put in front of
    every node name (reference io/_tree.py:193-211)."""
    pre = node_name_prefix
    out = [f"{tree.num_nodes()} nodes\n"]
    out += [f"{pre + v}\n" for v in tree.nodes()]
    out.append(f"{tree.num_edges()} edges\n")
    out += [f"{pre + u} {pre + v} {d * scaling_factor}\n" for u, v, d in tree.edges()]
    _makedirs_for(tree_path)
    with open(tree_path, "w") as f:
        f.write("".join(out))


# ----------------------------------------------------------------------------- MSAs
def read_msa(msa_path: str) -> Dict[str, str]:
    with open(msa_path, "r") as f:
        lines = f.read().strip().split("\n")
    if len(lines) % 2 != 0:
        raise Exception(f"The MSA at {msa_path} should have an even number of lines")
    msa = {}
    for i in range(len(lines) // 2):
        if not lines[2 * i].startswith(">"):
            raise Exception(
                f"MSA at {msa_path}: at line {2 * i} expected '>[seq_name]' but"
                f" found {lines[2 * i]}"
            )
        msa[lines[2 * i][1:]] = lines[2 * i + 1]
    return msa


def write_msa(msa: Dict[str, str], msa_path: str) -> None:
    _makedirs_for(msa_path)
    with open(msa_path, "w") as f:
        f.write("".join(f">{k}\n{msa[k]}\n" for k in sorted(msa.keys())))


# ----------------------------------------------------------------------- site rates
def read_site_rates(site_rates_path: str) -> List[float]:
    lines = open(site_rates_path).read().strip().split("\n")
    try:
        num_sites, s = lines[0].split(" ")
        if s != "sites":
            raise Exception
        num_sites = int(num_sites)
    except Exception:
        raise Exception(
            f"Site rates file: {site_rates_path} should start with line "
            f"'[num_sites] sites', but started with: {lines[0]} instead."
        )
    try:
        res = list(map(float, lines[1].split(" ")))
    except Exception:
        raise Exception(f"Could nor read site rates in file: {site_rates_path}")
    if len(res) != num_sites:
        raise Exception(
            f"Site rates file: {site_rates_path} was supposed to have "
            f"{num_sites} sites, but it has {len(res)}"
        )
    return res


def write_site_rates(site_rates: Sequence[float], site_rates_path: str) -> None:
    _makedirs_for(site_rates_path)
    with open(site_rates_path, "w") as f:
        f.write(f"{len(site_rates)} sites\n" + " ".join(map(str, site_rates)))


# --------------------------------------------------------------------- contact maps
def read_contact_map(contact_map_path: str) -> np.ndarray:
    lines = open(contact_map_path).read().strip().split("\n")
    try:
        num_sites, s = lines[0].split(" ")
        if s != "sites":
            raise Exception
        num_sites = int(num_sites)
    except Exception:
        raise Exception(
            f"Contact map file should start with line '[num_sites] sites', "
            f"but started with: {lines[0]} instead."
        )
    if len(lines) != num_sites + 1:
        raise Exception(
            f"Contact Map at: {contact_map_path} should have {num_sites} rows, "
            f"but has {len(lines) - 1}"
        )
    rows = np.frombuffer("".join(lines[1:]).encode("ascii"), dtype=np.uint8)
    if rows.size != num_sites * num_sites:
        raise Exception(f"Contact Map at: {contact_map_path} is not square")
    return (rows.reshape(num_sites, num_sites) - ord("0")).astype(int)


def write_contact_map(contact_map: np.ndarray, contact_map_path: str) -> None:
    """``<L> sites`` then L rows of L digits (reference io/_contact_map.py:31-43, np.savetxt with
    fmt="%i" and no delimiter); written as bytes, not through savetxt."""
    _makedirs_for(contact_map_path)
    cm = np.asarray(contact_map)
    ints = cm.astype(np.int64)
    if ints.ndim == 2 and ints.size and ints.min() >= 0 and ints.max() <= 9:
        digits = (ints + ord("0")).astype(np.uint8)
        body = b"".join(row.tobytes() + b"\n" for row in digits)
        with open(contact_map_path, "wb") as f:
            f.write(f"{cm.shape[0]} sites\n".encode() + body)
        return
    with open(contact_map_path, "w") as f:
        f.write(f"{cm.shape[0]} sites\n")
        np.savetxt(f, cm, delimiter="", fmt="%i")


# ------------------------------------------------------------------- count matrices
def _io_threads() -> int:
    return min(16, os.cpu_count() or 1)


def read_count_matrices_array(
    count_matrices_path: str,
) -> Tuple[np.ndarray, List[str], np.ndarray]:
    """Parse ``result.txt`` into ``(q[K], states[S], counts[K,S,S])`` (fp64) with the library's
    multithreaded reader (``cherry_read_count_matrices``; a 129 x 400 x 400 file takes 0.2 s
    instead of 4.4 s).  ``read_count_matrices_array_py`` is the plain-Python definition of the
    format, kept for the tests."""
    import ctypes

    from . import _lib

    lib = _lib.load()
    K, S = ctypes.c_int(0), ctypes.c_int(0)
    path = os.fspath(count_matrices_path).encode()
    _lib.check(lib.cherry_read_count_matrices_header(path, ctypes.byref(K), ctypes.byref(S)),
               "cherry_read_count_matrices_header")
    K, S = K.value, S.value
    q = np.zeros(K)
    counts = np.zeros((K, S, S))
    cap = 64 * S + 64
    names = ctypes.create_string_buffer(cap)
    _lib.check(lib.cherry_read_count_matrices(path, K, S, _lib.ptr(q), _lib.ptr(counts), names, cap, _io_threads()),
               "cherry_read_count_matrices")
    states = names.value.decode().split("\n")[:-1] if K else []
    return q, states, counts


def read_count_matrices_array_py(
    count_matrices_path: str,
) -> Tuple[np.ndarray, List[str], np.ndarray]:
    """Parse ``result.txt`` into ``(q[K], states[S], counts[K,S,S])`` (fp64).

    Accepts both writers' output (the pandas one and the C++ one, whose header line
    starts with a tab and ends with a trailing tab).
    """
    with open(count_matrices_path, "r") as f:
        lines = f.read().strip().split("\n")
    num_matrices, s = lines[0].strip().split(" ")
    if s != "matrices":
        raise Exception(
            f"In file {count_matrices_path}, expected line '[num_matrices] "
            f"matrices', but found: '{lines[0]}'"
        )
    num_states, s = lines[1].strip().split(" ")
    if s != "states":
        raise Exception(
            f"In file {count_matrices_path}, expected line '[num_states] "
            f"states', but found: '{lines[1]}'"
        )
    K, S = int(num_matrices), int(num_states)
    q = np.zeros(K)
    counts = np.zeros((K, S, S))
    states: List[str] = []
    pos = 2
    for k in range(K):
        q[k] = float(lines[pos])
        hdr = lines[pos + 1].strip().split()
        if len(hdr) != S:
            raise Exception(
                f"Error reading count matrices file: {count_matrices_path}\n"
                f"Expected {S} states in line {pos + 1}, but instead "
                f"found {len(hdr)} states: {hdr}"
            )
        states = hdr
        block = " ".join(
            " ".join(lines[pos + 2 + i].split()[1:]) for i in range(S)
        )
        vals = np.array(block.split(), dtype=np.float64)
        if vals.size != S * S:
            raise Exception(f"Could not read count matrices. Matrix {k} is ragged")
        counts[k] = vals.reshape(S, S)
        pos += 2 + S
    return q, states, counts


def read_count_matrices(count_matrices_path: str):
    """Drop-in for ``cherryml.io.read_count_matrices``: ``List[(q, DataFrame)]``."""
    import pandas as pd

    q, states, counts = read_count_matrices_array(count_matrices_path)
    return [
        (float(q[k]), pd.DataFrame(counts[k], index=states, columns=states))
        for k in range(len(q))
    ]


def _fmt_py(x: float) -> str:
    return repr(float(x))


def _fmt_cpp(x: float) -> str:
    # ostream << double at default precision == printf("%g")
    return "%g" % x


def write_count_matrices_array(
    q: Sequence[float],
    states: Sequence[str],
    counts: np.ndarray,
    count_matrices_path: str,
    style: str = "python",
) -> None:
    """Write ``result.txt`` with the library's multithreaded writer (``cherry_write_count_matrices``),
    byte-identical to ``write_count_matrices_array_py`` below (tests/test_fc_native_io.py)."""
    import ctypes

    from . import _lib

    if style not in ("python", "cpp"):
        raise ValueError(f"Unknown style: {style}")
    _makedirs_for(count_matrices_path)
    lib = _lib.load()
    qa = np.ascontiguousarray(np.asarray(q, dtype=np.float64))
    ca = np.ascontiguousarray(np.asarray(counts, dtype=np.float64))
    names = (ctypes.c_char_p * len(states))()
    names[:] = [s.encode() for s in states]
    _lib.check(lib.cherry_write_count_matrices(os.fspath(count_matrices_path).encode(), _lib.ptr(qa), len(qa), names,
                                               len(states), _lib.ptr(ca), int(style == "cpp"), _io_threads()),
               "cherry_write_count_matrices")


def write_count_matrices_array_py(
    q: Sequence[float],
    states: Sequence[str],
    counts: np.ndarray,
    count_matrices_path: str,
    style: str = "python",
) -> None:
    """Write ``result.txt``.

    ``style="python"`` reproduces the pandas ``to_csv(sep="\\t")`` layout with full
    ``repr`` floats (``io/_count_matrices.py:66-81``); ``style="cpp"`` reproduces the
    C++ binary's layout with 6 significant digits (``_count_transitions.cpp:524-548``).
    """
    _makedirs_for(count_matrices_path)
    K, S = len(q), len(states)
    out = [f"{K} matrices\n{S} states\n"]
    if style == "python":
        header = "\t" + "\t".join(states) + "\n"
        for k in range(K):
            out.append(f"{_fmt_py(q[k])}\n")
            out.append(header)
            rows = counts[k].tolist()
            for i in range(S):
                out.append(states[i] + "\t" + "\t".join(map(repr, rows[i])) + "\n")
    elif style == "cpp":
        header = "\t" + "".join(s + "\t" for s in states) + "\n"
        for k in range(K):
            out.append(f"{_fmt_cpp(q[k])}\n")
            out.append(header)
            rows = counts[k].tolist()
            for i in range(S):
                out.append(
                    states[i] + "\t" + "\t".join("%g" % v for v in rows[i]) + "\n"
                )
    else:
        raise ValueError(f"Unknown style: {style}")
    with open(count_matrices_path, "w") as f:
        f.write("".join(out))


def write_count_matrices(count_matrices, count_matrices_path: str) -> None:
    """Drop-in for ``cherryml.io.write_count_matrices`` (list of ``(q, DataFrame)``)."""
    q = [x[0] for x in count_matrices]
    states = list(count_matrices[0][1].index)
    counts = np.stack([np.asarray(x[1], dtype=np.float64) for x in count_matrices])
    write_count_matrices_array(q, states, counts, count_matrices_path, "python")


# ------------------------------------------------------------- rate / mask matrices
def _read_labelled_table(path: str) -> Tuple[List[str], List[str], np.ndarray]:
    with open(path, "r") as f:
        lines = [ln for ln in f.read().split("\n") if ln.strip() != ""]
    cols = lines[0].split()
    rows, data = [], []
    for ln in lines[1:]:
        toks = ln.split()
        rows.append(toks[0])
        data.append([float("nan") if t == "_" else float(t) for t in toks[1:]])
    arr = np.array(data, dtype=np.float64)
    if arr.ndim == 2 and arr.shape[1] == len(cols) - 1:
        cols = cols[1:]  # the header names the index column too ("state\tprob"), as pandas writes it
    if arr.ndim != 2 or arr.shape[1] != len(cols):
        raise Exception(f"Malformed matrix file: {path}")
    return rows, cols, arr


def read_rate_matrix(rate_matrix_path: str):
    import pandas as pd

    rows, cols, arr = _read_labelled_table(rate_matrix_path)
    return pd.DataFrame(arr, index=rows, columns=cols)


def read_mask_matrix(mask_matrix_path: str):
    import pandas as pd

    rows, cols, arr = _read_labelled_table(mask_matrix_path)
    return pd.DataFrame(arr.astype(int), index=rows, columns=cols)


def read_probability_distribution(probability_distribution_path: str):
    import pandas as pd

    rows, cols, arr = _read_labelled_table(probability_distribution_path)
    if arr.shape[1] != 1:
        raise Exception(
            f"Probability distribution at {probability_distribution_path} should be one-dimensional."
        )
    if abs(arr.sum() - 1.0) > 1e-6:
        raise Exception(
            f"Probability distribution at {probability_distribution_path} should add to 1.0, with a "
            "tolerance of 1e-6."
        )
    return pd.DataFrame(arr, index=rows, columns=cols)


def write_rate_matrix(
    rate_matrix: np.ndarray, states: Sequence[str], rate_matrix_path: str
) -> None:
    """Tab-separated labelled square table, floats printed like pandas (repr of the stored
    dtype, so an fp32 matrix prints with fp32 digits as in the reference); fp64 matrices go through
    the library's writer (``cherry_write_labelled_matrix``; 400 x 400: 10 ms instead of 170 ms),
    everything else through ``write_rate_matrix_py``, the plain-Python definition."""
    import ctypes

    arr = np.asarray(rate_matrix)
    n = len(states)
    if arr.dtype != np.float64 or arr.shape != (n, n) or n == 0:
        # (numpy prints float32 scalars with its own exponent thresholds: plain-Python path)
        return write_rate_matrix_py(rate_matrix, states, rate_matrix_path)
    from . import _lib

    _makedirs_for(rate_matrix_path)
    arr = np.ascontiguousarray(arr)
    names = (ctypes.c_char_p * n)()
    names[:] = [s.encode() for s in states]
    _lib.check(_lib.load().cherry_write_labelled_matrix(os.fspath(rate_matrix_path).encode(), names, n, _lib.ptr(arr),
                                                        _io_threads()),
               "cherry_write_labelled_matrix")


def write_rate_matrix_py(
    rate_matrix: np.ndarray, states: Sequence[str], rate_matrix_path: str
) -> None:
    """Tab-separated labelled square table, floats printed like pandas (repr of the
    stored dtype, so an fp32 matrix prints with fp32 digits as in the reference)."""
    _makedirs_for(rate_matrix_path)
    arr = np.asarray(rate_matrix)
    out = ["\t" + "\t".join(states) + "\n"]
    for i, s in enumerate(states):
        # str() of a numpy scalar is the shortest round-trip repr of ITS dtype
        out.append(s + "\t" + "\t".join(str(v) for v in arr[i]) + "\n")
    with open(rate_matrix_path, "w") as f:
        f.write("".join(out))


def write_probability_distribution(
    probability_distribution: np.ndarray, states: Sequence[str], probability_distribution_path: str
) -> None:
    p = np.asarray(probability_distribution).reshape(-1)
    if len(states) != p.shape[0]:
        raise Exception(
            f"probability_distribution has shape {p.shape}, inconsistent with "
            f"states: {states}"
        )
    _makedirs_for(probability_distribution_path)
    with open(probability_distribution_path, "w") as f:
        f.write("state\tprob\n" + "".join(f"{s}\t{v!r}\n" for s, v in zip(states, p.tolist())))


def write_log_likelihood(log_likelihood, log_likelihood_path: str) -> None:
    """``(total, per-site values or None)`` -> total, then ``<n> sites`` and the values
    (reference io/_log_likelihood.py:5-18)."""
    _makedirs_for(log_likelihood_path)
    ll, lls = log_likelihood
    res = f"{ll}\n"
    if lls is not None:
        res += f"{len(lls)} sites\n" + " ".join(map(str, lls))
    with open(log_likelihood_path, "w") as f:
        f.write(res)


def read_log_likelihood(log_likelihood_path: str):
    """-> ``(total, per-site values or None)`` (reference io/_log_likelihood.py:21-47)."""
    with open(log_likelihood_path) as f:
        lines = f.read().strip().split("\n")
    ll = float(lines[0])
    if len(lines) == 1:
        return ll, None
    try:
        num_sites, s = lines[1].split(" ")
        if s != "sites":
            raise Exception
        num_sites = float(num_sites)
    except Exception:
        raise Exception(
            f"Log likelihood file at:{log_likelihood_path} should have second line '[num_sites] sites', "
            f"but had second line: {lines[1]} instead."
        )
    lls = list(map(float, lines[2].split(" "))) if len(lines) > 2 and lines[2] else []
    if len(lls) != num_sites:
        raise Exception(
            f"Log likelihood file at:{log_likelihood_path} should have {num_sites} values in line 3,"
            f"but had {len(lls)} values instead."
        )
    return ll, lls


def read_computed_cherries_from_file(file_path: str):
    """FastCherries' raw output (name, name, distance per cherry) -> ``(cherries, distances)``
    (reference io/_rate_matrix.py:101-119)."""
    with open(file_path) as f:
        lines = [ln.strip() for ln in f.readlines()]
    cherries = [(lines[i], lines[i + 1]) for i in range(0, len(lines) - 2, 3)]
    distances = [float(lines[i + 2]) for i in range(0, len(lines) - 2, 3)]
    return cherries, distances


def get_msa_num_sites(msa_path: str) -> int:
    """Length of the first sequence line, without reading the rest (reference io/_msa.py:5-14)."""
    with open(msa_path) as f:
        f.readline()
        line = f.readline()
    if line == "":
        raise Exception("We shouldn't be here!")
    return len(line.strip())


def get_msa_num_sequences(msa_path: str) -> int:
    """Reference io/_msa.py:41-48."""
    return len(read_msa(msa_path))


def get_msa_num_residues(msa_path: str, exclude_gaps: bool) -> int:
    """Cells of the alignment, optionally without the gap characters ``. - _`` (reference io/_msa.py:17-38)."""
    seqs = list(read_msa(msa_path).values())
    if not exclude_gaps:
        return len(seqs) * len(seqs[0])
    return sum(len(q) - q.count(".") - q.count("-") - q.count("_") for q in seqs)


TransitionsType = List[Tuple[str, str, float]]


def _counted_lines(path: str, what: str) -> List[str]:
    """Body of a ``<n> transitions`` file, after checking the header against the line count."""
    with open(path) as f:
        lines = f.read().strip().split("\n")
    tokens = lines[0].split(" ")
    if len(tokens) != 2 or tokens[1] != "transitions":
        raise ValueError(f"{what} file at '{path}' should start with '[NUM_TRANSITIONS] transitions'.")
    if len(lines) - 1 != int(tokens[0]):
        raise ValueError(f"Expected {int(tokens[0])} transitions at '{path}', but found only {len(lines) - 1}.")
    return lines[1:]


def read_transitions(transitions_path: str) -> TransitionsType:
    """``<n> transitions`` then ``x y t`` per line (reference io/_transitions.py:7-36)."""
    res = []
    for line in _counted_lines(transitions_path, "Transitions"):
        x, y, t = line.split(" ")
        res.append((x, y, float(t)))
    return res


def write_transitions(transitions: TransitionsType, transitions_path: str) -> None:
    """Reference io/_transitions.py:39-52; times are written with Python's ``str``."""
    _makedirs_for(transitions_path)
    with open(transitions_path, "w") as f:
        f.write(f"{len(transitions)} transitions\n" + "\n".join(f"{x} {y} {t}" for x, y, t in transitions) + "\n")


def read_transitions_log_likelihood(transitions_log_likelihood_path: str) -> List[float]:
    """``<n> transitions`` then one log-likelihood per line (reference io/_transitions_log_likelihood.py:7-43)."""
    return [float(line) for line in _counted_lines(transitions_log_likelihood_path, "Transitions log likelihood")]


def write_transitions_log_likelihood(transitions_log_likelihood: Sequence[float], transitions_log_likelihood_path: str) -> None:
    """Reference io/_transitions_log_likelihood.py:46-62."""
    _makedirs_for(transitions_log_likelihood_path)
    with open(transitions_log_likelihood_path, "w") as f:
        f.write(
            f"{len(transitions_log_likelihood)} transitions\n"
            + "\n".join(str(ll) for ll in transitions_log_likelihood)
            + "\n"
        )


def read_pickle(pickle_path: str):
    """Reference io/_pickle.py:5-9."""
    import pickle

    with open(pickle_path, "rb") as f:
        return pickle.load(f)


def write_pickle(obj, output_path: str) -> None:
    """Reference io/_pickle.py:12-18."""
    import pickle

    with open(output_path, "wb") as f:
        pickle.dump(obj, f)


def read_transitions_log_likelihood_per_site(transitions_log_likelihood_per_site_path: str) -> List[List[float]]:
    """A pickled list of per-site log-likelihood lists (reference io/_transitions_log_likelihood_per_site.py:6-15)."""
    res = read_pickle(transitions_log_likelihood_per_site_path)
    if len(res) == 0:
        raise Exception(
            f"The transitions log likelihood file at {transitions_log_likelihood_per_site_path} is empty"
        )
    return res


def write_transitions_log_likelihood_per_site(
    transitions_log_likelihood_per_site: List[List[float]], transitions_log_likelihood_per_site_path: str
) -> None:
    """Reference io/_transitions_log_likelihood_per_site.py:17-31."""
    _makedirs_for(transitions_log_likelihood_per_site_path)
    write_pickle(transitions_log_likelihood_per_site, transitions_log_likelihood_per_site_path)


def read_str(s_path: str) -> str:
    with open(s_path) as f:
        return f.read()


def write_str(s: str, s_path: str) -> None:
    with open(s_path, "w") as f:
        f.write(s)


def read_sites_subset(sites_subset_path: str) -> List[int]:
    """``<n> sites`` then n space-separated site indices (reference io/_sites_subset.py:5-32)."""
    with open(sites_subset_path) as f:
        lines = f.read().strip().split("\n")
    try:
        num_sites, s = lines[0].split(" ")
        if s != "sites":
            raise Exception
        num_sites = int(num_sites)
    except Exception:
        raise Exception(
            f"Sites subset file: {sites_subset_path} should start with line "
            f"'[num_sites] sites', but started with: {lines[0]} instead."
        )
    try:
        res = [] if num_sites == 0 else list(map(int, lines[1].split(" ")))
    except Exception:
        raise Exception(f"Could nor read sites subset in file: {sites_subset_path}. Lines: {lines}")
    if len(res) != num_sites:
        raise Exception(
            f"Sites subset file: {sites_subset_path} was supposed to have {num_sites} sites, but it has {len(res)}"
        )
    return res


def write_sites_subset(sites_subset: Sequence[int], sites_subset_path: str) -> None:
    _makedirs_for(sites_subset_path)
    with open(sites_subset_path, "w") as f:
        f.write(f"{len(sites_subset)} sites\n" + " ".join(map(str, sites_subset)))


def _makedirs_for(path: str) -> None:
    d = os.path.dirname(path)
    if d != "" and not os.path.exists(d):
        os.makedirs(d, exist_ok=True)
