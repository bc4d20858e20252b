"""Host-side ingest for counting: trees -> leaf pairs, MSAs -> integer-encoded rows.

Everything here is host logic that runs once per family; the per-site / per-contact work
happens on the GPU.  What is mirrored from the reference:

* pair extraction for ``cherry++`` (post-order pairing of unmatched leaves, fp64 distance
  sums in traversal order), ``cherry`` and ``edge``:
  ``counting/_count_transitions.py:65-186`` / ``counting/_count_transitions.cpp:316-390,
  444-506``;
* the two "personalities" of branch lengths: the C++ binary parses them with ``std::stof``
  (float32, ``_count_transitions.cpp:247``) while the Python implementation keeps fp64
  (``io/_tree.py:251``);
* contact pairs ``(i < j, j - i >= d, map[i][j] == 1)``:
  ``counting/_count_co_transitions.py:74-79`` / ``.cpp:433-442``.

Layout produced (see DESIGN.md "Data layout in HBM"): one flat uint8 residue buffer; per
family a descriptor; rows 16-byte aligned; for LG the columns are sorted by site-rate
category with every category padded to a multiple of 4 sites; for co-transitions the rows
are contact-paired (bytes 2c, 2c+1 = the two sites of contact c).
"""
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .._lib import FAM_DESC_DTYPE, TILE_DTYPE
from ..io import Tree, read_contact_map, read_msa, read_site_rates, read_tree

TARGET_CHUNKS_PER_TILE = 16384  # ~16 loop trips of a 1024-thread CTA
TARGET_ITEMS_PER_CO_TILE = 8192


def _f32(x: float) -> float:
    return float(np.float32(x))


def extract_pairs(
    tree: Tree, edge_or_cherry: str, float32_branch_lengths: bool
) -> List[Tuple[str, str, float]]:
    """``[(node_a, node_b, t)]`` for one tree, in the reference's visiting order."""
    bl = _f32 if float32_branch_lengths else float
    pairs: List[Tuple[str, str, float]] = []
    if edge_or_cherry == "cherry++":
        # iterative post-order; result[v] = (unmatched_leaf or None, distance)
        result: Dict[str, Tuple[Optional[str], float]] = {}
        root = tree.root()
        stack: List[Tuple[str, bool]] = [(root, False)]
        while stack:
            node, expanded = stack.pop()
            if tree.is_leaf(node):
                result[node] = (node, 0.0)
                continue
            if not expanded:
                stack.append((node, True))
                for child, _ in reversed(tree.children(node)):
                    stack.append((child, False))
                continue
            leaves_under: List[str] = []
            dists_under: List[float] = []
            for child, length in tree.children(node):
                leaf, dist = result.pop(child)
                if leaf is not None:
                    leaves_under.append(leaf)
                    dists_under.append(dist + bl(length))
            for i in range(0, len(leaves_under) - 1, 2):
                pairs.append(
                    (leaves_under[i], leaves_under[i + 1], dists_under[i] + dists_under[i + 1])
                )
            if len(leaves_under) % 2 == 0:
                result[node] = (None, -1.0)
            else:
                result[node] = (leaves_under[-1], dists_under[-1])
        n_leaves = len(tree.leaves())
        if len(pairs) != n_leaves // 2:
            raise AssertionError(
                f"cherry++ produced {len(pairs)} pairs for {n_leaves} leaves"
            )
    elif edge_or_cherry == "cherry":
        for node in tree.nodes():
            ch = tree.children(node)
            if len(ch) == 2 and all(tree.is_leaf(c) for c, _ in ch):
                pairs.append((ch[0][0], ch[1][0], bl(ch[0][1]) + bl(ch[1][1])))
    elif edge_or_cherry == "edge":
        for node in tree.nodes():
            for child, length in tree.children(node):
                pairs.append((node, child, bl(length)))
    else:
        raise ValueError(f"Unknown edge_or_cherry: {edge_or_cherry!r}")
    return pairs


def contacting_pairs(contact_map: np.ndarray, minimum_distance: int) -> np.ndarray:
    """``int32 [n, 2]`` of ``(i, j)``, ``i < j``, ``j - i >= minimum_distance``, row-major."""
    ii, jj = np.nonzero(contact_map == 1)
    keep = (jj > ii) & (jj - ii >= minimum_distance)
    return np.stack([ii[keep], jj[keep]], axis=1).astype(np.int32)


def contact_paired_rows(enc: np.ndarray, contacts: np.ndarray, skip: int) -> np.ndarray:
    """Rows in the co-transition kernels' layout: bytes ``2c`` and ``2c+1`` are the residues at
    the two sites of contact ``c``; padded with the skip code to a multiple of 16 bytes.  The
    kernel then streams 4 bytes per (pair, contact) and never gathers."""
    n_rows, P = enc.shape[0], len(contacts)
    stride = max(16, (2 * P + 15) // 16 * 16)
    rows = np.full((n_rows, stride), skip, dtype=np.uint8)
    if n_rows and P:
        rows[:, 0 : 2 * P : 2] = enc[:, contacts[:, 0]]
        rows[:, 1 : 2 * P : 2] = enc[:, contacts[:, 1]]
    return rows


def alphabet_lut(states: Sequence[str]) -> np.ndarray:
    if len(states) > 254:
        raise ValueError("at most 254 states are supported")
    # every byte that is not a state maps to the skip code S == len(states)
    lut = np.full(256, len(states), dtype=np.uint8)
    for i, s in enumerate(states):
        if len(s) != 1 or ord(s) > 255:
            raise ValueError(f"states must be single one-byte characters, got {s!r}")
        lut[ord(s)] = i
    return lut


@dataclass
class CountBatch:
    """Integer-encoded families ready for the counting kernels (host numpy arrays)."""

    kind: str  # "lg" or "co"
    msa: np.ndarray  # uint8 flat
    fams: np.ndarray  # FAM_DESC_DTYPE [F]
    pair_a: np.ndarray  # int32 [P] row index inside the family
    pair_b: np.ndarray  # int32 [P]
    pair_t: np.ndarray  # float64 [P]
    pair_fam: np.ndarray  # int32 [P]
    rate_vals: np.ndarray  # float64 flat (LG: distinct site rates per family; co: 1.0)
    aux: np.ndarray  # LG: uint16 group categories; co: int32 [n,2] contacts (host only)
    tiles: np.ndarray  # TILE_DTYPE [T]
    r_pad: int
    n_sites_examined: int = 0  # (pair, site) or (pair, contact) items, before validity
    family_names: List[str] = field(default_factory=list)

    @property
    def n_pairs(self) -> int:
        return int(self.pair_a.shape[0])

    @property
    def n_fams(self) -> int:
        return int(self.fams.shape[0])


class _BatchBuilder:
    def __init__(self, kind: str):
        self.kind = kind
        self.msa_parts: List[np.ndarray] = []
        self.msa_bytes = 0
        self.fams: List[tuple] = []
        self.pair_a: List[np.ndarray] = []
        self.pair_b: List[np.ndarray] = []
        self.pair_t: List[np.ndarray] = []
        self.pair_fam: List[np.ndarray] = []
        self.rate_vals: List[np.ndarray] = []
        self.n_rate_vals = 0
        self.aux: List[np.ndarray] = []
        self.n_aux = 0
        self.tiles: List[tuple] = []
        self.n_pairs = 0
        self.max_rates = 1
        self.examined = 0
        self.names: List[str] = []

    def add_family(
        self,
        name: str,
        rows: np.ndarray,  # uint8 [n_rows, row_stride], already laid out
        pair_a: np.ndarray,
        pair_b: np.ndarray,
        pair_t: np.ndarray,
        rate_vals: np.ndarray,
        aux: np.ndarray,
        aux_cnt: int,
        items_per_pair: int,
    ) -> None:
        f = len(self.fams)
        n_rows, stride = rows.shape
        assert stride % 16 == 0
        self.fams.append(
            (self.msa_bytes, stride, stride // 16, self.n_aux, aux_cnt, self.n_rate_vals,
             len(rate_vals))
        )
        self.msa_parts.append(np.ascontiguousarray(rows).reshape(-1))
        self.msa_bytes += n_rows * stride
        self.rate_vals.append(np.asarray(rate_vals, dtype=np.float64))
        self.n_rate_vals += len(rate_vals)
        self.max_rates = max(self.max_rates, len(rate_vals))
        self.aux.append(aux)
        self.n_aux += len(aux)
        npairs = len(pair_a)
        self.pair_a.append(np.asarray(pair_a, dtype=np.int32))
        self.pair_b.append(np.asarray(pair_b, dtype=np.int32))
        self.pair_t.append(np.asarray(pair_t, dtype=np.float64))
        self.pair_fam.append(np.full(npairs, f, dtype=np.int32))
        if self.kind == "lg":
            per_tile = max(1, TARGET_CHUNKS_PER_TILE // max(1, stride // 16))
        else:
            per_tile = max(1, TARGET_ITEMS_PER_CO_TILE // max(1, aux_cnt))
        for b in range(0, npairs, per_tile):
            self.tiles.append((f, self.n_pairs + b, min(per_tile, npairs - b), 0))
        self.n_pairs += npairs
        self.examined += npairs * items_per_pair
        self.names.append(name)

    def finish(self) -> CountBatch:
        def cat(parts, dtype, shape_tail=()):
            if parts:
                return np.ascontiguousarray(np.concatenate(parts).astype(dtype, copy=False))
            return np.zeros((0,) + shape_tail, dtype=dtype)

        fams = np.array(self.fams, dtype=FAM_DESC_DTYPE) if self.fams else np.zeros(0, FAM_DESC_DTYPE)
        tiles = np.array(self.tiles, dtype=TILE_DTYPE) if self.tiles else np.zeros(0, TILE_DTYPE)
        if self.kind == "lg":
            aux = cat(self.aux, np.uint16)
            r_pad = (self.max_rates + 3) // 4 * 4
        else:
            aux = cat(self.aux, np.int32, (2,)).reshape(-1, 2)
            r_pad = 4
        msa = cat(self.msa_parts, np.uint8)
        if msa.size == 0:
            msa = np.zeros(16, dtype=np.uint8)
        return CountBatch(
            kind=self.kind,
            msa=msa,
            fams=fams,
            pair_a=cat(self.pair_a, np.int32),
            pair_b=cat(self.pair_b, np.int32),
            pair_t=cat(self.pair_t, np.float64),
            pair_fam=cat(self.pair_fam, np.int32),
            rate_vals=cat(self.rate_vals, np.float64),
            aux=aux,
            tiles=tiles,
            r_pad=r_pad,
            n_sites_examined=self.examined,
            family_names=self.names,
        )


def _rows_for_pairs(pairs: Sequence[Tuple[str, str, float]]):
    """Assign a row to every distinct node, in order of first use (partners adjacent)."""
    row_of: Dict[str, int] = {}
    a = np.empty(len(pairs), dtype=np.int32)
    b = np.empty(len(pairs), dtype=np.int32)
    t = np.empty(len(pairs), dtype=np.float64)
    for i, (u, v, d) in enumerate(pairs):
        a[i] = row_of.setdefault(u, len(row_of))
        b[i] = row_of.setdefault(v, len(row_of))
        t[i] = d
    return list(row_of.keys()), a, b, t


def _encode_rows(msa: Dict[str, str], names: Sequence[str], lut: np.ndarray, family: str) -> np.ndarray:
    if not names:
        return np.zeros((0, 0), dtype=np.uint8)
    try:
        seqs = [msa[n] for n in names]
    except KeyError as e:
        raise Exception(f"Family {family}: node {e} of the tree is not in the MSA")
    L = len(seqs[0])
    if any(len(s) != L for s in seqs):
        raise Exception(f"Family {family}: sequences in the MSA have different lengths")
    raw = np.frombuffer("".join(seqs).encode("latin-1"), dtype=np.uint8).reshape(len(seqs), L)
    return lut[raw]


def lg_column_layout(site_rates: Sequence[float]):
    """Sort columns by rate category, pad each category to a multiple of 4 sites.

    Returns ``(rate_vals[R], dest_col[L], group_cat[row_stride/4], row_stride)``."""
    rates = np.asarray(site_rates, dtype=np.float64)
    vals, inv = np.unique(rates, return_inverse=True)
    if len(vals) > 65535:
        raise ValueError("more than 65535 distinct site rates in one family")
    order = np.argsort(inv, kind="stable")
    counts = np.bincount(inv, minlength=len(vals))
    padded = (counts + 3) // 4 * 4
    starts = np.concatenate([[0], np.cumsum(padded)[:-1]])
    dest = np.empty(len(rates), dtype=np.int64)
    within = np.arange(len(rates)) - np.repeat(np.concatenate([[0], np.cumsum(counts)[:-1]]), counts)
    dest[order] = np.repeat(starts, counts) + within
    total = int(padded.sum())
    stride = max(16, (total + 15) // 16 * 16)
    group_cat = np.zeros(stride // 4, dtype=np.uint16)
    group_cat[: total // 4] = np.repeat(np.arange(len(vals), dtype=np.uint16), padded // 4)
    return vals, dest, group_cat, stride


def encode_lg_family(
    builder: _BatchBuilder,
    name: str,
    tree: Tree,
    msa: Dict[str, str],
    site_rates: Sequence[float],
    lut: np.ndarray,
    edge_or_cherry: str,
    float32_branch_lengths: bool,
) -> None:
    pairs = extract_pairs(tree, edge_or_cherry, float32_branch_lengths)
    names, a, b, t = _rows_for_pairs(pairs)
    enc = _encode_rows(msa, names, lut, name)
    if enc.shape[0] and enc.shape[1] > len(site_rates):
        raise Exception(
            f"Family {name}: MSA has {enc.shape[1]} sites but there are only "
            f"{len(site_rates)} site rates"
        )
    if enc.shape[0]:
        # the reference indexes site_rates by MSA position, so surplus rates are ignored
        site_rates = list(site_rates)[: enc.shape[1]]
    L = len(site_rates)
    vals, dest, group_cat, stride = lg_column_layout(site_rates) if L else (
        np.ones(1), np.zeros(0, dtype=np.int64), np.zeros(4, dtype=np.uint16), 16)
    rows = np.full((enc.shape[0], stride), int(lut.max()), dtype=np.uint8)
    if enc.shape[0] and L:
        rows[:, dest] = enc
    builder.add_family(name, rows, a, b, t, vals, group_cat, stride // 4, L)


def encode_co_family(
    builder: _BatchBuilder,
    name: str,
    tree: Tree,
    msa: Dict[str, str],
    contact_map: np.ndarray,
    lut: np.ndarray,
    edge_or_cherry: str,
    minimum_distance: int,
    float32_branch_lengths: bool,
) -> None:
    pairs = extract_pairs(tree, edge_or_cherry, float32_branch_lengths)
    names, a, b, t = _rows_for_pairs(pairs)
    enc = _encode_rows(msa, names, lut, name)
    contacts = contacting_pairs(contact_map, minimum_distance)
    L = enc.shape[1] if enc.shape[0] else contact_map.shape[0]
    if len(contacts) and contacts.max() >= L:
        raise Exception(f"Family {name}: contact map is larger than the MSA")
    rows = contact_paired_rows(enc, contacts, int(lut.max()))
    builder.add_family(name, rows, a, b, t, np.ones(1), contacts, len(contacts), len(contacts))


def build_lg_batch(
    tree_dir: str,
    msa_dir: str,
    site_rates_dir: str,
    families: Sequence[str],
    states: Sequence[str],
    edge_or_cherry: str,
    float32_branch_lengths: bool,
) -> CountBatch:
    lut = alphabet_lut(states)
    builder = _BatchBuilder("lg")
    for fam in families:
        tree = read_tree(os.path.join(tree_dir, fam + ".txt"))
        msa = read_msa(os.path.join(msa_dir, fam + ".txt"))
        rates = read_site_rates(os.path.join(site_rates_dir, fam + ".txt"))
        encode_lg_family(builder, fam, tree, msa, rates, lut, edge_or_cherry, float32_branch_lengths)
    return builder.finish()


def build_co_batch(
    tree_dir: str,
    msa_dir: str,
    contact_map_dir: str,
    families: Sequence[str],
    states: Sequence[str],
    edge_or_cherry: str,
    minimum_distance: int,
    float32_branch_lengths: bool,
) -> CountBatch:
    lut = alphabet_lut(states)
    builder = _BatchBuilder("co")
    for fam in families:
        tree = read_tree(os.path.join(tree_dir, fam + ".txt"))
        msa = read_msa(os.path.join(msa_dir, fam + ".txt"))
        cmap = read_contact_map(os.path.join(contact_map_dir, fam + ".txt"))
        encode_co_family(
            builder, fam, tree, msa, cmap, lut, edge_or_cherry, minimum_distance,
            float32_branch_lengths,
        )
    return builder.finish()


# ------------------------------------------------------------------ native (C++) ingest
class _NativeHandle:
    """Keeps a ``cherry_ingest_result`` alive while numpy views of its arrays exist."""

    def __init__(self, lib, ptr):
        self.lib, self.ptr = lib, ptr

    def __del__(self):
        try:
            if self.ptr:
                self.lib.cherry_ingest_free(self.ptr)
                self.ptr = None
        except Exception:
            pass


def _native_batch(kind: str, call, families: Sequence[str]) -> CountBatch:
    import ctypes

    from .. import _lib

    lib = _lib.load()
    res = ctypes.POINTER(_lib.IngestResult)()
    rc = call(lib, ctypes.byref(res))
    _lib.check(rc, "cherry_ingest_" + kind)
    r = res.contents
    handle = _NativeHandle(lib, ctypes.cast(res, ctypes.c_void_p))

    def arr(addr, n, dtype):
        if n == 0 or not addr:
            return np.zeros((0,), dtype=dtype)
        nbytes = int(n) * np.dtype(dtype).itemsize
        buf = (ctypes.c_uint8 * nbytes).from_address(addr)
        buf._cherry_owner = handle  # numpy keeps `buf` alive, `buf` keeps the allocation alive
        return np.frombuffer(buf, dtype=dtype, count=int(n))

    if kind == "lg":
        aux = arr(r.aux, r.n_aux, np.uint16)
    else:
        aux = arr(r.aux, r.n_aux * 2, np.int32).reshape(-1, 2)
    batch = CountBatch(
        kind=kind,
        msa=arr(r.msa, r.msa_bytes, np.uint8),
        fams=arr(r.fams, r.n_fams, FAM_DESC_DTYPE),
        pair_a=arr(r.pair_a, r.n_pairs, np.int32),
        pair_b=arr(r.pair_b, r.n_pairs, np.int32),
        pair_t=arr(r.pair_t, r.n_pairs, np.float64),
        pair_fam=arr(r.pair_fam, r.n_pairs, np.int32),
        rate_vals=arr(r.rate_vals, r.n_rate_vals, np.float64),
        aux=aux,
        tiles=arr(r.tiles, r.n_tiles, TILE_DTYPE),
        r_pad=int(r.r_pad),
        n_sites_examined=int(r.n_items_examined),
        family_names=list(families),
    )
    batch.msa_pinned = bool(r.pinned)
    return batch


def _c_strings(items: Sequence[str]):
    import ctypes

    enc = [s.encode("utf-8") for s in items]
    return (ctypes.c_char_p * max(1, len(enc)))(*enc) if enc else (ctypes.c_char_p * 1)()


def default_ingest_threads() -> int:
    return max(1, min(64, os.cpu_count() or 1))


def build_lg_batch_native(
    tree_dir: str, msa_dir: str, site_rates_dir: str, families: Sequence[str], states: Sequence[str],
    edge_or_cherry: str, float32_branch_lengths: bool, n_threads: Optional[int] = None, pinned: bool = False,
) -> CountBatch:
    """``build_lg_batch`` done by the library's multithreaded C++ ingest (``cherry_ingest_lg``)."""
    fam_c, st_c = _c_strings(families), _c_strings(states)
    nt = n_threads or default_ingest_threads()

    def call(lib, out):
        return lib.cherry_ingest_lg(
            tree_dir.encode(), msa_dir.encode(), site_rates_dir.encode(), fam_c, len(families), st_c,
            len(states), edge_or_cherry.encode(), int(float32_branch_lengths), nt, int(pinned), out)

    return _native_batch("lg", call, families)


def build_co_batch_native(
    tree_dir: str, msa_dir: str, contact_map_dir: str, families: Sequence[str], states: Sequence[str],
    edge_or_cherry: str, minimum_distance: int, float32_branch_lengths: bool,
    n_threads: Optional[int] = None, pinned: bool = False,
) -> CountBatch:
    """``build_co_batch`` done by the library's multithreaded C++ ingest (``cherry_ingest_co``)."""
    fam_c, st_c = _c_strings(families), _c_strings(states)
    nt = n_threads or default_ingest_threads()

    def call(lib, out):
        return lib.cherry_ingest_co(
            tree_dir.encode(), msa_dir.encode(), contact_map_dir.encode(), fam_c, len(families), st_c,
            len(states), edge_or_cherry.encode(), int(minimum_distance), int(float32_branch_lengths), nt,
            int(pinned), out)

    return _native_batch("co", call, families)
