"""FastCherries tree estimation on the GPU.

``fast_cherries`` keeps the signature, defaults, caching behaviour and output files of the
reference's ``cherryml/phylogeny_estimation/_fast_cherries.py:191-281``; the C++ program it
shells out to (``FastCherries/fast_cherries.cpp``) is replaced by two CUDA kernels
(``cherry_fc_pair``, ``cherry_fc_ble``) that process all families of the call in one launch
each.  The small set-up computations of ``fast_cherries.cpp:47-215`` (quantization grid, rate
categories, prior weights) are host scalars; the table of log transition probabilities
(``io_helpers.cpp:150-176``) is computed on the device with ``cherry_expm_batched``.

There is no CPU fallback: without the CUDA library / a GPU the stage raises.
"""
import ctypes
import math
import os
import time
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .. import _lib
from ..caching import cached_parallel_computation
from ..io import Tree, read_rate_matrix
from ..markov_chain import expm_batched


# ----------------------------------------------------------------------------- set-up scalars

def quantization_grid(center: float, step: float, num_steps: int) -> np.ndarray:
    """``compute_quantization_points`` (io_helpers.cpp:178-194): a chain of multiplications /
    divisions in x87 extended precision starting at the centre, narrowed to fp64 at the end."""
    if np.finfo(np.longdouble).nmant != 63:
        raise _lib.CherryError("FastCherries grid needs an 80-bit numpy longdouble (x86-64)")
    ext = np.zeros(2 * num_steps + 1, dtype=np.longdouble)
    ext[num_steps] = center
    ratio = np.longdouble(step)
    for i in range(1, num_steps + 1):
        ext[num_steps + i] = ext[num_steps + i - 1] * ratio
        ext[num_steps - i] = ext[num_steps - i + 1] / ratio
    return np.asarray(ext, dtype=np.float64)


def ble_rate_categories(num_rate_categories: int) -> np.ndarray:
    """Geometric ladder from 1/R to R (fast_cherries.cpp:205-213); a single category is 1."""
    R = int(num_rate_categories)
    if R < 1:
        raise ValueError("num_rate_categories must be >= 1")
    if R == 1:
        return np.ones(1)
    first = 1.0 / R
    ratio = math.pow(R / first, 1.0 / (R - 1))
    cats = [first]
    while len(cats) < R:
        cats.append(cats[-1] * ratio)
    return np.array(cats)


def _log_gamma(a: float) -> float:
    # Pike & Hill, CACM Algorithm 291 (what fast_cherries.cpp:47-66 uses): push the argument
    # up to >= 7 by the recurrence, then Stirling's series.
    shift = 0.0
    x = a
    if x < 7:
        prod = 1.0
        z = x
        while z < 7:
            prod *= z
            z += 1.0
        x = z
        shift = -math.log(prod)
    inv2 = 1.0 / (x * x)
    series = (((-.000595238095238 * inv2 + .000793650793651) * inv2 - .002777777777778) * inv2 + .083333333333333) / x
    return shift + (x - 0.5) * math.log(x) - x + .918938533204673 + series


def _gamma_cdf(x: float, shape: float) -> float:
    """Regularised lower incomplete gamma P(shape, x) by Bhattacharjee's AS 32 (series for small
    x, continued fraction otherwise) with the 1e-8 stopping rule of fast_cherries.cpp:69-128."""
    tol, big = 1e-8, 1e30
    if x == 0:
        return 0.0
    if x < 0 or shape <= 0:
        return -1.0
    scale = math.exp(shape * math.log(x) - x - _log_gamma(shape))
    if x <= 1 or x < shape:
        total, term, denom = 1.0, 1.0, shape
        while True:
            denom += 1
            term *= x / denom
            total += term
            if term <= tol:
                return total * (scale / shape)
    a = 1 - shape
    b = a + x + 1
    n = 0.0
    p = [1.0, x, x + 1, x * b]
    cur = p[2] / p[3]
    while True:
        a += 1
        b += 2
        n += 1
        an = a * n
        p4 = b * p[2] - an * p[0]
        p5 = b * p[3] - an * p[1]
        if p5 != 0:
            nxt = p4 / p5
            gap = abs(cur - nxt)
            if gap <= tol and gap <= tol * nxt:
                return 1 - scale * cur
            cur = nxt
        p = [p[2], p[3], p4, p5]
        if abs(p4) >= big:
            p = [v / big for v in p]


def initial_rate_weights(cats: np.ndarray) -> np.ndarray:
    """CDF of gamma(shape 3, mean 1) at the geometric midpoints of neighbouring categories, last
    entry 1 (fast_cherries.cpp:137-160)."""
    shape = 3.0
    w = [_gamma_cdf(math.sqrt(cats[i - 1] * cats[i]) * shape, shape) for i in range(1, len(cats))]
    return np.array(w + [1.0])


# ----------------------------------------------------------------------------- MSAs -> residue rows

def parse_msa_like_cpp(path: str) -> Tuple[List[str], List[bytes]]:
    """``read_msa`` of io_helpers.cpp:35-74: a line starting with '>' names a sequence, the
    line after it is the sequence."""
    with open(path, "rb") as f:
        lines = f.read().split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()
    names: List[str] = []
    seqs: List[bytes] = []
    i = 0
    while i < len(lines):
        ln = lines[i]
        if ln[:1] == b">":
            if i + 1 >= len(lines):
                break
            names.append(ln[1:].decode("utf-8", "replace"))
            seqs.append(lines[i + 1])
            i += 2
        else:
            i += 1
    return names, seqs


def encode_families(msa_paths: Sequence[str], alphabet: Sequence[str]):
    """-> (names per family, flat uint8 residue buffer, FC_FAMILY_DTYPE array)."""
    S = len(alphabet)
    lut = bytearray([S]) * 256
    for i, ch in enumerate(alphabet):
        if len(ch) != 1:
            raise ValueError("FastCherries alphabets are single characters")
        lut[ord(ch)] = i
    lut = bytes(lut)
    fams = np.zeros(len(msa_paths), dtype=_lib.FC_FAMILY_DTYPE)
    all_names, blocks = [], []
    off = cherry_off = site_off = seq_off = 0
    for f, path in enumerate(msa_paths):
        names, seqs = parse_msa_like_cpp(path)
        n = len(names)
        L = len(seqs[0]) if n else 0
        if any(len(s) != L for s in seqs):
            raise ValueError(f"MSA {path}: sequences of different lengths")
        if n > 65535:
            raise _lib.CherryError(f"MSA {path}: more than 65535 sequences")
        stride = max(16, (L + 15) // 16 * 16)
        rows = np.full((n, stride), S, dtype=np.uint8)
        if n and L:
            rows[:, :L] = np.frombuffer(b"".join(s.translate(lut) for s in seqs), dtype=np.uint8).reshape(n, L)
        fams[f] = (off, n, stride, L, cherry_off, site_off, seq_off)
        blocks.append(rows.reshape(-1))
        all_names.append(names)
        off += n * stride
        cherry_off += n // 2
        site_off += L
        seq_off += n
    buf = np.concatenate(blocks) if blocks else np.zeros(0, dtype=np.uint8)
    return all_names, buf, fams


# ----------------------------------------------------------------------------- device pipeline

def log_transition_table(Q: np.ndarray, grid: np.ndarray, cats: np.ndarray, device) -> torch.Tensor:
    """fp64 [K][R][S][S] on the device: log P + (log P)^T, P = expm(q_k * rate_r * Q)."""
    K, R, S = len(grid), len(cats), Q.shape[0]
    exponents = (grid[:, None] * cats[None, :]).reshape(-1)
    logp = torch.log(expm_batched(Q, exponents, device))
    return (logp + logp.transpose(1, 2)).reshape(K, R, S, S).contiguous()


def fast_cherries_device(msa: np.ndarray, fams: np.ndarray, S: int, sym_table: torch.Tensor, priors: np.ndarray,
                         weights: np.ndarray, seed: int, max_iters: int, device="cuda",
                         keep_device: bool = False) -> Dict[str, np.ndarray]:
    """Runs both kernels on an encoded batch.  Returns host arrays: pair_a, pair_b (row indices
    per cherry, reference emission order), unpaired (per family), len_idx (per cherry),
    site_cat (per site), iters (per family), and the two kernel times in ms."""
    lib = _lib.load()
    dev = torch.device(device)
    n_fams = len(fams)
    total_seqs = int(fams["n_seqs"].sum())
    total_sites = int(fams["n_sites"].sum())
    n_cherries = int((fams["n_seqs"] // 2).sum())
    K, R = int(sym_table.shape[0]), int(sym_table.shape[1])
    with torch.cuda.device(dev):
        d_msa = torch.from_numpy(msa).to(dev)
        d_fams = torch.from_numpy(fams.view(np.uint8).reshape(-1)).to(dev)
        pair_a = torch.empty(max(1, n_cherries), dtype=torch.int32, device=dev)
        pair_b = torch.empty_like(pair_a)
        len_idx = torch.zeros_like(pair_a)
        unpaired = torch.empty(max(1, n_fams), dtype=torch.int32, device=dev)
        iters = torch.zeros_like(unpaired)
        site_cat = torch.zeros(max(1, total_sites), dtype=torch.int32, device=dev)
        nbytes = int(lib.cherry_fc_scratch_bytes(total_seqs, total_sites, n_fams, K, R, S))
        scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        d_priors = torch.from_numpy(np.ascontiguousarray(priors, dtype=np.float64)).to(dev)
        d_weights = torch.from_numpy(np.ascontiguousarray(weights, dtype=np.float64)).to(dev)
        stream = _lib.current_stream_ptr()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record()
        _lib.check(
            lib.cherry_fc_pair(_lib.ptr(d_msa), _lib.ptr(d_fams), n_fams, total_seqs, S, seed & 0xFFFFFFFF,
                               _lib.ptr(pair_a), _lib.ptr(pair_b), _lib.ptr(unpaired), _lib.ptr(scratch), nbytes,
                               stream),
            "cherry_fc_pair",
        )
        ev[1].record()
        _lib.check(
            lib.cherry_fc_ble(_lib.ptr(d_msa), _lib.ptr(d_fams), n_fams, total_sites, S, _lib.ptr(pair_a),
                              _lib.ptr(pair_b), _lib.ptr(sym_table), K, R, _lib.ptr(d_priors), _lib.ptr(d_weights),
                              int(max_iters), _lib.ptr(len_idx), _lib.ptr(site_cat), _lib.ptr(iters),
                              _lib.ptr(scratch), nbytes, stream),
            "cherry_fc_ble",
        )
        ev[2].record()
        torch.cuda.synchronize()
    extra = {}
    if keep_device:  # for the in-memory hand-off to the counting kernels (_pipeline.py)
        extra = {"d_msa": d_msa, "d_fams": d_fams, "d_pair_a": pair_a, "d_pair_b": pair_b}
    return {
        **extra,
        "pair_a": pair_a[:n_cherries].cpu().numpy(),
        "pair_b": pair_b[:n_cherries].cpu().numpy(),
        "unpaired": unpaired[:n_fams].cpu().numpy(),
        "len_idx": len_idx[:n_cherries].cpu().numpy(),
        "site_cat": site_cat[:total_sites].cpu().numpy(),
        "iters": iters[:n_fams].cpu().numpy(),
        "pair_ms": ev[0].elapsed_time(ev[1]),
        "ble_ms": ev[1].elapsed_time(ev[2]),
    }


def normalise_lengths_and_rates(len_idx: np.ndarray, site_cat: np.ndarray, grid: np.ndarray, cats: np.ndarray):
    """fast_cherries.cpp:268-279: site rates are rescaled to mean 1 (the mean is a left-to-right
    fp64 sum) and the branch lengths absorb the factor."""
    rates = cats[site_cat]
    if len(rates) == 0:
        return grid[len_idx], rates
    mean = float(np.cumsum(rates)[-1]) / len(rates)
    return grid[len_idx] * mean, rates / mean


def _fixed17(x: float) -> str:
    return "%.17f" % x  # std::fixed << std::setprecision(max_digits10)


def cherries_tree(names: Sequence[str], pairs: Sequence[Tuple[int, int]], lengths: Sequence[float],
                  unpaired: int) -> Tree:
    """The star-of-cherries tree of _fast_cherries.py:121-141.  The reference round-trips the
    lengths through the program's '%.17f' text output and builds the tree with ete3, whose new
    nodes hang at distance 1.0 from their parent."""
    tree = Tree()
    tree.add_node("root")
    for i, ((a, b), length) in enumerate(zip(pairs, lengths)):
        half = float(_fixed17(length)) / 2.0
        inner = "internal-" + str(i)
        tree.add_node(inner)
        tree.add_edge("root", inner, 1.0)
        for leaf in (names[a], names[b]):
            tree.add_node(leaf)
            tree.add_edge(inner, leaf, half)
    if unpaired >= 0:
        tree.add_node(names[unpaired])
        tree.add_edge("root", names[unpaired], 1.0)
    return tree


def _newick(names, pairs, lengths, unpaired) -> str:
    parts = []
    for i, ((a, b), length) in enumerate(zip(pairs, lengths)):
        half = float(_fixed17(length)) / 2.0
        parts.append("(%s:%g,%s:%g)internal-%d:1" % (names[a], half, names[b], half, i))
    if unpaired >= 0:
        parts.append("%s:1" % names[unpaired])
    return "(" + ",".join(parts) + ");"


def _c_paths(paths: Sequence[Optional[str]]):
    arr = (ctypes.c_char_p * len(paths))()
    arr[:] = [None if p is None else p.encode() for p in paths]
    return arr


class NativeMsas:
    """MSA files read and encoded by the library's host threads (``cherry_fc_read_msas``); the
    residue buffer is page-locked.  Use as a context manager."""

    def __init__(self, msa_paths: Sequence[str], alphabet: Sequence[str], n_threads: Optional[int] = None,
                 pinned: bool = True) -> None:
        lib = _lib.load()
        self._lib = lib
        self.n_threads = n_threads or os.cpu_count() or 1
        handle = ctypes.POINTER(_lib.FcMsas)()
        _lib.check(lib.cherry_fc_read_msas(_c_paths(msa_paths), len(msa_paths), _c_paths(list(alphabet)),
                                           len(alphabet), self.n_threads, int(pinned), ctypes.byref(handle)),
                   "cherry_fc_read_msas")
        self.handle = handle
        r = handle.contents
        self.n_fams = int(r.n_fams)
        self.msa = np.ctypeslib.as_array(ctypes.cast(r.msa, ctypes.POINTER(ctypes.c_uint8)),
                                         shape=(max(1, int(r.msa_bytes)),))[: int(r.msa_bytes)]
        self.fams = np.ctypeslib.as_array(ctypes.cast(r.fams, ctypes.POINTER(ctypes.c_uint8)),
                                          shape=(max(1, self.n_fams) * 32,))[: self.n_fams * 32].view(
            _lib.FC_FAMILY_DTYPE)

    def names(self, family_index: int) -> List[str]:
        r = self.handle.contents
        fam = self.fams[family_index]
        s0, n = int(fam["seq_off"]), int(fam["n_seqs"])
        off = np.ctypeslib.as_array(ctypes.cast(r.name_off, ctypes.POINTER(ctypes.c_int64)),
                                    shape=(int(r.total_seqs) + 1,))
        blob = ctypes.string_at(r.name_blob + int(off[s0]), int(off[s0 + n] - off[s0]))
        base = int(off[s0])
        return [blob[int(off[s0 + i]) - base: int(off[s0 + i + 1]) - base].decode("utf-8", "replace")
                for i in range(n)]

    def write_outputs(self, out: Dict[str, np.ndarray], grid: np.ndarray, cats: np.ndarray, tree_paths,
                      newick_paths, site_rate_paths, likelihood_paths, profiling_paths, profiling: np.ndarray) -> None:
        c = lambda a, dt: np.ascontiguousarray(a, dtype=dt)  # noqa: E731
        pa, pb, un = c(out["pair_a"], np.int32), c(out["pair_b"], np.int32), c(out["unpaired"], np.int32)
        li, sc = c(out["len_idx"], np.int32), c(out["site_cat"], np.int32)
        g, ct, prof = c(grid, np.float64), c(cats, np.float64), c(profiling, np.float64)
        _lib.check(
            self._lib.cherry_fc_write_outputs(
                self.handle, _lib.ptr(pa), _lib.ptr(pb), _lib.ptr(un), _lib.ptr(li), _lib.ptr(sc), _lib.ptr(g),
                len(g), _lib.ptr(ct), len(ct), _c_paths(tree_paths), _c_paths(newick_paths),
                _c_paths(site_rate_paths), _c_paths(likelihood_paths), _c_paths(profiling_paths), _lib.ptr(prof),
                self.n_threads),
            "cherry_fc_write_outputs",
        )

    def close(self) -> None:
        if self.handle:
            self.msa = self.fams = None
            self._lib.cherry_fc_free_msas(self.handle)
            self.handle = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


@cached_parallel_computation(
    parallel_arg="families",
    exclude_args=["num_processes", "device", "process_group"],
    exclude_args_if_default=["_version"],
    output_dirs=[
        "output_tree_dir",
        "output_site_rates_dir",
        "output_likelihood_dir",
    ],
    write_extra_log_files=True,
)
def fast_cherries(
    msa_dir: str,
    families: List[str],
    rate_matrix_path: str,
    num_rate_categories: int,
    max_iters: int,
    num_processes: int,
    _version="2",
    output_tree_dir: Optional[str] = None,
    output_site_rates_dir: Optional[str] = None,
    output_likelihood_dir: Optional[str] = None,
    remake=False,
    quantization_grid_center=0.03,
    quantization_grid_step=1.1,
    quantization_grid_num_steps=64,
    verbose=True,
    seed=1234,
    device: str = "cuda",
    process_group=None,
) -> None:
    """Same contract as the reference stage: per family ``<tree_dir>/<family>.txt`` (tree),
    ``.newick``, ``.profiling``; ``<site_rates_dir>/<family>.txt``; ``<likelihood_dir>/<family>.txt``
    (the constant 0.0).  ``num_processes`` and ``remake`` are accepted and ignored (one GPU
    launch handles every family; there is no binary to rebuild).

    ``process_group`` (a torch.distributed group, one process per GPU): families are independent,
    so rank r takes ``families[r::world]`` -- the striping of the reference's worker processes
    (``get_process_args``, _fast_cherries.py:175-183) -- writes their files, and a barrier makes
    every file exist before any rank returns.  No data-path collective."""
    for d in (output_tree_dir, output_site_rates_dir, output_likelihood_dir):
        os.makedirs(d, exist_ok=True)
    if process_group is not None:
        import torch.distributed as dist

        rank, world = dist.get_rank(process_group), dist.get_world_size(process_group)
        try:
            _fast_cherries_local(msa_dir, families[rank::world], rate_matrix_path, num_rate_categories, max_iters,
                                 output_tree_dir, output_site_rates_dir, output_likelihood_dir,
                                 quantization_grid_center, quantization_grid_step, quantization_grid_num_steps,
                                 seed, device)
        finally:
            dist.barrier(process_group)
        return
    _fast_cherries_local(msa_dir, families, rate_matrix_path, num_rate_categories, max_iters, output_tree_dir,
                         output_site_rates_dir, output_likelihood_dir, quantization_grid_center,
                         quantization_grid_step, quantization_grid_num_steps, seed, device)


# In-process hand-off to count_transitions (reference estimation_end_to_end/_cherry.py:279-336: the tree-free LG
# pipeline estimates cherries and then counts on them): the encoded residues and the FastCherries results of
# the LAST stage call stay on the device, keyed by the directories and families the counting stage will be
# given, so that it does not parse the trees, site rates and MSAs it would otherwise read back from the files
# this stage wrote (they are still written).  One entry; taken (removed) by the first matching call.
_HANDOFF: Dict = {}


def take_handoff(tree_dir, site_rates_dir, msa_dir, families, alphabet):
    """The resident FastCherries results for exactly these directories / families / alphabet, or None."""
    key = (os.path.realpath(tree_dir), os.path.realpath(site_rates_dir), os.path.realpath(msa_dir),
           tuple(families), tuple(alphabet))
    entry = _HANDOFF.pop("entry", None)
    if entry is None or entry["key"] != key:
        return None
    # the files this stage wrote must still be the ones on disk (a later writer wins over the resident copy)
    for path, stamp in entry["stamps"].items():
        try:
            st = os.stat(path)
        except OSError:
            return None
        if (st.st_size, st.st_mtime_ns) != stamp:
            return None
    return entry


def clear_handoff() -> None:
    _HANDOFF.clear()


def _fast_cherries_local(msa_dir, families, rate_matrix_path, num_rate_categories, max_iters, output_tree_dir,
                         output_site_rates_dir, output_likelihood_dir, quantization_grid_center,
                         quantization_grid_step, quantization_grid_num_steps, seed, device) -> None:
    if not families:
        return
    t_start = time.time()
    rate_matrix = read_rate_matrix(rate_matrix_path)
    alphabet = list(rate_matrix.columns)
    Q = rate_matrix.to_numpy(dtype=np.float64)
    # the reference passes the grid parameters through str() on a command line
    grid = quantization_grid(float(str(quantization_grid_center)), float(str(quantization_grid_step)),
                             int(quantization_grid_num_steps))
    cats = ble_rate_categories(num_rate_categories)
    weights = initial_rate_weights(cats)
    priors = np.array([2 * math.log(r) - 3 * r for r in cats])
    table = log_transition_table(Q, grid, cats, device)
    join = lambda d, ext: [os.path.join(d, f + ext) for f in families]  # noqa: E731
    keep = os.environ.get("CHERRY_FC_HANDOFF", "1") != "0"
    with NativeMsas(join(msa_dir, ".txt"), alphabet) as msas:
        out = fast_cherries_device(msas.msa, msas.fams, len(alphabet), table, priors, weights, int(seed),
                                   int(max_iters), device, keep_device=keep)
        n = len(families)
        elapsed = time.time() - t_start
        profiling = np.empty((n, 4))
        profiling[:, 0] = out["pair_ms"] * 1e-3 / n
        profiling[:, 1] = out["ble_ms"] * 1e-3 / n
        profiling[:, 2] = elapsed / n
        profiling[:, 3] = elapsed / n
        msas.write_outputs(out, grid, cats, join(output_tree_dir, ".txt"), join(output_tree_dir, ".newick"),
                           join(output_site_rates_dir, ".txt"), join(output_likelihood_dir, ".txt"),
                           join(output_tree_dir, ".profiling"), profiling)
        _HANDOFF.clear()
        if keep:
            stamps = {}
            for pth in join(output_tree_dir, ".txt") + join(output_site_rates_dir, ".txt"):
                st = os.stat(pth)
                stamps[pth] = (st.st_size, st.st_mtime_ns)
            _HANDOFF["entry"] = dict(
                key=(os.path.realpath(output_tree_dir), os.path.realpath(output_site_rates_dir),
                     os.path.realpath(msa_dir), tuple(families), tuple(alphabet)),
                fams=np.array(msas.fams, copy=True), out=out, grid=grid, cats=cats, stamps=stamps, device=device)
