"""CPU restatement of the reference's FastCherries program (TEST INFRASTRUCTURE).

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module; the
product path (cherryml_b200.phylogeny_estimation) never does.

Reference files followed (paths relative to cherryml/phylogeny_estimation/FastCherries/):

* pairing: ``pairing_algorithms.cpp:15-175`` (negated normalised Hamming distance,
  ``find_farthest``, ``partition_subset_distance``, the recursive ``divide``);
* random pivots: ``std::mt19937 rng(seed)`` re-seeded per family (``fast_cherries.cpp:226``)
  drawn through ``std::uniform_int_distribution<size_t>`` (``pairing_algorithms.cpp:97-98``).
  The distribution's algorithm is libstdc++'s (GCC >= 11, ``bits/uniform_int_dist.h``):
  Lemire's nearly-divisionless method on the 32-bit engine output;
* branch lengths and site rates: ``branch_length_estimation.cpp:10-241`` (gamma-bin initial
  site rates, binary searches over the quantization grid / the rate categories with
  sequential fp64 sums in site / cherry order, coordinate ascent until the lengths repeat);
* grid, rate categories, prior weights: ``io_helpers.cpp:178-194``, ``fast_cherries.cpp:47-160,
  205-215`` (the incomplete-gamma routine there is Bhattacharjee's AS 32 as distributed with
  FastTree 2.1);
* log transition table: ``io_helpers.cpp:150-176``.  The reference exponentiates with a
  third-party routine (John Burkardt's ``r8mat_expm1``: Pade(6) + scaling and squaring, vendored
  under FastCherries/matrix_exponential/).  This restatement takes the table as an INPUT
  (``log_table``) so that the CUDA path and the oracle can be compared on identical tables;
  ``log_table_scipy`` builds one with scipy.linalg.expm, which agrees with the reference's to
  ~1e-15 relative -- decisions then differ only on ties closer than that, none on the goldens.

PINNED against: the KATs of the reference's own C++ tests (FastCherries/tests/
test_branch_length_estimation.cpp, test_pairing_algorithms.cpp -- restated in
tests/test_oracle_fast_cherries.py) and outputs of the UNMODIFIED reference program compiled
into oracle/_ref/fast_cherries (tests/golden/fast_cherries, made by
tests/golden/make_golden_fast_cherries.py).
"""
import math
import os
import subprocess
import tempfile
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_BIN = os.path.join(HERE, "_ref", "fast_cherries")


# ----------------------------------------------------------------------------- RNG

class MT19937:
    """std::mt19937 (32-bit Mersenne twister, standard initialisation by a 32-bit seed)."""

    def __init__(self, seed: int) -> None:
        mt = [0] * 624
        mt[0] = seed & 0xFFFFFFFF
        for i in range(1, 624):
            mt[i] = (1812433253 * (mt[i - 1] ^ (mt[i - 1] >> 30)) + i) & 0xFFFFFFFF
        self.mt = mt
        self.pos = 624

    def _twist(self) -> None:
        mt = self.mt
        for i in range(624):
            y = (mt[i] & 0x80000000) | (mt[(i + 1) % 624] & 0x7FFFFFFF)
            v = mt[(i + 397) % 624] ^ (y >> 1)
            if y & 1:
                v ^= 0x9908B0DF
            mt[i] = v
        self.pos = 0

    def __call__(self) -> int:
        if self.pos >= 624:
            self._twist()
        y = self.mt[self.pos]
        self.pos += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9D2C5680
        y ^= (y << 15) & 0xEFC60000
        y ^= y >> 18
        return y & 0xFFFFFFFF


def uniform_index(rng: MT19937, n: int) -> int:
    """std::uniform_int_distribution<size_t>(0, n-1)(rng) as libstdc++ (GCC >= 11) computes it."""
    rng_range = n  # __uerange
    product = rng() * rng_range
    low = product & 0xFFFFFFFF
    if low < rng_range:
        threshold = ((1 << 32) - rng_range) % rng_range
        while low < threshold:
            product = rng() * rng_range
            low = product & 0xFFFFFFFF
    return product >> 32


# ----------------------------------------------------------------------------- pairing

def _neg_hamming(rows: np.ndarray, pivot: np.ndarray) -> np.ndarray:
    """pairing_algorithms.cpp:15-40 for every row: -(#differing)/(#both valid), 0 if none valid."""
    valid = (rows >= 0) & (pivot >= 0)[None, :]
    count = valid.sum(axis=1)
    dist = (valid & (rows != pivot[None, :])).sum(axis=1)
    out = np.zeros(rows.shape[0], dtype=np.float64)
    nz = count > 0
    out[nz] = (dist[nz] * -1.0) / count[nz]
    return out


def divide_and_pair(seqs: np.ndarray, seed: int) -> List[Tuple[int, int]]:
    """Row-index pairs in the order the reference emits them (pairing_algorithms.cpp:79-175).
    seqs: int array [N, L], -1 = not in the alphabet."""
    rng = MT19937(seed)
    cherries: List[Tuple[int, int]] = []

    def first_argmin(d: np.ndarray) -> int:
        return int(np.argmin(d))  # first occurrence == the strict '<' scan of find_farthest

    def divide(lst: List[int]) -> int:
        n = len(lst)
        if n == 2:
            cherries.append((lst[0], lst[1]))
            return -1
        if n == 1:
            return lst[0]
        if n == 0:
            return -1
        x = lst[uniform_index(rng, n)]
        rows = seqs[lst]
        x = lst[first_argmin(_neg_hamming(rows, seqs[x]))]
        dist_x = _neg_hamming(rows, seqs[x])
        y = lst[first_argmin(dist_x)]
        closer_left = dist_x >= _neg_hamming(rows, seqs[y])
        close_x = [v for v, c in zip(lst, closer_left) if c and v != y]
        close_y = [v for v, c in zip(lst, closer_left) if not (c and v != y)]
        ux = divide(close_x)
        uy = divide(close_y)
        if ux >= 0 and uy >= 0:
            cherries.append((ux, uy))
            return -1
        return ux if ux >= 0 else uy

    import sys

    old = sys.getrecursionlimit()
    sys.setrecursionlimit(max(old, 4 * seqs.shape[0] + 100))
    try:
        divide(list(range(seqs.shape[0])))
    finally:
        sys.setrecursionlimit(old)
    return cherries


# ----------------------------------------------------------------------------- setup

def quantization_points(center: float, step: float, num_steps: int) -> np.ndarray:
    """io_helpers.cpp:178-194: the chain is evaluated in long double, then narrowed."""
    q = np.zeros(2 * num_steps + 1, dtype=np.longdouble)
    q[num_steps] = np.longdouble(center)
    s = np.longdouble(step)
    for i in range(1, num_steps + 1):
        q[num_steps + i] = q[num_steps + i - 1] * s
        q[num_steps - i] = q[num_steps - i + 1] / s
    return q.astype(np.float64)


def rate_categories(R: int) -> np.ndarray:
    """fast_cherries.cpp:205-213."""
    if R == 1:
        return np.array([1.0])
    start = 1.0 / R
    ratio = math.pow(R / start, 1.0 / (R - 1))
    out = [start]
    for _ in range(1, R):
        out.append(out[-1] * ratio)
    return np.array(out, dtype=np.float64)


def _ln_gamma(alpha: float) -> float:
    """Pike & Hill (1966) Algorithm 291 as used at fast_cherries.cpp:47-66."""
    x, f = alpha, 0.0
    if x < 7:
        f = 1.0
        z = x - 1.0
        while True:
            z += 1.0
            if not z < 7:
                break
            f *= z
        x = z
        f = -math.log(f)
    z = 1.0 / (x * x)
    return (f + (x - 0.5) * math.log(x) - x + .918938533204673
            + (((-.000595238095238 * z + .000793650793651) * z - .002777777777778) * z + .083333333333333) / x)


def _incomplete_gamma(x: float, alpha: float, ln_gamma_alpha: float) -> float:
    """Bhattacharjee (1970) AS 32 with the constants of fast_cherries.cpp:69-128."""
    p, g = alpha, ln_gamma_alpha
    accurate, overflow = 1e-8, 1e30
    if x == 0:
        return 0.0
    if x < 0 or p <= 0:
        return -1.0
    factor = math.exp(p * math.log(x) - x - g)
    if not (x > 1 and x >= p):
        gin, term, rn = 1.0, 1.0, p
        while True:
            rn += 1
            term *= x / rn
            gin += term
            if not term > accurate:
                break
        return gin * (factor / p)
    a = 1 - p
    b = a + x + 1
    term = 0.0
    pn = [1.0, x, x + 1, x * b, 0.0, 0.0]
    gin = pn[2] / pn[3]
    while True:
        a += 1
        b += 2
        term += 1
        an = a * term
        for i in range(2):
            pn[i + 4] = b * pn[i + 2] - an * pn[i]
        if pn[5] != 0:
            rn = pn[4] / pn[5]
            dif = abs(gin - rn)
            if not dif > accurate and dif <= accurate * rn:
                return 1 - factor * gin
            gin = rn
        for i in range(4):
            pn[i] = pn[i + 2]
        if abs(pn[4]) >= overflow:
            for i in range(4):
                pn[i] /= overflow


def initial_site_rate_weights(cats: np.ndarray) -> np.ndarray:
    """fast_cherries.cpp:137-160: gamma(shape 3, scale 1/3) CDF at the geometric midpoints."""
    shape = 3.0
    w = [
        _incomplete_gamma(math.sqrt(cats[i - 1] * cats[i]) * shape, shape, _ln_gamma(shape))
        for i in range(1, len(cats))
    ]
    w.append(1.0)
    return np.array(w, dtype=np.float64)


def log_table_scipy(Q: np.ndarray, q: np.ndarray, cats: np.ndarray) -> np.ndarray:
    """[K, R, S, S] log expm(q_k * rate_r * Q) (io_helpers.cpp:150-176) with scipy's expm."""
    from scipy.linalg import expm

    out = np.zeros((len(q), len(cats), Q.shape[0], Q.shape[0]))
    for i, qi in enumerate(q):
        for r, rr in enumerate(cats):
            with np.errstate(divide="ignore"):
                out[i, r] = np.log(expm(qi * rr * Q))
    return out


# ----------------------------------------------------------------------------- BLE

def initial_site_categories(seqs: np.ndarray, weights: np.ndarray, S: int) -> np.ndarray:
    """branch_length_estimation.cpp:10-62 (all sequences, not only the paired ones)."""
    L = seqs.shape[1]
    counts = np.zeros((L, S), dtype=np.int64)
    for k in range(S):
        counts[:, k] = (seqs == k).sum(axis=0)
    non_missing = counts.sum(axis=1)
    total = ((non_missing[:, None] - counts) * counts).sum(axis=1)
    order = sorted(range(L), key=lambda j: (int(total[j]), j))
    w = [float(int(round_half_away(wr * L))) for wr in weights]
    out = np.zeros(L, dtype=np.int64)
    rc = 0
    for i in range(L):
        rc += i >= w[rc]
        out[order[i]] = rc
    return out


def round_half_away(x: float) -> float:
    return math.floor(x + 0.5) if x >= 0 else math.ceil(x - 0.5)


def _seq_sum(start: float, vals: np.ndarray) -> float:
    """start + vals[0] + vals[1] + ... strictly left to right (np.cumsum is sequential)."""
    if vals.size == 0:
        return start
    return float(np.cumsum(np.concatenate(([start], vals)))[-1])


def branch_length_indices(xa: np.ndarray, xb: np.ndarray, sym: np.ndarray, site_cat: np.ndarray) -> np.ndarray:
    """get_branch_lengths, branch_length_estimation.cpp:64-108.  sym[k, r, x, y] = T + T^T."""
    K = sym.shape[0]
    out = np.zeros(xa.shape[0], dtype=np.int64)
    for c in range(xa.shape[0]):
        v = np.nonzero((xa[c] >= 0) & (xb[c] >= 0))[0]
        x, y, r = xa[c, v], xb[c, v], site_cat[v]
        low, high = 0, K - 1
        while low < high:
            mid = low + (high - low) // 2
            if _seq_sum(0.0, sym[mid, r, x, y]) > _seq_sum(0.0, sym[mid + 1, r, x, y]):
                high = mid
            else:
                low = mid + 1
        out[c] = low
    return out


def site_rate_indices(xa: np.ndarray, xb: np.ndarray, sym: np.ndarray, len_idx: np.ndarray,
                      priors: np.ndarray) -> np.ndarray:
    """get_site_rates, branch_length_estimation.cpp:110-148."""
    R = len(priors)
    L = xa.shape[1]
    out = np.zeros(L, dtype=np.int64)
    for j in range(L):
        v = np.nonzero((xa[:, j] >= 0) & (xb[:, j] >= 0))[0]
        x, y, k = xa[v, j], xb[v, j], len_idx[v]
        low, high = 0, R - 1
        while low < high:
            mid = low + (high - low) // 2
            if _seq_sum(priors[mid], sym[k, mid, x, y]) > _seq_sum(priors[mid + 1], sym[k, mid + 1, x, y]):
                high = mid
            else:
                low = mid + 1
        out[j] = low
    return out


def ble(seqs: np.ndarray, cherries: Sequence[Tuple[int, int]], log_table: np.ndarray, cats: np.ndarray,
        weights: np.ndarray, max_iters: int) -> Tuple[np.ndarray, np.ndarray, int]:
    """ble(), branch_length_estimation.cpp:150-241 -> (length index per cherry, category per site,
    coordinate-ascent iterations run)."""
    S = log_table.shape[2]
    sym = log_table + np.swapaxes(log_table, 2, 3)  # a + b of the inner loops, commutative in fp64
    a = np.array([c[0] for c in cherries], dtype=np.int64)
    b = np.array([c[1] for c in cherries], dtype=np.int64)
    xa, xb = seqs[a], seqs[b]
    site_cat = initial_site_categories(seqs, weights, S)
    len_idx = branch_length_indices(xa, xb, sym, site_cat)
    priors = np.array([2 * math.log(r) - 3 * r for r in cats])
    match, iters = False, 0
    while not match and max_iters:
        site_cat = site_rate_indices(xa, xb, sym, len_idx, priors)
        new_len = branch_length_indices(xa, xb, sym, site_cat)
        match = bool(np.array_equal(new_len, len_idx))
        len_idx = new_len
        max_iters -= 1
        iters += 1
    return len_idx, site_cat, iters


def finalize(len_idx: np.ndarray, site_cat: np.ndarray, q: np.ndarray, cats: np.ndarray):
    """fast_cherries.cpp:258-272: rates normalised to mean 1 (sequential sum), lengths scaled up."""
    rates = cats[site_cat]
    mean = float(np.cumsum(rates)[-1]) / len(rates)
    return q[len_idx] * mean, rates / mean


def encode(seqs: Sequence[str], alphabet: Sequence[str]) -> np.ndarray:
    lut = np.full(256, -1, dtype=np.int64)
    for i, ch in enumerate(alphabet):
        lut[ord(ch)] = i
    return np.stack([lut[np.frombuffer(s.encode("latin-1"), dtype=np.uint8)] for s in seqs])


def fast_cherries_oracle(seqs: np.ndarray, log_table: np.ndarray, q: np.ndarray, cats: np.ndarray, seed: int,
                         max_iters: int):
    """One family end to end -> (cherries, length per cherry, rate per site, len_idx, site_cat)."""
    cherries = divide_and_pair(seqs, seed)
    weights = initial_site_rate_weights(cats)
    len_idx, site_cat, _ = ble(seqs, cherries, log_table, cats, weights, max_iters)
    lengths, rates = finalize(len_idx, site_cat, q, cats)
    return cherries, lengths, rates, len_idx, site_cat


# ----------------------------------------------------------------------------- reference binary

def have_reference_binary() -> bool:
    return os.path.exists(REF_BIN)


def _write_list(paths: Sequence[str], fn: str) -> None:
    with open(fn, "w") as f:
        f.write(str(len(paths)) + "\n" + "\n".join(paths))


def run_reference_binary(msa_paths: Sequence[str], rate_matrix_path: str, num_rate_categories: int,
                         max_iters: int = 50, seed: int = 1234, center: float = 0.03, step: float = 1.1,
                         num_steps: int = 64, out_dir: Optional[str] = None) -> List[Dict]:
    """Run oracle/_ref/fast_cherries the way the reference's Python wrapper does
    (phylogeny_estimation/_fast_cherries.py:85-104, 229-236).  rate_matrix_path: labelled table."""
    own = out_dir is None
    tmp = tempfile.mkdtemp() if own else out_dir
    lines = open(rate_matrix_path).read().strip().split("\n")
    alphabet = lines[0].split()
    with open(os.path.join(tmp, "Q.txt"), "w") as f:
        f.writelines([ln[1:] + "\n" for ln in lines[1:]])
    with open(os.path.join(tmp, "alphabet.txt"), "w") as f:
        f.write(str(len(alphabet)) + " " + " ".join(alphabet))
    n = len(msa_paths)
    outs = [os.path.join(tmp, f"{i}.output") for i in range(n)]
    profs = [os.path.join(tmp, f"{i}.profiling") for i in range(n)]
    rates = [os.path.join(tmp, f"{i}.rates") for i in range(n)]
    _write_list(list(msa_paths), os.path.join(tmp, "msas.txt"))
    _write_list(outs, os.path.join(tmp, "outs.txt"))
    _write_list(profs, os.path.join(tmp, "profs.txt"))
    _write_list(rates, os.path.join(tmp, "rates.txt"))
    cmd = [
        REF_BIN, "-seed", str(seed), "-quantization_grid_center", str(center), "-quantization_grid_step",
        str(step), "-quantization_grid_num_steps", str(num_steps), "-output_list_path",
        os.path.join(tmp, "outs.txt"), "-rate_matrix_path", os.path.join(tmp, "Q.txt"), "-msa_list_path",
        os.path.join(tmp, "msas.txt"), "-profiling_list_path", os.path.join(tmp, "profs.txt"),
        "-site_rate_list_path", os.path.join(tmp, "rates.txt"), "-num_rate_categories_ble",
        str(num_rate_categories), "-max_iters_ble", str(max_iters), "-alphabet_path",
        os.path.join(tmp, "alphabet.txt"),
    ]
    subprocess.run(cmd, check=True)
    res = []
    for i in range(n):
        toks = open(outs[i]).read().split("\n")
        toks = toks[: len(toks) - 1] if toks and toks[-1] == "" else toks
        cherries = [(toks[j], toks[j + 1]) for j in range(0, len(toks), 3)]
        dist_text = [toks[j + 2] for j in range(0, len(toks), 3)]
        rl = open(rates[i]).read().split("\n")
        res.append({
            "cherries": cherries,
            "distances_text": dist_text,
            "site_rates_text": rl[1].split() if len(rl) > 1 else [],
            "output_text": open(outs[i]).read(),
            "site_rates_file_text": open(rates[i]).read(),
        })
    if own:
        import shutil

        shutil.rmtree(tmp, ignore_errors=True)
    return res
