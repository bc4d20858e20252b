"""Synthetic Pfam-shaped families, generated directly in the encoded layout.

Used by bench.py (BASELINE.json configs 3 and 4) and by the parity tests.  Statistics follow
SURVEY.md section 8(d): residues uniform over the 20 amino acids with 13 % gaps per
position, the second leaf of a cherry is the first with each site resampled with
probability 0.34, cherry lengths t ~ LogNormal(ln 0.52, 0.95), site rates in R equally
sized gamma(shape=1) quantile categories normalised to mean 1, star-shaped trees (every
family is N/2 cherries: rows 2i and 2i+1).  Families all have the same shape, which lets
the whole batch be generated with a few tensor ops on the target device.

Also renders a slice of a batch as the reference's text files so that the reference
programs can consume the identical data (``write_text_rendering``).
"""
import math
import os
from typing import List, Optional, Sequence

import numpy as np
import torch

from ._lib import FAM_DESC_DTYPE, TILE_DTYPE

SKIP = 20  # skip code of a residue byte == number of states (20 amino acids)
from .counting._ingest import TARGET_CHUNKS_PER_TILE, TARGET_ITEMS_PER_CO_TILE, CountBatch
from .utils import amino_acids


def gamma_category_rates(n_cats: int) -> np.ndarray:
    """Means of the n equal-probability slices of Exp(1) (= gamma shape 1), mean 1."""
    # slice i covers quantiles [i/n, (i+1)/n); its mean has a closed form for Exp(1)
    with np.errstate(divide="ignore"):
        edges = -np.log1p(-np.arange(n_cats + 1) / n_cats)  # last edge = inf
    means = np.empty(n_cats)
    for i in range(n_cats):
        a, b = edges[i], edges[i + 1]
        ea = math.exp(-a) * (a + 1)
        eb = 0.0 if math.isinf(b) else math.exp(-b) * (b + 1)
        means[i] = (ea - eb) * n_cats
    return means / means.mean()


def quantization_grid(center: float = 0.03, step: float = 1.1, lo: int = -50, hi: int = 49) -> List[float]:
    """K = hi-lo+1 points ``float("%.8f" % (center*step**i))`` (default K = 100)."""
    return [float("%.8f" % (center * step**i)) for i in range(lo, hi + 1)]


def _lg_layout(n_sites: int, n_cats: int):
    counts = np.array([n_sites // n_cats + (1 if c < n_sites % n_cats else 0) for c in range(n_cats)])
    padded = (counts + 3) // 4 * 4
    starts = np.concatenate([[0], np.cumsum(padded)[:-1]])
    total = int(padded.sum())
    stride = max(16, (total + 15) // 16 * 16)
    col_of_site = np.concatenate([starts[c] + np.arange(counts[c]) for c in range(n_cats)])
    cat_of_site = np.repeat(np.arange(n_cats), counts)
    group_cat = np.zeros(stride // 4, dtype=np.uint16)
    group_cat[: total // 4] = np.repeat(np.arange(n_cats, dtype=np.uint16), padded // 4)
    return stride, col_of_site, cat_of_site, group_cat


def _residue_rows(n_fams, n_seqs, n_sites, stride, cols, gen, device, gap_frac, mut_frac, fam_chunk=128):
    """uint8 [n_fams * n_seqs * stride] with rows (2i, 2i+1) forming cherries."""
    n_pairs = n_seqs // 2
    out = torch.full((n_fams * n_seqs * stride,), SKIP, dtype=torch.uint8, device=device)
    view = out.view(n_fams, n_pairs, 2, stride)
    cols_t = torch.as_tensor(cols, device=device, dtype=torch.long)
    for f0 in range(0, n_fams, fam_chunk):
        f1 = min(n_fams, f0 + fam_chunk)
        shape = (f1 - f0, n_pairs, n_sites)
        a = torch.randint(0, 20, shape, dtype=torch.uint8, device=device, generator=gen)
        fresh = torch.randint(0, 20, shape, dtype=torch.uint8, device=device, generator=gen)
        mutate = torch.rand(shape, device=device, generator=gen) < mut_frac
        b = torch.where(mutate, fresh, a)
        gap = torch.tensor(SKIP, dtype=torch.uint8, device=device)
        a = torch.where(torch.rand(shape, device=device, generator=gen) < gap_frac, gap, a)
        b = torch.where(torch.rand(shape, device=device, generator=gen) < gap_frac, gap, b)
        view[f0:f1, :, 0, :].index_copy_(2, cols_t, a)
        view[f0:f1, :, 1, :].index_copy_(2, cols_t, b)
    return out


def _pair_arrays(n_fams, n_pairs, gen, device):
    idx = torch.arange(n_pairs, dtype=torch.int32, device=device).repeat(n_fams)
    pair_a, pair_b = idx * 2, idx * 2 + 1
    pair_fam = torch.arange(n_fams, dtype=torch.int32, device=device).repeat_interleave(n_pairs)
    z = torch.randn(n_fams * n_pairs, dtype=torch.float64, device=device, generator=gen)
    pair_t = torch.exp(math.log(0.52) + 0.95 * z)
    return pair_a, pair_b, pair_t, pair_fam


def _tiles(n_fams, n_pairs, per_tile) -> np.ndarray:
    per_fam = [(b, min(per_tile, n_pairs - b)) for b in range(0, n_pairs, per_tile)]
    tiles = np.zeros(n_fams * len(per_fam), dtype=TILE_DTYPE)
    fam = np.repeat(np.arange(n_fams, dtype=np.int64), len(per_fam))
    begin = np.tile(np.array([b for b, _ in per_fam], dtype=np.int64), n_fams)
    tiles["fam"] = fam
    tiles["pair_begin"] = fam * n_pairs + begin
    tiles["n_pairs"] = np.tile(np.array([n for _, n in per_fam], dtype=np.int32), n_fams)
    return tiles


def synthetic_lg(
    n_fams: int, n_seqs: int, n_sites: int, n_rate_cats: int = 4, seed: int = 0,
    device="cpu", gap_frac: float = 0.13, mut_frac: float = 0.34,
):
    """Returns a dict of torch tensors on ``device`` plus host descriptors; see
    ``as_count_batch`` / ``as_device_batch``."""
    device = torch.device(device)
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    n_pairs = n_seqs // 2
    stride, cols, _, group_cat = _lg_layout(n_sites, n_rate_cats)
    msa = _residue_rows(n_fams, n_seqs, n_sites, stride, cols, gen, device, gap_frac, mut_frac)
    pair_a, pair_b, pair_t, pair_fam = _pair_arrays(n_fams, n_pairs, gen, device)
    fams = np.zeros(n_fams, dtype=FAM_DESC_DTYPE)
    f = np.arange(n_fams, dtype=np.int64)
    fams["msa_off"] = f * n_seqs * stride
    fams["row_stride"] = stride
    fams["n_chunks"] = stride // 16
    fams["aux_off"] = f * (stride // 4)
    fams["aux_cnt"] = stride // 4
    fams["rate_off"] = f * n_rate_cats
    fams["n_rates"] = n_rate_cats
    rates = gamma_category_rates(n_rate_cats)
    per_tile = max(1, TARGET_CHUNKS_PER_TILE // (stride // 16))
    return dict(
        kind="lg", msa=msa, fams=fams, pair_a=pair_a, pair_b=pair_b, pair_t=pair_t, pair_fam=pair_fam,
        rate_vals=np.tile(rates, n_fams), aux=np.tile(group_cat, n_fams),
        tiles=_tiles(n_fams, n_pairs, per_tile), r_pad=(n_rate_cats + 3) // 4 * 4,
        n_sites_examined=n_fams * n_pairs * n_sites,
        shape=dict(n_fams=n_fams, n_seqs=n_seqs, n_sites=n_sites, n_rate_cats=n_rate_cats, stride=stride),
    )


def synthetic_co(
    n_fams: int, n_seqs: int, n_sites: int, seed: int = 0, device="cpu",
    gap_frac: float = 0.13, mut_frac: float = 0.34, min_dist: int = 7,
):
    """Co-transition batch: each family's contact list is a random matching of the sites
    (what maximal matching yields) with pairs closer than ``min_dist`` dropped."""
    device = torch.device(device)
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    n_pairs = n_seqs // 2
    rng = np.random.default_rng(seed)
    perm = np.argsort(rng.random((n_fams, n_sites)), axis=1)
    half = n_sites // 2
    i = np.minimum(perm[:, 0 : 2 * half : 2], perm[:, 1 : 2 * half : 2])
    j = np.maximum(perm[:, 0 : 2 * half : 2], perm[:, 1 : 2 * half : 2])
    keep = (j - i) >= min_dist
    cnt = keep.sum(axis=1)
    contacts = np.stack([i[keep], j[keep]], axis=1).astype(np.int32)  # row-major == family order
    # Rows are contact-paired (bytes 2c, 2c+1 = the two sites of contact c).  Residues are
    # i.i.d. over sites, so the rows are drawn directly in that layout; bytes past a
    # family's 2*cnt are the skip code.
    n_cols = 2 * int(cnt.max()) if n_fams else 0
    stride = max(16, (n_cols + 15) // 16 * 16)
    msa = _residue_rows(n_fams, n_seqs, n_cols, stride, np.arange(n_cols), gen, device, gap_frac, mut_frac)
    live = torch.arange(stride, device=device)[None, :] < torch.as_tensor(2 * cnt, device=device)[:, None]
    msa.view(n_fams, n_seqs, stride).masked_fill_(~live[:, None, :], SKIP)
    pair_a, pair_b, pair_t, pair_fam = _pair_arrays(n_fams, n_pairs, gen, device)
    fams = np.zeros(n_fams, dtype=FAM_DESC_DTYPE)
    f = np.arange(n_fams, dtype=np.int64)
    fams["msa_off"] = f * n_seqs * stride
    fams["row_stride"] = stride
    fams["n_chunks"] = stride // 16
    fams["aux_off"] = np.concatenate([[0], np.cumsum(cnt)[:-1]])
    fams["aux_cnt"] = cnt
    fams["rate_off"] = f
    fams["n_rates"] = 1
    per_tile = max(1, TARGET_ITEMS_PER_CO_TILE // max(1, half))
    return dict(
        kind="co", msa=msa, fams=fams, pair_a=pair_a, pair_b=pair_b, pair_t=pair_t, pair_fam=pair_fam,
        rate_vals=np.ones(n_fams), aux=contacts, tiles=_tiles(n_fams, n_pairs, per_tile), r_pad=4,
        n_sites_examined=int(n_pairs * cnt.sum()),
        shape=dict(n_fams=n_fams, n_seqs=n_seqs, n_sites=n_sites, stride=stride),
    )


def as_count_batch(syn: dict) -> CountBatch:
    """Host (numpy) view of a synthetic batch, e.g. for the oracle or the host-buffer path."""
    def host(x):
        return x.cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)

    return CountBatch(
        kind=syn["kind"], msa=host(syn["msa"]), fams=syn["fams"], pair_a=host(syn["pair_a"]),
        pair_b=host(syn["pair_b"]), pair_t=host(syn["pair_t"]), pair_fam=host(syn["pair_fam"]),
        rate_vals=np.asarray(syn["rate_vals"], dtype=np.float64), aux=syn["aux"], tiles=syn["tiles"],
        r_pad=syn["r_pad"], n_sites_examined=syn["n_sites_examined"],
    )


def as_device_batch(syn: dict, device="cuda"):
    """DeviceBatch without a round trip through the host for the big residue buffer."""
    from .counting._device import DeviceBatch, _as_device

    device = torch.device(device)

    def dev(x):
        if isinstance(x, torch.Tensor):
            return x.to(device).contiguous().view(torch.uint8).reshape(-1)
        return _as_device(np.asarray(x), device)

    tile_items = None
    if syn["kind"] == "co":
        tile_items = syn["tiles"]["n_pairs"].astype(np.int64) * syn["fams"]["aux_cnt"][syn["tiles"]["fam"]]
    return DeviceBatch(
        kind=syn["kind"], msa=dev(syn["msa"]), fams=dev(syn["fams"]), pair_a=dev(syn["pair_a"]),
        pair_b=dev(syn["pair_b"]), pair_t=dev(syn["pair_t"]), pair_fam=dev(syn["pair_fam"]),
        rate_vals=dev(np.asarray(syn["rate_vals"], dtype=np.float64)), aux=dev(syn["aux"]),
        tiles=dev(syn["tiles"]), r_pad=syn["r_pad"], n_pairs=int(syn["pair_a"].shape[0]),
        n_tiles=int(syn["tiles"].shape[0]), n_sites_examined=syn["n_sites_examined"], tile_items=tile_items,
        max_row_stride=int(syn["fams"]["row_stride"].max()) if len(syn["fams"]) else 16,
    )


def write_text_rendering(
    syn: dict, out_dir: str, families: Optional[Sequence[int]] = None, states: Sequence[str] = amino_acids,
) -> List[str]:
    """Write families of a synthetic batch as the reference's text files
    (``tree_dir``, ``msa_dir``, ``site_rates_dir`` or ``contact_map_dir`` under ``out_dir``).

    Trees are FastCherries-shaped stars: root -> internal_i -> two leaves, each leaf edge
    t/2 (reference ``phylogeny_estimation/_fast_cherries.py:120-131``).  Note that the
    C++ reference reads branch lengths as float32, so renderings meant for it should be
    counted with ``float32_branch_lengths`` semantics (t is re-derived from the text)."""
    batch = as_count_batch(syn)
    shape = syn["shape"]
    n_seqs, n_sites, stride = shape["n_seqs"], shape["n_sites"], shape["stride"]
    n_pairs = n_seqs // 2
    fam_ids = list(range(len(batch.fams))) if families is None else list(families)
    letters = np.frombuffer(("".join(states)).encode("ascii"), dtype=np.uint8)
    table = np.full(256, ord("-"), dtype=np.uint8)
    table[: len(letters)] = letters
    third = "site_rates_dir" if batch.kind == "lg" else "contact_map_dir"
    for sub in ("tree_dir", "msa_dir", third):
        os.makedirs(os.path.join(out_dir, sub), exist_ok=True)
    if batch.kind == "lg":
        _, cols, cat_of_site, _ = _lg_layout(n_sites, shape["n_rate_cats"])
    names = []
    for f in fam_ids:
        fd = batch.fams[f]
        name = f"fam{f:06d}"
        names.append(name)
        rows = batch.msa[fd["msa_off"] : fd["msa_off"] + n_seqs * stride].reshape(n_seqs, stride)
        if batch.kind == "lg":
            text = table[rows[:, cols]]
        else:
            # contact-paired rows -> site order; sites in no contact are never read: gaps
            cs = batch.aux[fd["aux_off"] : fd["aux_off"] + fd["aux_cnt"]]
            text = np.full((n_seqs, n_sites), ord("-"), dtype=np.uint8)
            text[:, cs[:, 0]] = table[rows[:, 0 : 2 * len(cs) : 2]]
            text[:, cs[:, 1]] = table[rows[:, 1 : 2 * len(cs) : 2]]
        with open(os.path.join(out_dir, "msa_dir", name + ".txt"), "w") as fh:
            fh.write("".join(f">seq{r}\n{text[r].tobytes().decode('ascii')}\n" for r in range(n_seqs)))
        t = batch.pair_t[f * n_pairs : (f + 1) * n_pairs]
        nodes = ["root"] + [f"internal-{i}" for i in range(n_pairs)] + [f"seq{r}" for r in range(n_seqs)]
        edges = []
        for i in range(n_pairs):
            edges.append(f"root internal-{i} 1.0")
            edges.append(f"internal-{i} seq{2 * i} {repr(float(t[i]) / 2)}")
            edges.append(f"internal-{i} seq{2 * i + 1} {repr(float(t[i]) / 2)}")
        with open(os.path.join(out_dir, "tree_dir", name + ".txt"), "w") as fh:
            fh.write(f"{len(nodes)} nodes\n" + "\n".join(nodes) + f"\n{len(edges)} edges\n" + "\n".join(edges) + "\n")
        if batch.kind == "lg":
            rv = batch.rate_vals[fd["rate_off"] : fd["rate_off"] + fd["n_rates"]]
            with open(os.path.join(out_dir, third, name + ".txt"), "w") as fh:
                fh.write(f"{n_sites} sites\n" + " ".join(repr(float(rv[c])) for c in cat_of_site))
        else:
            cmap = np.zeros((n_sites, n_sites), dtype=np.uint8)
            cs = batch.aux[fd["aux_off"] : fd["aux_off"] + fd["aux_cnt"]]
            cmap[cs[:, 0], cs[:, 1]] = 1
            cmap[cs[:, 1], cs[:, 0]] = 1
            with open(os.path.join(out_dir, third, name + ".txt"), "w") as fh:
                fh.write(f"{n_sites} sites\n" + "\n".join("".join(map(str, row)) for row in cmap) + "\n")
    return names


def synthetic_fc(n_fams: int, n_seqs: int, n_sites: int, seed: int = 0, gap_frac: float = 0.13,
                 cherry_div: float = 0.5, leaf_div: float = 0.19):
    """Families for the FastCherries kernels: (flat uint8 residue buffer, FC_FAMILY_DTYPE array).
    Per family: a root sequence; every cherry's ancestor is the root with ``cherry_div`` of its
    sites resampled; both leaves are the ancestor with ``leaf_div`` resampled (so partners agree at
    ~66 % of their sites as on the Pfam demo data); 13 % gaps; rows in random order, natural
    column order, rows padded to a multiple of 16 with the skip code."""
    from ._lib import FC_FAMILY_DTYPE

    rng = np.random.default_rng(seed)
    stride = max(16, (n_sites + 15) // 16 * 16)
    fams = np.zeros(n_fams, dtype=FC_FAMILY_DTYPE)
    f = np.arange(n_fams, dtype=np.int64)
    fams["msa_off"] = f * n_seqs * stride
    fams["n_seqs"] = n_seqs
    fams["row_stride"] = stride
    fams["n_sites"] = n_sites
    fams["cherry_off"] = f * (n_seqs // 2)
    fams["site_off"] = f * n_sites
    fams["seq_off"] = f * n_seqs
    msa = np.full((n_fams, n_seqs, stride), SKIP, dtype=np.uint8)
    n_anc = (n_seqs + 1) // 2

    def resample(x, frac):
        fresh = rng.integers(0, 20, x.shape, dtype=np.uint8)
        return np.where(rng.random(x.shape) < frac, fresh, x)

    for i in range(n_fams):
        root = rng.integers(0, 20, (1, n_sites), dtype=np.uint8)
        anc = resample(np.repeat(root, n_anc, axis=0), cherry_div)
        leaves = resample(np.repeat(anc, 2, axis=0)[:n_seqs], leaf_div)
        leaves[rng.random(leaves.shape) < gap_frac] = SKIP
        msa[i, :, :n_sites] = leaves[rng.permutation(n_seqs)]
    return msa.reshape(-1), fams
