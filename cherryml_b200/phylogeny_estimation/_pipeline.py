"""FastCherries -> LG transition counting in device memory.

The reference hands trees and site rates from the tree estimator to the counting stage through
text files (``estimation_end_to_end/_cherry.py:279-336``).  Here the residues uploaded for
FastCherries stay on the device: the cherries, branch lengths and site-rate categories it
produced are turned into the counting kernels' batch layout by ``cherry_fc_lengths_and_rates``
(host: the exact values the text files would carry) and ``cherry_fc_relayout_lg`` (device: rows
in cherry order, columns sorted by rate category), and ``cherry_count_lg`` runs on it.  The batch
is array-for-array the one ``cherry_ingest_lg`` builds from the files ``fast_cherries`` writes, so
the counts are identical (tests/test_gpu_fast_cherries.py).
"""
import math
import os
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch

from .. import _lib
from ..counting._device import DeviceBatch, count_raw, sorted_grid, symmetrize
from ..counting._ingest import TARGET_CHUNKS_PER_TILE
from . import _fast_cherries as fc


def count_layout(fams: np.ndarray, site_cat: np.ndarray, rate_table: np.ndarray, n_threads: Optional[int] = None):
    """Host metadata of the LG counting batch for FastCherries results by the library's host threads
    (``cherry_fc_count_layout``); same dictionary as ``count_layout_numpy`` (kept as the readable
    definition and for the tests)."""
    lib = _lib.load()
    F, R = len(fams), rate_table.shape[1]
    fams_c = np.ascontiguousarray(fams)
    sc = np.ascontiguousarray(site_cat, dtype=np.int32)
    rt = np.ascontiguousarray(rate_table, dtype=np.float64)
    sizes = np.zeros(6, dtype=np.int64)
    nt = n_threads or os.cpu_count() or 1
    _lib.check(lib.cherry_fc_count_layout(_lib.ptr(fams_c), F, _lib.ptr(sc), _lib.ptr(rt), R, TARGET_CHUNKS_PER_TILE,
                                          0, 0, 0, 0, 0, _lib.ptr(sizes), nt), "cherry_fc_count_layout")
    dest = np.empty(max(1, len(sc)), dtype=np.int32)
    aux = np.empty(max(1, int(sizes[1])), dtype=np.uint16)
    rate_vals = np.empty(max(1, int(sizes[2])), dtype=np.float64)
    out = np.zeros(max(1, F), dtype=_lib.FAM_DESC_DTYPE)
    tiles = np.zeros(max(1, int(sizes[3])), dtype=_lib.TILE_DTYPE)
    _lib.check(lib.cherry_fc_count_layout(_lib.ptr(fams_c), F, _lib.ptr(sc), _lib.ptr(rt), R, TARGET_CHUNKS_PER_TILE,
                                          _lib.ptr(dest), _lib.ptr(aux), _lib.ptr(rate_vals), _lib.ptr(out),
                                          _lib.ptr(tiles), _lib.ptr(sizes), nt), "cherry_fc_count_layout")
    return dict(dest=dest[: len(sc)], aux=aux[: int(sizes[1])], rate_vals=rate_vals[: int(sizes[2])], fams=out[:F],
                tiles=tiles[: int(sizes[3])], r_pad=int(sizes[4]) if F else 4, msa_bytes=int(sizes[0]),
                examined=int(sizes[5]))


def count_layout_numpy(fams: np.ndarray, site_cat: np.ndarray, rate_table: np.ndarray):
    """Host metadata of the LG counting batch for FastCherries results (vectorised over all
    families): per-site destination column, group categories, distinct rate values, family
    descriptors and tiles -- the layout rules of counting/_ingest.py (categories = distinct site
    rates ascending, sites in order inside a category, categories padded to 4 sites, row stride
    a multiple of 16)."""
    F = len(fams)
    R = rate_table.shape[1]
    L = fams["n_sites"].astype(np.int64)
    n_pairs = (fams["n_seqs"] // 2).astype(np.int64)
    T = int(L.sum())
    fam_of_site = np.repeat(np.arange(F, dtype=np.int64), L)
    key = fam_of_site * R + site_cat.astype(np.int64)
    uniq, inv, cnt = np.unique(key, return_inverse=True, return_counts=True)
    G = len(uniq)
    fam_of_group = uniq // R
    cat_of_group = (uniq % R).astype(np.int64)
    first_group = np.searchsorted(fam_of_group, np.arange(F))          # first group of every family
    groups_per_fam = np.bincount(fam_of_group, minlength=F).astype(np.int64)
    padded = (cnt + 3) // 4 * 4
    gstart = np.cumsum(padded) - padded
    fam_base = np.zeros(F, dtype=np.int64)
    has = groups_per_fam > 0
    fam_base[has] = gstart[first_group[has]]
    start_in_fam = gstart - fam_base[fam_of_group]
    total = np.bincount(fam_of_group, weights=padded, minlength=F).astype(np.int64)
    stride = np.maximum(16, (total + 15) // 16 * 16)
    # destination column of every site: stable inside its (family, category) group
    order = np.argsort(key, kind="stable")
    group_first_pos = np.cumsum(cnt) - cnt
    dest = np.empty(T, dtype=np.int32)
    dest[order] = (start_in_fam[inv[order]] + (np.arange(T) - group_first_pos[inv[order]])).astype(np.int32)
    # group categories: one uint16 per 4 sites, stride / 4 entries per family
    aux_cnt = stride // 4
    aux_off = np.cumsum(aux_cnt) - aux_cnt
    aux = np.zeros(int(aux_cnt.sum()), dtype=np.uint16)
    words = padded // 4
    cprime = np.arange(G) - first_group[fam_of_group]
    pos = np.repeat(aux_off[fam_of_group] + start_in_fam // 4, words) + (np.arange(int(words.sum())) -
                                                                          np.repeat(np.cumsum(words) - words, words))
    aux[pos] = np.repeat(cprime, words).astype(np.uint16)
    # distinct rate values per family (a family without sites gets the single rate 1.0, like the ingest)
    n_rates = np.where(has, groups_per_fam, 1)
    rate_off = np.cumsum(n_rates) - n_rates
    rate_vals = np.ones(int(n_rates.sum()), dtype=np.float64)
    rate_vals[rate_off[fam_of_group] + cprime] = rate_table[fam_of_group, cat_of_group]
    out = np.zeros(F, dtype=_lib.FAM_DESC_DTYPE)
    rows = 2 * n_pairs
    out["msa_off"] = np.cumsum(rows * stride) - rows * stride
    out["row_stride"] = stride
    out["n_chunks"] = stride // 16
    out["aux_off"] = aux_off
    out["aux_cnt"] = aux_cnt
    out["rate_off"] = rate_off
    out["n_rates"] = n_rates
    per_tile = np.maximum(1, TARGET_CHUNKS_PER_TILE // np.maximum(1, stride // 16))
    pair_off = np.cumsum(n_pairs) - n_pairs
    n_tiles_f = (n_pairs + per_tile - 1) // per_tile
    tile_fam = np.repeat(np.arange(F), n_tiles_f)
    tile_k = np.arange(int(n_tiles_f.sum())) - np.repeat(np.cumsum(n_tiles_f) - n_tiles_f, n_tiles_f)
    tl = np.zeros(len(tile_fam), dtype=_lib.TILE_DTYPE)
    tl["fam"] = tile_fam
    tl["pair_begin"] = pair_off[tile_fam] + tile_k * per_tile[tile_fam]
    tl["n_pairs"] = np.minimum(per_tile[tile_fam], n_pairs[tile_fam] - tile_k * per_tile[tile_fam])
    r_pad = (int(n_rates.max()) + 3) // 4 * 4 if F else 4
    msa_bytes = int((rows * stride).sum())
    return dict(dest=dest, aux=aux, rate_vals=rate_vals, fams=out, tiles=tl, r_pad=r_pad, msa_bytes=msa_bytes,
                examined=int((n_pairs * L).sum()))


def lg_batch_from_fast_cherries(fams: np.ndarray, out: Dict, grid: np.ndarray, cats: np.ndarray, S: int,
                                float32_branch_lengths: bool, device="cuda",
                                n_threads: Optional[int] = None) -> DeviceBatch:
    """``out`` = ``fast_cherries_device(..., keep_device=True)``.  Returns the device batch for
    ``count_raw`` (counting/_device.py)."""
    lib = _lib.load()
    dev = torch.device(device)
    F = len(fams)
    n_cherries = int((fams["n_seqs"] // 2).sum())
    R = len(cats)
    pair_t = np.zeros(max(1, n_cherries))
    rate_table = np.zeros((max(1, F), R))
    fams_c = np.ascontiguousarray(fams)
    li = np.ascontiguousarray(out["len_idx"], dtype=np.int32)
    sc = np.ascontiguousarray(out["site_cat"], dtype=np.int32)
    g = np.ascontiguousarray(grid, dtype=np.float64)
    c = np.ascontiguousarray(cats, dtype=np.float64)
    _lib.check(lib.cherry_fc_lengths_and_rates(_lib.ptr(fams_c), F, _lib.ptr(li), _lib.ptr(sc), _lib.ptr(g), len(g),
                                               _lib.ptr(c), R, int(float32_branch_lengths), _lib.ptr(pair_t),
                                               _lib.ptr(rate_table), n_threads or os.cpu_count() or 1),
               "cherry_fc_lengths_and_rates")
    lay = count_layout(fams_c, sc, rate_table[:F])
    n_pairs_f = (fams["n_seqs"] // 2).astype(np.int64)
    local = np.arange(n_cherries, dtype=np.int64) - np.repeat(np.cumsum(n_pairs_f) - n_pairs_f, n_pairs_f)
    with torch.cuda.device(dev):
        up = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).to(dev)  # noqa: E731
        d_out_fams = up(lay["fams"])
        d_dest = up(lay["dest"]) if len(lay["dest"]) else torch.zeros(4, dtype=torch.uint8, device=dev)
        msa_out = torch.full((max(16, lay["msa_bytes"]),), S, dtype=torch.uint8, device=dev)
        _lib.check(lib.cherry_fc_relayout_lg(_lib.ptr(out["d_msa"]), _lib.ptr(out["d_fams"]), _lib.ptr(d_out_fams), F,
                                             _lib.ptr(out["d_pair_a"]), _lib.ptr(out["d_pair_b"]), _lib.ptr(d_dest),
                                             _lib.ptr(msa_out), _lib.current_stream_ptr()), "cherry_fc_relayout_lg")
        return DeviceBatch(
            kind="lg", msa=msa_out, fams=d_out_fams, pair_a=up((2 * local).astype(np.int32)),
            pair_b=up((2 * local + 1).astype(np.int32)), pair_t=up(pair_t[:n_cherries]),
            pair_fam=up(np.repeat(np.arange(F, dtype=np.int32), n_pairs_f)), rate_vals=up(lay["rate_vals"]),
            aux=up(lay["aux"]), tiles=up(lay["tiles"]), r_pad=lay["r_pad"], n_pairs=n_cherries,
            n_tiles=len(lay["tiles"]), n_sites_examined=lay["examined"],
        )


def fast_cherries_then_count_lg(msa: np.ndarray, fams: np.ndarray, alphabet: Sequence[str], Q: np.ndarray,
                                quantization_points: Sequence[float], num_rate_categories: int = 20,
                                max_iters: int = 50, seed: int = 1234, quantization_grid_center: float = 0.03,
                                quantization_grid_step: float = 1.1, quantization_grid_num_steps: int = 64,
                                float32_branch_lengths: bool = True, device="cuda") -> Tuple[torch.Tensor, Dict]:
    """Encoded MSAs -> (symmetrised LG count tensor [K, S, S] on the device, FastCherries results):
    tree estimation and counting back to back on the resident residues."""
    S = len(alphabet)
    grid_fc = fc.quantization_grid(quantization_grid_center, quantization_grid_step, quantization_grid_num_steps)
    cats = fc.ble_rate_categories(num_rate_categories)
    weights = fc.initial_rate_weights(cats)
    priors = np.array([2 * math.log(r) - 3 * r for r in cats])
    table = fc.log_transition_table(np.asarray(Q, dtype=np.float64), grid_fc, cats, device)
    out = fc.fast_cherries_device(msa, fams, S, table, priors, weights, seed, max_iters, device, keep_device=True)
    batch = lg_batch_from_fast_cherries(fams, out, grid_fc, cats, S, float32_branch_lengths, device)
    grid = sorted_grid(quantization_points)
    K = int(grid.size)
    with torch.cuda.device(torch.device(device)):
        grid_dev = torch.from_numpy(grid).to(batch.msa.device)
        counts = symmetrize(count_raw(batch, grid_dev, K, S), "lg", K, S, False)
    return counts, out
