"""numpy emulation of what the counting kernels do with an encoded CountBatch.

Test-only: lets the CPU test-suite check the host-side ingest/encoding (column sorting,
padding, descriptors, tiles, pair rows) against the oracle without a GPU.  Mirrors
cherryml_b200/csrc/count_kernels.cu one-to-one (bucket table, then per tile/pair/word).
"""
import numpy as np

from oracle.counting_oracle import quantization_idx_vec


def emulate_bucket_table(batch, grid):
    P, R = batch.n_pairs, batch.r_pad
    tab = np.full((P, R), 255, dtype=np.uint8)
    for p in range(P):
        fd = batch.fams[batch.pair_fam[p]]
        rv = batch.rate_vals[fd["rate_off"] : fd["rate_off"] + fd["n_rates"]]
        b = quantization_idx_vec(batch.pair_t[p] * rv, grid)
        tab[p, : len(rv)] = np.where(b >= 0, b, 255).astype(np.uint8)
    return tab


def emulate_count(batch, grid, S):
    grid = np.asarray(grid, dtype=np.float64)
    K = len(grid)
    tab = emulate_bucket_table(batch, grid)
    seen_pairs = np.zeros(batch.n_pairs, dtype=np.int64)
    if batch.kind == "lg":
        raw = np.zeros((K, S, S), dtype=np.int64)
        for tl in batch.tiles:
            fd = batch.fams[tl["fam"]]
            stride = int(fd["row_stride"])
            assert stride % 16 == 0 and fd["n_chunks"] * 16 == stride and fd["msa_off"] % 16 == 0
            gc = batch.aux[fd["aux_off"] : fd["aux_off"] + stride // 4].astype(np.int64)
            for p in range(tl["pair_begin"], tl["pair_begin"] + tl["n_pairs"]):
                assert batch.pair_fam[p] == tl["fam"]
                seen_pairs[p] += 1
                ra = batch.msa[fd["msa_off"] + batch.pair_a[p] * stride :][:stride].astype(np.int64)
                rb = batch.msa[fd["msa_off"] + batch.pair_b[p] * stride :][:stride].astype(np.int64)
                bucket = np.repeat(tab[p][gc].astype(np.int64), 4)
                ok = (bucket != 255) & (ra < S) & (rb < S)
                np.add.at(raw, (bucket[ok], ra[ok], rb[ok]), 1)
    else:
        n = S * S
        raw = np.zeros((K, n, n), dtype=np.int64)
        for tl in batch.tiles:
            fd = batch.fams[tl["fam"]]
            stride = int(fd["row_stride"])
            cs = batch.aux[fd["aux_off"] : fd["aux_off"] + fd["aux_cnt"]]
            for p in range(tl["pair_begin"], tl["pair_begin"] + tl["n_pairs"]):
                seen_pairs[p] += 1
                b = int(tab[p, 0])
                if b == 255 or len(cs) == 0:
                    continue
                ra = batch.msa[fd["msa_off"] + batch.pair_a[p] * stride :][:stride].astype(np.int64)
                rb = batch.msa[fd["msa_off"] + batch.pair_b[p] * stride :][:stride].astype(np.int64)
                assert stride % 16 == 0 and 2 * len(cs) <= stride
                assert np.all(ra[2 * len(cs):] == S) and np.all(rb[2 * len(cs):] == S)
                # contact-paired rows: bytes 2c, 2c+1 = residues at the two sites of contact c
                xi, xj, yi, yj = ra[0:2 * len(cs):2], ra[1:2 * len(cs):2], rb[0:2 * len(cs):2], rb[1:2 * len(cs):2]
                ok = (xi < S) & (xj < S) & (yi < S) & (yj < S)
                np.add.at(raw[b], (xi[ok] * S + xj[ok], yi[ok] * S + yj[ok]), 1)
    assert np.all(seen_pairs == 1), "every pair must be covered by exactly one tile"
    return raw


def emulate_symmetrize(raw, kind, S, directed):
    raw = raw.astype(np.float64)
    if kind == "lg":
        return raw if directed else 0.5 * (raw + raw.transpose(0, 2, 1))
    n = S * S
    perm = (np.arange(n) % S) * S + np.arange(n) // S
    swapped = raw[:, perm][:, :, perm]
    if directed:
        return 0.5 * (raw + swapped)
    return 0.25 * (raw + raw.transpose(0, 2, 1) + swapped + swapped.transpose(0, 2, 1))
