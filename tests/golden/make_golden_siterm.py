"""Goldens for the batched per-site fit: RUN THE UNMODIFIED reference function
``quantized_transitions_mle_vectorized_over_sites`` (build container only).

    python tests/golden/make_golden_siterm.py  ->  tests/golden/siterm/*.npz
"""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests", "golden"))
OUT = os.path.join(REPO, "tests/golden/siterm")


def dna_case(L=16, B=11):
    base = (np.ones((4, 4)) - np.eye(4)) / 3.0 - np.eye(4)
    Qs = np.stack([base * (1.0 if l % 2 == 0 else 3.0) for l in range(L)])
    n = B // 2
    times = np.array([[(1.1**i) / (1.0 if l % 4 in (0, 1) else 10.0) for i in range(-n, n + 1)] for l in range(L)])
    return Qs, times


def main():
    from make_golden_fit import import_reference

    import_reference()
    import torch
    from cherryml._siterm._cherryml_vectorized import quantized_transitions_mle_vectorized_over_sites as ref_fn
    from cherryml.io import read_rate_matrix

    os.makedirs(OUT, exist_ok=True)
    # 1. DNA-shaped synthetic counts (counts := expm(t Q_true), the reference's own KAT), no init
    Qs, times = dna_case()
    counts = np.stack([torch.matrix_exp(torch.tensor(times[l])[:, None, None] * torch.tensor(Qs[l])).numpy()
                       for l in range(len(Qs))])
    r = ref_fn(counts, times, num_epochs=40, initialization=None)
    np.savez_compressed(os.path.join(OUT, "dna_noinit.npz"), counts=counts, times=times, Q_true=Qs,
                        res=r["res"], loss_per_epoch=r["loss_per_epoch"],
                        loss_per_epoch_per_site=r["loss_per_epoch_per_site"])
    r0 = ref_fn(counts, times, num_epochs=0, initialization=Qs)
    np.savez_compressed(os.path.join(OUT, "dna_init_0epochs.npz"), counts=counts, times=times, init=Qs, res=r0["res"])
    # 2. 20 states, random integer counts, initialisation = LG scaled by a per-site rate (what SiteRM passes)
    rng = np.random.default_rng(0)
    L, B, N = 12, 6, 20
    lg = read_rate_matrix("/root/reference/data/rate_matrices/lg.txt").to_numpy()
    rates = rng.uniform(0.2, 3.0, L)
    init = lg[None] * rates[:, None, None]
    times = np.sort(np.exp(rng.uniform(np.log(0.02), np.log(2.0), (L, B))), axis=1)
    counts = rng.integers(0, 6, size=(L, B, N, N)).astype(np.float64) + 5.0 * np.eye(N)[None, None]
    counts[:, -1] = 0.0
    times[:, -1] = 1.0  # padded bucket: time 1.0, zero counts (SiteRM pads like this)
    r = ref_fn(counts, times, num_epochs=30, initialization=init)
    np.savez_compressed(os.path.join(OUT, "aa_init.npz"), counts=counts, times=times, init=init, res=r["res"],
                        loss_per_epoch=r["loss_per_epoch"], loss_per_epoch_per_site=r["loss_per_epoch_per_site"])
    print("written", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
