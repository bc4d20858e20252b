"""Goldens for the SiteRM orchestration: RUN THE UNMODIFIED reference function
``_estimate_site_specific_rate_matrices_given_tree_and_site_rates`` (vectorised branch, CPU;
build container only) on seeded trees / MSAs.

    python tests/golden/make_golden_siterm_estimate.py  ->  tests/golden/siterm/estimate_*.npz
"""
import json
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests", "golden"))
OUT = os.path.join(REPO, "tests/golden/siterm")
AA = "ARNDCQEGHILKMFPSTWYV"


def main():
    from make_golden_fit import import_reference
    from make_golden_likelihood import random_tree

    import_reference()
    from cherryml._siterm._site_specific_rate_matrix import (
        _estimate_site_specific_rate_matrices_given_tree_and_site_rates as ref_fn,
    )
    from cherryml.io import Tree, read_rate_matrix

    lg = read_rate_matrix("/root/reference/data/rate_matrices/lg.txt").to_numpy()
    rng = np.random.default_rng(11)
    for name, n_leaves, L, alphabet, gap, lam, steps, epochs in [
        ("aa", 14, 9, list(AA), 0.15, 0.5, 8, 25),
        ("aa_gap_state", 21, 6, list(AA) + ["-"], 0.2, 0.3, 11, 20),
    ]:
        names, parent, length = random_tree(rng, n_leaves)
        tree = Tree()
        tree.add_nodes(names)
        for i in range(1, len(names)):
            tree.add_edge(names[parent[i]], names[i], length[i])
        msa = {}
        for lf in [n for n in names if n.startswith("leaf")]:
            s = rng.choice(list(AA), L)
            s[rng.random(L) < gap] = "-"
            msa[lf] = "".join(s)
        if name == "aa":
            for lf in msa:  # one all-gap column: its site keeps the prior
                msa[lf] = msa[lf][:3] + "-" + msa[lf][4:]
        S = len(alphabet)
        if S == 21:
            Q0 = np.zeros((21, 21))
            Q0[:20, :20] = lg
            Q0[:20, 20] = 0.05
            Q0[20, :20] = 0.05 * 20 / 20
            np.fill_diagonal(Q0, 0)
            Q0 = (Q0 + Q0.T) / 2  # symmetric => reversible with uniform stationary distribution
            np.fill_diagonal(Q0, -Q0.sum(axis=1))
        else:
            Q0 = lg
        rates = [float(r) for r in rng.uniform(0.3, 2.5, L)]
        grid = [0.03 * 1.5 ** i for i in range(-steps, steps + 1)]
        r = ref_fn(tree=tree, site_rates=rates, msa=msa, alphabet=alphabet, regularization_strength=lam,
                   regularization_rate_matrix=Q0, quantization_points=grid, optimization_num_epochs=epochs,
                   use_vectorized_cherryml_implementation=True)
        meta = dict(names=names, parent=parent, length=length, msa=msa, alphabet=alphabet, lam=lam, grid=grid,
                    epochs=epochs, rates=rates)
        np.savez_compressed(os.path.join(OUT, f"estimate_{name}.npz"), res=r["res"], Q0=Q0, meta=json.dumps(meta))
        print(name, r["res"].shape, float(np.abs(r["res"]).max()))


if __name__ == "__main__" and "--public-api" not in sys.argv:
    main()


def public_api_golden():
    """learn_site_rate_matrices with a given tree (site-rate estimation + SiteRM), unmodified
    reference, non-Cython site-rate path."""
    from make_golden_fit import import_reference
    from make_golden_likelihood import random_tree

    import_reference()
    import pandas as pd
    from cherryml._siterm._learn_site_rate_matrix import (get_standard_site_rate_grid, get_standard_site_rate_prior,
                                                           learn_site_rate_matrices)
    from cherryml.io import Tree, read_rate_matrix

    lg = read_rate_matrix("/root/reference/data/rate_matrices/lg.txt")
    rng = np.random.default_rng(5)
    names, parent, length = random_tree(rng, 16)
    tree = Tree()
    tree.add_nodes(names)
    for i in range(1, len(names)):
        tree.add_edge(names[parent[i]], names[i], length[i])
    L = 12
    msa = {}
    anc = rng.choice(list(AA), L)
    for lf in [n for n in names if n.startswith("leaf")]:
        s = anc.copy()
        flip = rng.random(L) < np.linspace(0.02, 0.9, L)  # slow sites first, fast sites last
        s[flip] = rng.choice(list(AA), int(flip.sum()))
        s[rng.random(L) < 0.1] = "-"
        msa[lf] = "".join(s)
    r = learn_site_rate_matrices(
        tree=tree, leaf_states=msa, alphabet=list(AA), regularization_rate_matrix=lg, regularization_strength=0.5,
        use_vectorized_implementation=True, site_rate_grid=get_standard_site_rate_grid(),
        site_rate_prior=get_standard_site_rate_prior(), num_epochs=20, use_fast_site_rate_implementation=False,
        quantization_grid_num_steps=8,
    )
    meta = dict(names=names, parent=parent, length=length, msa=msa)
    np.savez_compressed(os.path.join(OUT, "public_api_tree_given.npz"), res=r["learnt_rate_matrices"],
                        site_rates=np.array(r["learnt_site_rates"]), meta=json.dumps(meta))
    print("public api:", r["learnt_rate_matrices"].shape, r["learnt_site_rates"])


if __name__ == "__main__" and "--public-api" in sys.argv:
    public_api_golden()
