"""bench.py's tree-likelihood section: ``dp_likelihood_computation`` on synthetic Pfam-shaped
families, beside the numpy/scipy oracle port timed on the host (rank 0, N = 1) with the values
compared (1e-9 relative)."""
import time
from typing import Dict

import numpy as np

from cherryml_b200.io import Tree, read_rate_matrix
from cherryml_b200.markov_chain import chain_product, compute_stationary_distribution, get_lg_path
from cherryml_b200.utils import amino_acids
from cherryml_b200.evaluation._likelihood import dp_likelihood_computation


def random_binary_tree(rng, n_leaves: int) -> Tree:
    """Random topology by repeatedly joining two random subtrees (Yule-like), lognormal lengths."""
    tree = Tree()
    roots = []
    for i in range(n_leaves):
        tree.add_node(f"leaf{i}")
        roots.append(f"leaf{i}")
    k = 0
    edges = []
    while len(roots) > 1:
        i, j = sorted(rng.choice(len(roots), 2, replace=False).tolist(), reverse=True)
        a, b = roots.pop(i), roots.pop(j)
        parent = f"int{k}"
        k += 1
        tree.add_node(parent)
        edges.append((parent, a, float(rng.lognormal(-2.5, 1.0))))
        edges.append((parent, b, float(rng.lognormal(-2.5, 1.0))))
        roots.append(parent)
    # the Tree class wants parents before children in add_edge order only for lookups: any order works
    for e in reversed(edges):
        tree.add_edge(*e)
    return tree


def _family(rng, n_leaves, n_sites, n_pairs, n_cats):
    tree = random_binary_tree(rng, n_leaves)
    aa = np.array(amino_acids)
    msa = {}
    for v in tree.leaves():
        s = aa[rng.integers(0, 20, n_sites)]
        s[rng.random(n_sites) < 0.13] = "-"
        msa[v] = "".join(s)
    cmap = None
    if n_pairs:
        cmap = np.zeros((n_sites, n_sites))
        sites = rng.permutation(n_sites)[: 2 * n_pairs]
        for a, b in zip(sites[0::2], sites[1::2]):
            cmap[a, b] = cmap[b, a] = 1
    rates = [float(r) for r in rng.choice(np.linspace(0.3, 2.0, n_cats), n_sites)]
    return tree, msa, cmap, rates


def bench_likelihood(device, cpu_baseline: bool = True, seed: int = 0) -> Dict:
    rng = np.random.default_rng(seed)
    Q1 = read_rate_matrix(get_lg_path()).to_numpy(dtype=np.float64)
    pi1 = compute_stationary_distribution(Q1)
    Q2 = chain_product(Q1, Q1)
    pi2 = compute_stationary_distribution(Q2)
    out: Dict = {"metric": "seconds per family, dp_likelihood_computation (tree, MSA, rates -> per-site log-likelihoods)"}
    for key, (n_leaves, n_sites, n_pairs, n_cats) in (("sites_1024x300", (1024, 300, 0, 4)),
                                                       ("pairs_256x300_60_contacts", (256, 300, 60, 4))):
        tree, msa, cmap, rates = _family(rng, n_leaves, n_sites, n_pairs, n_cats)
        kw = dict(tree=tree, msa=msa, contact_map=cmap, site_rates=rates, amino_acids=amino_acids, pi_1=pi1, Q_1=Q1,
                  pi_2=pi2 if n_pairs else None, Q_2=Q2 if n_pairs else None, device_1=str(device))
        dp_likelihood_computation(**kw)  # warm-up
        best = float("inf")
        for _ in range(3):
            t0 = time.perf_counter()
            ll, lls = dp_likelihood_computation(**kw)
            best = min(best, time.perf_counter() - t0)
        res = {"workload": f"{n_leaves} leaves x {n_sites} sites, {n_pairs} contacting pairs (400-state model), "
                           f"{n_cats} rate categories, random binary tree", "seconds": best,
               "log_likelihood": ll, "note": "host set-up (tree arrays, H2D) + cherry_expm_batched + "
                                             "cherry_tree_log_likelihood + D2H"}
        if cpu_baseline:
            from oracle.likelihood_oracle import log_likelihood

            t0 = time.perf_counter()
            ll_o, lls_o = log_likelihood(tree, msa, cmap, rates, amino_acids, pi1, Q1, pi2 if n_pairs else None,
                                         Q2 if n_pairs else None)
            cpu_s = time.perf_counter() - t0
            res["cpu_baseline"] = {"value": cpu_s, "unit": "s", "cores": 1, "kind": "port",
                                   "sample": "the same family, numpy/scipy restatement of the reference function "
                                             "(oracle/likelihood_oracle.py)"}
            res["matches_oracle_1e-9"] = bool(abs(ll - ll_o) <= 1e-9 * abs(ll_o)
                                              and np.allclose(lls, lls_o, rtol=1e-8, atol=1e-9))
        out[key] = res
    return out
