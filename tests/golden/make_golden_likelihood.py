"""Regenerate tests/golden/likelihood/cases.npz by RUNNING THE UNMODIFIED REFERENCE FUNCTION
``cherryml.evaluation._likelihood.dp_likelihood_computation`` (imported from /root/reference
with the stubs of make_golden.import_reference; build container only) on seeded random trees,
MSAs with gaps, matchings of contacting sites and site rates.

    python tests/golden/make_golden_likelihood.py

Stored per case: the tree (parent index + branch length per node, nodes in insertion order),
the leaf sequences, the contact pairs, the site rates, which rate matrices were used, and the
reference's (ll, lls).  Q_1 is the LG or WAG matrix shipped in cherryml_b200/data; Q_2 is
chain_product(Q_1, Q_1) or a seeded reversible perturbation of it.
"""
import json
import os
import sys
import tempfile

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests", "golden"))
OUT = os.path.join(REPO, "tests/golden/likelihood")
AA = "ARNDCQEGHILKMFPSTWYV"


def random_tree(rng, n_leaves, max_children=3):
    """-> (names, parent index (-1 root), length); internal nodes get 2..max_children children."""
    names, parent, length = ["root"], [-1], [0.0]
    open_nodes, leaves = [0], 0
    while leaves < n_leaves:
        p = open_nodes.pop(rng.integers(0, len(open_nodes))) if open_nodes else 0
        kids = int(rng.integers(2, max_children + 1))
        for _ in range(kids):
            idx = len(names)
            parent.append(p)
            length.append(float(rng.lognormal(-1.5, 1.2)))  # zero lengths: see the FastTree KATs
            if leaves < n_leaves and (rng.random() < 0.6 or leaves + len(open_nodes) + 1 >= n_leaves):
                names.append(f"leaf{leaves}")
                leaves += 1
            else:
                names.append(f"int{idx}")
                open_nodes.append(idx)
    # internal nodes left without children become leaves too
    has_child = set(parent)
    for i, nm in enumerate(names):
        if nm.startswith("int") and i not in has_child:
            names[i] = f"leaf{leaves}"
            leaves += 1
    return names, parent, length


def perturbed_pair_matrix(rng, Q1):
    from cherryml_b200.markov_chain import chain_product, compute_stationary_distribution

    base = chain_product(Q1, Q1)
    pi = compute_stationary_distribution(base)
    n = base.shape[0]
    sym = rng.uniform(0.5, 1.5, (n, n))
    sym = (sym + sym.T) / 2
    exch = base / pi[None, :]  # symmetric exchangeabilities
    Q = exch * sym * pi[None, :]
    Q[np.arange(n), np.arange(n)] = 0
    Q[np.arange(n), np.arange(n)] = -Q.sum(axis=1)
    return Q


def main():
    from make_golden import import_reference

    import_reference()
    from cherryml.evaluation._likelihood import dp_likelihood_computation
    from cherryml.io import Tree
    from cherryml.markov_chain import FactorizedReversibleModel

    from cherryml_b200.io import read_rate_matrix
    from cherryml_b200.markov_chain import chain_product, compute_stationary_distribution

    from cherryml_b200.markov_chain import _rate_matrix_path

    mats = {k: read_rate_matrix(_rate_matrix_path(k)).to_numpy() for k in ("lg", "wag")}
    rng = np.random.default_rng(7)
    cases = []
    specs = [  # n_leaves, n_sites, n_pairs, gap, n_cats, Q1, pair model, reversible flags
        (3, 6, 0, 0.0, 1, "wag", None, True), (4, 9, 2, 0.2, 2, "wag", "product", True),
        (8, 30, 5, 0.15, 4, "lg", "product", False), (17, 40, 8, 0.3, 20, "lg", "perturbed", True),
        (40, 25, 6, 0.1, 4, "wag", "perturbed", False), (90, 60, 0, 0.13, 20, "lg", None, True),
        (12, 20, 10, 0.5, 3, "lg", "product", True),
    ]
    os.makedirs(OUT, exist_ok=True)
    arrays = {}
    for ci, (n_leaves, L, n_pairs, gap, n_cats, q1, pair_model, reversible) in enumerate(specs):
        names, parent, length = random_tree(rng, n_leaves)
        tree = Tree()
        tree.add_nodes(names)
        for i in range(1, len(names)):
            tree.add_edge(names[parent[i]], names[i], length[i])
        leaves = [n for n in names if n.startswith("leaf")]
        msa = {}
        for lf in leaves:
            s = rng.choice(list(AA), L)
            s[rng.random(L) < gap] = "-"
            msa[lf] = "".join(s)
        sites = rng.permutation(L)[: 2 * n_pairs]
        pairs = [(int(min(a, b)), int(max(a, b))) for a, b in zip(sites[0::2], sites[1::2])]
        cmap = np.eye(L) if n_pairs == 0 and ci % 2 == 0 else np.zeros((L, L))
        for a, b in pairs:
            cmap[a, b] = cmap[b, a] = 1
        rates = [float(r) for r in rng.choice(0.3 * np.log(2 + np.arange(n_cats)), L)]
        Q1 = mats[q1]
        pi1 = compute_stationary_distribution(Q1)
        if pair_model is None:
            Q2 = pi2 = None
        else:
            Q2 = chain_product(Q1, Q1) if pair_model == "product" else perturbed_pair_matrix(np.random.default_rng(ci), Q1)
            pi2 = compute_stationary_distribution(Q2)
        with tempfile.TemporaryDirectory() as tmp:
            ll, lls = dp_likelihood_computation(
                tree=tree, msa=msa, contact_map=cmap if (pairs or ci % 2 == 0) else None, site_rates=rates,
                amino_acids=list(AA), pi_1=pi1, Q_1=Q1,
                fact_1=FactorizedReversibleModel(Q1) if reversible else None, reversible_1=reversible, device_1="cpu",
                pi_2=pi2, Q_2=Q2, fact_2=FactorizedReversibleModel(Q2) if (reversible and Q2 is not None) else None,
                reversible_2=reversible if Q2 is not None else None, device_2="cpu" if Q2 is not None else None,
                output_profiling_path=os.path.join(tmp, "prof.txt"),
            )
        cases.append({
            "names": names, "parent": parent, "length": length, "msa": msa, "pairs": pairs,
            "contact_map_given": bool(pairs or ci % 2 == 0), "site_rates": rates, "Q1": q1, "pair_model": pair_model,
            "pair_seed": ci, "ll": float(ll), "lls": [float(x) for x in lls],
        })
        print(ci, n_leaves, len(names), ll)
    with open(os.path.join(OUT, "cases.json"), "w") as f:
        json.dump(cases, f)


if __name__ == "__main__":
    main()
