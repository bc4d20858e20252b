"""Shared fixtures of the likelihood tests: the FastTree-verified known answers of the reference's
tests/evaluation_tests/likelihood_test.py (restated here with its 1a92_1_A data fixture under
tests/golden/likelihood/1a92) and the goldens made by tests/golden/make_golden_likelihood.py."""
import json
import os

import numpy as np

from cherryml_b200.io import Tree, read_msa, read_rate_matrix, read_site_rates, read_tree
from cherryml_b200.markov_chain import chain_product, compute_stationary_distribution
from tests.conftest import GOLDEN

AA = list("ARNDCQEGHILKMFPSTWYV")
LL_DIR = os.path.join(GOLDEN, "likelihood")


def rate_matrix(name):
    from cherryml_b200.markov_chain import _rate_matrix_path

    return read_rate_matrix(_rate_matrix_path(name)).to_numpy()


def _tree(nodes, edges):
    t = Tree()
    t.add_nodes(nodes)
    t.add_edges(edges)
    return t


def fasttree_kats():
    """(name, tree, msa, contact_map, site_rates, pi_1, Q_1, pi_2, Q_2, ll, lls, decimals).
    likelihood_test.py:242-284 (3 seqs), :291-331 (4 seqs), :338-378 (gaps), :385-429 (equ x equ),
    :918-953 (1a92_1_A, 1/2/4/20 rate categories)."""
    wag, equ = rate_matrix("wag"), rate_matrix("equ")
    pi_wag = compute_stationary_distribution(wag)
    equ2 = chain_product(equ, equ)
    pi_equ2 = compute_stationary_distribution(equ2)
    out = []
    t3 = _tree(["r", "l1", "l2", "l3"], [("r", "l1", 0.0), ("r", "l2", 1.120547166), ("r", "l3", 3.402392896)])
    out.append(("wag_3_seqs", t3, {"l1": "S", "l2": "T", "l3": "G"}, np.eye(1), [1.0], pi_wag, wag, None, None,
                -7.343870, [-7.343870], 4))
    t4 = _tree(["r", "i1", "l1", "l2", "l3", "l4"],
               [("r", "l1", 0.0), ("r", "l2", 1.121562482), ("r", "i1", 1.719057732), ("i1", "l3", 1.843908633),
                ("i1", "l4", 2.740236263)])
    out.append(("wag_4_seqs_gaps", t4, {"l1": "SS", "l2": "TT", "l3": "GG", "l4": "D-"}, np.eye(2), [1.0, 1.0],
                pi_wag, wag, pi_equ2, equ2, -17.436349, [-10.092142, -7.344207], 4))
    out.append(("equ_x_equ_3_seqs", t3, {"l1": "SK", "l2": "TI", "l3": "GL"}, np.ones((2, 2)), [1.0, 1.0],
                compute_stationary_distribution(equ), equ, pi_equ2, equ2, -9.382765 * 2, [-9.382765, -9.382765], 4))
    d = os.path.join(LL_DIR, "1a92")
    msa = read_msa(os.path.join(d, "msa.txt"))
    for cats, ll in ((1, -4649.6146), (2, -4397.8184), (4, -4337.8688), (20, -4307.0638)):
        rates = read_site_rates(os.path.join(d, f"site_rates_{cats}_cat.txt"))
        out.append((f"1a92_{cats}_cat", read_tree(os.path.join(d, f"tree_{cats}_cat.txt")), msa, np.eye(len(rates)),
                    rates, pi_wag, wag, None, None, ll, None, 4))
    return out


def pair_matrix(case):
    Q1 = rate_matrix(case["Q1"])
    if case["pair_model"] is None:
        return None
    base = chain_product(Q1, Q1)
    if case["pair_model"] == "product":
        return base
    rng = np.random.default_rng(case["pair_seed"])
    pi = compute_stationary_distribution(base)
    n = base.shape[0]
    sym = rng.uniform(0.5, 1.5, (n, n))
    sym = (sym + sym.T) / 2
    Q = (base / pi[None, :]) * sym * pi[None, :]
    Q[np.arange(n), np.arange(n)] = 0
    Q[np.arange(n), np.arange(n)] = -Q.sum(axis=1)
    return Q


def golden_cases():
    with open(os.path.join(LL_DIR, "cases.json")) as f:
        cases = json.load(f)
    out = []
    for c in cases:
        t = Tree()
        t.add_nodes(c["names"])
        for i in range(1, len(c["names"])):
            t.add_edge(c["names"][c["parent"][i]], c["names"][i], c["length"][i])
        L = len(c["site_rates"])
        cmap = None
        if c["contact_map_given"]:
            cmap = np.eye(L) if not c["pairs"] else np.zeros((L, L))
            for a, b in c["pairs"]:
                cmap[a, b] = cmap[b, a] = 1
        Q1 = rate_matrix(c["Q1"])
        Q2 = pair_matrix(c)
        out.append(dict(tree=t, msa=c["msa"], contact_map=cmap, site_rates=c["site_rates"], Q1=Q1,
                        pi1=compute_stationary_distribution(Q1), Q2=Q2,
                        pi2=None if Q2 is None else compute_stationary_distribution(Q2), ll=c["ll"], lls=c["lls"]))
    return out
