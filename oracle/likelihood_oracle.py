"""CPU restatement of the reference's tree log-likelihood (TEST INFRASTRUCTURE).

Follows ``cherryml/evaluation/_likelihood.py:47-326`` (``dp_likelihood_computation``): sites
split into independent sites (model 1: pi_1, Q_1, per-site rate) and contacting pairs (model
2 on S*S states, rate 1); leaves observe a one-hot vector (all ones for a character outside
the alphabet, the matching 20 states when one site of a pair is unknown); Felsenstein pruning
in log space where the message of child c to its parent is
``log(max(0, expm(t_c * rate * Q) @ (exp(dp_c - max dp_c) * obs_c))) + max dp_c`` and a node's
dp is the sum of its children's messages in ``tree.children`` order; the root combines with
pi the same way; a pair's log-likelihood is split in halves over its two sites.

The matrix exponential is third-party in the reference (torch.matrix_exp or an
eigendecomposition through numpy, ``markov_chain/_markov_chain.py:22-155``); scipy.linalg.expm
here.  PINNED against the FastTree-verified constants of the reference's own tests
(tests/evaluation_tests/likelihood_test.py:236-431, 908-953, restated in
tests/test_oracle_likelihood.py with the reference's 1a92_1_A fixture) and against outputs of
the UNMODIFIED reference function (tests/golden/likelihood/*.npz, made by
tests/golden/make_golden_likelihood.py).
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np


def _expm_stack(Q: np.ndarray, exponents: Sequence[float]) -> np.ndarray:
    from scipy.linalg import expm

    return np.stack([expm(t * Q) for t in exponents]) if len(exponents) else np.zeros((0,) + Q.shape)


def split_sites(contact_map: Optional[np.ndarray], num_sites: int):
    """_likelihood.py:82-97."""
    if contact_map is not None:
        ii, jj = np.where(contact_map == 1)
        pairs = [(int(i), int(j)) for i, j in zip(ii, jj) if i < j]
    else:
        pairs = []
    flat = [s for p in pairs for s in p]
    if len(set(flat)) != len(flat):
        raise Exception(f"Each site can only be in contact with one other site. The contacting sites were: {pairs}")
    independent = [i for i in range(num_sites) if i not in set(flat)]
    return independent, pairs


def log_likelihood(tree, msa: Dict[str, str], contact_map: Optional[np.ndarray], site_rates: List[float],
                   amino_acids: List[str], pi_1: np.ndarray, Q_1: np.ndarray, pi_2: Optional[np.ndarray],
                   Q_2: Optional[np.ndarray]) -> Tuple[float, List[float]]:
    S = len(amino_acids)
    aa = {a: i for i, a in enumerate(amino_acids)}
    num_sites = len(site_rates)
    independent, pairs = split_sites(contact_map, num_sites)

    def obs_single(ch):
        v = np.zeros(S)
        if ch in aa:
            v[aa[ch]] = 1.0
        else:
            v[:] = 1.0
        return v

    def obs_pair(c1, c2):
        m = np.outer(obs_single(c1), obs_single(c2))  # state index = S * i + j
        return m.reshape(-1)

    nodes = tree.nodes()
    root = tree.root()
    non_root = [v for v in nodes if not tree.is_root(v)]
    cats = sorted(set(site_rates))
    cat_of = {r: c for c, r in enumerate(cats)}
    obs1, obs2 = {}, {}
    for v in nodes:
        if tree.is_leaf(v):
            seq = msa[v]
            obs1[v] = np.stack([obs_single(seq[i]) for i in independent]) if independent else np.zeros((0, S))
            obs2[v] = np.stack([obs_pair(seq[i], seq[j]) for i, j in pairs]) if pairs else np.zeros((0, S * S))
        else:
            obs1[v] = np.ones((len(independent), S))
            obs2[v] = np.ones((len(pairs), S * S))
    P1, P2 = {}, {}
    if independent:
        exps = [tree.parent(v)[1] * r for v in non_root for r in cats]
        E = _expm_stack(Q_1, exps)
        site_cat = np.array([cat_of[site_rates[i]] for i in independent])
        for k, v in enumerate(non_root):
            P1[v] = E[k * len(cats) + site_cat]  # [n_ind, S, S]
    if pairs:
        E = _expm_stack(Q_2, [tree.parent(v)[1] for v in non_root])
        for k, v in enumerate(non_root):
            P2[v] = E[k][None]

    def prune(obs, P, pi, n_units, n_states):
        dp = {}
        for v in tree.postorder_traversal():
            dp[v] = np.zeros((n_units, n_states))
            if tree.is_leaf(v):
                continue
            for c, _ in tree.children(v):
                mx = dp[c].max(axis=1, keepdims=True)
                arg = np.einsum("uij,uj->ui", np.broadcast_to(P[c], (n_units, n_states, n_states)),
                                np.exp(dp[c] - mx) * obs[c])
                arg[arg < 0] = 0.0
                with np.errstate(divide="ignore"):
                    dp[v] += np.log(arg) + mx
        mx = dp[root].max(axis=1, keepdims=True)
        arg = (np.exp(dp[root] - mx) * obs[root]) @ pi.reshape(-1)
        arg[arg < 0] = 0.0
        with np.errstate(divide="ignore"):
            return np.log(arg) + mx[:, 0]

    lls = [0.0] * num_sites
    if independent:
        r1 = prune(obs1, P1, pi_1, len(independent), S)
        for k, i in enumerate(independent):
            lls[i] = float(r1[k])
    if pairs:
        r2 = prune(obs2, P2, pi_2, len(pairs), S * S)
        for k, (i, j) in enumerate(pairs):
            lls[i] = float(r2[k]) / 2.0
            lls[j] = float(r2[k]) / 2.0
    return sum(lls), lls
