"""SiteRM: site-specific rate matrices from a tree, site rates and an MSA.

``estimate_site_specific_rate_matrices_given_tree_and_site_rates`` keeps the arguments and
the returned dictionary of the reference's
``_estimate_site_specific_rate_matrices_given_tree_and_site_rates``
(``cherryml/_siterm/_site_specific_rate_matrix.py:442-731``) and runs its four stages on the
GPU: per-site counts (``cherry_count_per_site``), the prior matrices ``pi_x expm(t Q0)[x, y]``
(``cherry_expm_batched``), pseudocounts / regularised counts / compaction (tensor ops on the
resident count tensor) and the batched per-site fit (``cherry_fit_*`` with one problem per
site).  The reference's non-vectorised branch (one ``quantized_transitions_mle`` per site)
solves the same problems; here both values of ``use_vectorized_cherryml_implementation`` take
the batched path.
"""
import time
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from ..io import Tree
from ..markov_chain import compute_stationary_distribution, expm_batched
from ..utils import quantization_idx_array
from ._counting import get_raw_count_matrices_device
from ._vectorized import quantized_transitions_mle_vectorized_over_sites


def get_cherry_transitions(tree: Tree, msa: Dict[str, str]) -> List[Tuple[str, str, float]]:
    """(sequence, sequence, path length) of the ``cherry++`` pairs: post-order, every node pairs
    the unmatched leaves handed up by its children left to right and hands up the odd one
    (reference ``_site_specific_rate_matrix.py:87-139``).  Iterative, so deep trees are fine."""
    cherries: List[Tuple[str, str, float]] = []
    handed_up: Dict[str, Optional[Tuple[str, float]]] = {}
    for node in tree.postorder_traversal():
        if tree.is_leaf(node):
            handed_up[node] = (node, 0.0)
            continue
        pending = []
        for child, branch_length in tree.children(node):
            up = handed_up.pop(child)
            if up is not None:
                pending.append((up[0], up[1] + branch_length))
        for k in range(0, len(pending) - 1, 2):
            (leaf_1, d1), (leaf_2, d2) = pending[k], pending[k + 1]
            cherries.append((msa[leaf_1], msa[leaf_2], d1 + d2))
        handed_up[node] = pending[-1] if len(pending) % 2 == 1 else None
    assert len(cherries) == int(len(tree.leaves()) / 2)
    return cherries


def get_edge_transitions(tree: Tree, msa: Dict[str, str]) -> List[Tuple[str, str, float]]:
    assert sorted(tree.nodes()) == sorted(msa.keys())
    return [(msa[u], msa[v], t) for (u, v, t) in tree.edges()]


def count_prior_probability_matrices(rate_matrix: np.ndarray, quantization_points_sorted: List[float],
                                     device="cuda") -> torch.Tensor:
    """``[B, S, S]`` with entry ``pi[x] * expm(t_b Q0)[x, y]`` (reference :325-356)."""
    pi = compute_stationary_distribution(rate_matrix)
    P = expm_batched(rate_matrix, quantization_points_sorted, device)
    out = torch.from_numpy(np.asarray(pi, dtype=np.float64)).to(P.device)[None, :, None] * P
    sums = out.sum(dim=(1, 2))
    if bool((torch.abs(sums - 1.0) > 1e-6).any()):
        raise ValueError("count_prior_probability_matrices[b, :, :] does not add up to 1!")
    return out


def estimate_site_specific_rate_matrices_given_tree_and_site_rates(
    tree: Tree,
    site_rates: List[float],
    msa: Dict[str, str],
    alphabet: List[str],
    regularization_strength: float,
    regularization_rate_matrix: np.ndarray,
    quantization_points: List[float],
    optimization_num_epochs: int,
    transitions_strategy: str = "cherry++",
    include_reverse_transitions: bool = True,
    rate_matrix_parameterization: str = "pande_reversible",
    log_dir: Optional[str] = None,
    plot_site_specific_rate_matrices: int = 0,
    use_vectorized_cherryml_implementation: bool = True,
    vectorized_cherryml_implementation_device: str = "cuda",
    vectorized_cherryml_implementation_num_cores: int = 1,
) -> Dict:
    if rate_matrix_parameterization != "pande_reversible":
        raise NotImplementedError("only the pande_reversible parameterisation is implemented")
    prof: Dict = {}
    st = time.time()
    dev = vectorized_cherryml_implementation_device
    dev = dev if str(dev).startswith("cuda") else "cuda"
    q = sorted(float(x) for x in quantization_points)
    Q0 = np.asarray(regularization_rate_matrix, dtype=np.float64)
    if transitions_strategy == "cherry++":
        assert sorted(tree.leaves()) == sorted(msa.keys())
        transitions = get_cherry_transitions(tree, msa)
    elif transitions_strategy == "edges":
        transitions = get_edge_transitions(tree, msa)
    else:
        raise ValueError(f"Unknown transitions_strategy: {transitions_strategy}")
    L, B, S = len(transitions[0][0]), len(q), len(alphabet)
    prof["time_get_transitions"] = time.time() - st

    st = time.time()
    raw = get_raw_count_matrices_device(transitions, q, alphabet, include_reverse_transitions, dev)  # [L,B,S,S]
    prof["time_get_raw_count_matrices"] = time.time() - st
    st = time.time()
    prior = count_prior_probability_matrices(Q0, q, dev)
    prof["time_get_count_prior_probability_matrices"] = time.time() - st

    st = time.time()
    # bucket of t_b * site_rate (clamped to the grid's ends, reference :519-545)
    scaled = np.asarray(site_rates, dtype=np.float64)[:, None] * np.asarray(q)[None, :]  # t_b * rate_l
    b_adj = quantization_idx_array(scaled, q)
    b_adj = np.where(b_adj >= 0, b_adj, np.where(scaled > q[-1], B - 1, 0)).astype(np.int64)
    l1 = raw.sum(dim=(2, 3))  # [L, B]
    pseudo = l1[:, :, None, None] * prior[torch.from_numpy(b_adj).to(raw.device)]
    raw_sum, pseudo_sum = float(raw.sum()), float(pseudo.sum())
    if abs(raw_sum - pseudo_sum) > 0.4 and abs(raw_sum / pseudo_sum - 1.0) > 1e-6:
        raise ValueError(f"Raw counts matrix and pseudocounts matrix have different counts: {raw_sum} vs {pseudo_sum}")
    prof["time_get_pseudocount_matrices"] = time.time() - st
    st = time.time()
    counts = raw * (1.0 - regularization_strength) + pseudo * regularization_strength
    prof["time_get_count_matrices"] = time.time() - st

    st = time.time()
    rates = np.asarray(site_rates, dtype=np.float64)
    initialization = Q0[None, :, :] * rates[:, None, None]
    # keep, per site, the buckets that hold counts (in grid order); pad with (t = 1, zero counts)
    keep = counts.sum(dim=(2, 3)) > 0
    n_keep = keep.sum(dim=1)
    B_eff = int(n_keep.max().item()) if L else 0
    order = torch.argsort((~keep).to(torch.int8), dim=1, stable=True)[:, :B_eff]  # kept buckets first, in order
    valid = torch.arange(B_eff, device=counts.device)[None, :] < n_keep[:, None]
    compact = torch.gather(counts, 1, order[:, :, None, None].expand(L, B_eff, S, S)) * valid[:, :, None, None]
    q_dev = torch.tensor(q, dtype=torch.float64, device=counts.device)
    times = torch.where(valid, q_dev[order], torch.ones_like(q_dev[order]))
    prof["time_get_count_matrices_compactified"] = time.time() - st
    if B_eff == 0:
        res = {"res": initialization}
        return {**res, **prof}
    fit = quantized_transitions_mle_vectorized_over_sites(
        counts=compact.cpu().numpy(), times=times.cpu().numpy(), num_epochs=optimization_num_epochs,
        initialization=initialization, device=dev, num_cores=vectorized_cherryml_implementation_num_cores,
    )
    for k, v in fit.items():
        if k.startswith("time_"):
            prof[k] = v
    prof["time_plotting"] = 0.0
    return {"res": fit["res"], **prof}
