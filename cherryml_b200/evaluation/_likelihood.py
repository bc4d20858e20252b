"""``compute_log_likelihoods`` with the reference's signature, caching behaviour and output
files (``cherryml/evaluation/_likelihood.py:436-590``), and ``dp_likelihood_computation``
(:47-326) on the GPU: the transition matrices of all edges come from ``cherry_expm_batched``
(one matrix per DISTINCT ``branch length * site rate``), the dynamic programme is
``cherry_tree_log_likelihood`` (one launch per model: independent sites, contacting pairs).

No CPU fallback: the functions raise if the CUDA library or a GPU is missing.
"""
import os
import time
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from .. import _lib
from ..caching import cached_parallel_computation
from ..io import (Tree, read_contact_map, read_msa, read_probability_distribution, read_rate_matrix, read_site_rates,
                  read_tree, write_log_likelihood)
from ..markov_chain import expm_batched

_PAIR_EXPM_CHUNK = 128  # 400 x 400 exponentials per cherry_expm_batched call (bounds its workspace)


def _tree_arrays(tree: Tree):
    """Post-order node table for the kernel + per node branch length (one iterative walk; the
    order is ``tree.postorder_traversal()``'s: children in edge order, then their parent)."""
    kids_of = {v: tree.children(v) for v in tree.nodes()}
    depths: List[int] = []
    flags: List[int] = []
    obs_rows: List[int] = []
    lengths: List[float] = []
    leaves: List[str] = []
    stack = [(tree.root(), 0, 0.0, 0, False)]
    while stack:
        v, d, length, first, expanded = stack.pop()
        kids = kids_of[v]
        if kids and not expanded:
            stack.append((v, d, length, first, True))
            for j in range(len(kids) - 1, -1, -1):
                stack.append((kids[j][0], d + 1, kids[j][1], 2 if j == 0 else 0, False))
            continue
        depths.append(d)
        lengths.append(length)
        if kids:
            flags.append(first)
            obs_rows.append(-1)
        else:
            flags.append(first | 1)
            obs_rows.append(len(leaves))
            leaves.append(v)
    nodes = np.zeros(len(depths), dtype=_lib.LL_NODE_DTYPE)
    nodes["depth"] = depths
    nodes["flags"] = flags
    nodes["obs_row"] = obs_rows
    return nodes, np.array(lengths, dtype=np.float64), leaves, max(depths)


def _encode_leaves(msa: Dict[str, str], leaves: List[str], amino_acids: List[str]) -> np.ndarray:
    """uint8 [n_leaves, n_sites]: index in ``amino_acids``, S for any other character."""
    S = len(amino_acids)
    lut = np.full(256, S, dtype=np.uint8)
    for i, ch in enumerate(amino_acids):
        lut[ord(ch)] = i
    n_sites = len(msa[leaves[0]])
    seqs = [msa[v] for v in leaves]
    if any(len(q) != n_sites for q in seqs):
        raise ValueError("the sequences of the MSA have different lengths")
    flat = np.frombuffer("".join(seqs).encode("latin-1"), dtype=np.uint8)
    return lut[flat].reshape(len(leaves), n_sites)


def _matrices(Q: np.ndarray, exponents: np.ndarray, device) -> Tuple[torch.Tensor, np.ndarray]:
    """One expm per distinct exponent -> (the TRANSPOSED matrices [n_distinct, S, S] on the device, as the
    pruning kernel reads them, and the index of every exponent)."""
    lib = _lib.load()
    uniq, inverse = np.unique(exponents, return_inverse=True)
    S = Q.shape[0]
    chunk = _PAIR_EXPM_CHUNK if S > 32 else 1 << 16
    dev = torch.device(device)
    out = torch.empty((len(uniq), S, S), dtype=torch.float64, device=dev)
    for i in range(0, len(uniq), chunk):
        part = expm_batched(Q, uniq[i: i + chunk], dev)
        with torch.cuda.device(dev):
            _lib.check(lib.cherry_tree_ll_transpose(_lib.ptr(part), part.shape[0], S, _lib.ptr(out[i: i + chunk]),
                                                    _lib.current_stream_ptr()), "cherry_tree_ll_transpose")
    return out, inverse.astype(np.int32)


def _prune(nodes, max_depth, p_index, n_cats, P, obs, unit_cat, pi, S, c, device) -> np.ndarray:
    lib = _lib.load()
    n_units = obs.shape[1]
    dev = torch.device(device)
    with torch.cuda.device(dev):
        d_nodes = torch.from_numpy(nodes.view(np.uint8).reshape(-1)).to(dev)
        d_pidx = torch.from_numpy(np.ascontiguousarray(p_index, dtype=np.int32)).to(dev)
        d_obs = torch.from_numpy(np.ascontiguousarray(obs, dtype=np.uint8)).to(dev)
        d_cat = torch.from_numpy(np.ascontiguousarray(unit_cat, dtype=np.int32)).to(dev)
        d_pi = torch.from_numpy(np.array(pi, dtype=np.float64).reshape(-1)).to(dev)
        nbytes = int(lib.cherry_tree_ll_scratch_bytes(S, c, n_units, max_depth))
        scratch = torch.empty(max(8, nbytes), dtype=torch.uint8, device=dev)
        out = torch.empty(n_units, dtype=torch.float64, device=dev)
        _lib.check(
            lib.cherry_tree_log_likelihood(_lib.ptr(d_nodes), len(nodes), _lib.ptr(d_pidx), n_cats, _lib.ptr(P),
                                           _lib.ptr(d_obs), _lib.ptr(d_cat), _lib.ptr(d_pi), S, c, n_units,
                                           max_depth, _lib.ptr(scratch), nbytes, _lib.ptr(out),
                                           _lib.current_stream_ptr()),
            "cherry_tree_log_likelihood",
        )
        return out.cpu().numpy()


def dp_likelihood_computation(
    tree: Tree,
    msa: Dict[str, str],
    contact_map: Optional[np.ndarray],
    site_rates: List[float],
    amino_acids: List[str],
    pi_1: np.ndarray,
    Q_1: np.ndarray,
    fact_1=None,
    reversible_1: bool = False,
    device_1: str = "cuda",
    pi_2: Optional[np.ndarray] = None,
    Q_2: Optional[np.ndarray] = None,
    fact_2=None,
    reversible_2: Optional[bool] = None,
    device_2: Optional[str] = None,
    output_profiling_path: Optional[str] = None,
) -> Tuple[float, List[float]]:
    """Data log-likelihood and its per-site split, like the reference function.  ``fact_*``,
    ``reversible_*`` select among the reference's expm back ends, which all compute the same
    matrices; here they are accepted and ignored.  ``device_*`` other than a CUDA device name
    means "cuda"."""
    st_all = time.time()
    S = len(amino_acids)
    num_sites = len(site_rates)
    if contact_map is not None:
        ii, jj = np.where(contact_map == 1)
        pairs = [(int(i), int(j)) for i, j in zip(ii, jj) if i < j]
    else:
        pairs = []
    flat = [s for p in pairs for s in p]
    if len(set(flat)) != len(flat):
        raise Exception(
            f"Each site can only be in contact with one other site. The contacting sites were: {pairs}"
        )
    in_contact = set(flat)
    independent = [i for i in range(num_sites) if i not in in_contact]
    device = device_1 if str(device_1).startswith("cuda") else "cuda"

    nodes, lengths, leaves, max_depth = _tree_arrays(tree)
    enc = _encode_leaves(msa, leaves, amino_acids)
    lls = [0] * num_sites
    t_expm = t_dp = 0.0
    if independent:
        t0 = time.time()
        cats = sorted(set(site_rates))
        cat_of = {r: k for k, r in enumerate(cats)}
        exps = (lengths[:, None] * np.array(cats)[None, :]).reshape(-1)  # length * site_rate, as the reference
        P, p_index = _matrices(np.asarray(Q_1, dtype=np.float64), exps, device)
        t_expm += time.time() - t0
        t0 = time.time()
        res = _prune(nodes, max_depth, p_index, len(cats), P, enc[:, independent][:, :, None],
                     [cat_of[site_rates[i]] for i in independent], pi_1, S, 1, device)
        t_dp += time.time() - t0
        for k, i in enumerate(independent):
            lls[i] = float(res[k])
    if pairs:
        t0 = time.time()
        P, p_index = _matrices(np.asarray(Q_2, dtype=np.float64), lengths.copy(), device)
        t_expm += time.time() - t0
        t0 = time.time()
        obs = np.stack([enc[:, [i for i, _ in pairs]], enc[:, [j for _, j in pairs]]], axis=2)
        res = _prune(nodes, max_depth, p_index, 1, P, obs, [0] * len(pairs), pi_2, S, 2, device)
        t_dp += time.time() - t0
        for k, (i, j) in enumerate(pairs):
            lls[i] = float(res[k]) / 2.0
            lls[j] = float(res[k]) / 2.0
    if output_profiling_path is not None:
        with open(output_profiling_path, "w") as f:
            f.write(f"Time to populate_transition_mats: {t_expm}\nTime for dp: {t_dp}\n"
                    f"Total time: {time.time() - st_all}\n")
    return sum(lls), lls


@cached_parallel_computation(
    parallel_arg="families",
    exclude_args=[
        "device_1",
        "device_2",
        "num_processes",
        "use_cpp_implementation",
        "OMP_NUM_THREADS",
        "OPENBLAS_NUM_THREADS",
        "process_group",
    ],
    output_dirs=["output_likelihood_dir"],
    write_extra_log_files=True,
)
def compute_log_likelihoods(
    tree_dir: str,
    msa_dir: str,
    site_rates_dir: str,
    contact_map_dir: Optional[str],
    families: List[str],
    amino_acids: List[str],
    pi_1_path: str,
    Q_1_path: str,
    reversible_1: bool,
    device_1: str,
    pi_2_path: Optional[str],
    Q_2_path: Optional[str],
    reversible_2: Optional[bool],
    device_2: Optional[str],
    output_likelihood_dir: Optional[str],
    num_processes: int,
    use_cpp_implementation: bool = False,
    OMP_NUM_THREADS: Optional[int] = 1,
    OPENBLAS_NUM_THREADS: Optional[int] = 1,
    process_group=None,
) -> None:
    """Per family ``<output_likelihood_dir>/<family>.txt`` (total, then the per-site values) and
    ``<family>.profiling``.  Model validation as in the reference (:377-416).

    ``process_group`` (a torch.distributed group, one process per GPU): families are independent,
    so rank r evaluates ``families[r::world]`` -- the striping of the reference's worker processes
    (``get_process_args``, :565-575) -- and a barrier makes every file exist before any rank
    returns.  No data-path collective."""
    if use_cpp_implementation:
        raise NotImplementedError
    rank, world = 0, 1
    if process_group is not None:
        import torch.distributed as dist

        rank, world = dist.get_rank(process_group), dist.get_world_size(process_group)
        try:
            _compute_log_likelihoods_local(
                tree_dir, msa_dir, site_rates_dir, contact_map_dir, families[rank::world], amino_acids, pi_1_path,
                Q_1_path, reversible_1, device_1, pi_2_path, Q_2_path, reversible_2, device_2, output_likelihood_dir,
                write_total=rank == 0)
        finally:
            dist.barrier(process_group)
        return
    _compute_log_likelihoods_local(
        tree_dir, msa_dir, site_rates_dir, contact_map_dir, families, amino_acids, pi_1_path, Q_1_path, reversible_1,
        device_1, pi_2_path, Q_2_path, reversible_2, device_2, output_likelihood_dir, write_total=True)


def _compute_log_likelihoods_local(tree_dir, msa_dir, site_rates_dir, contact_map_dir, families, amino_acids,
                                   pi_1_path, Q_1_path, reversible_1, device_1, pi_2_path, Q_2_path, reversible_2,
                                   device_2, output_likelihood_dir, write_total: bool) -> None:
    os.makedirs(output_likelihood_dir, exist_ok=True)
    st = time.time()
    pi_1_df = read_probability_distribution(pi_1_path)
    Q_1_df = read_rate_matrix(Q_1_path)
    pi_2_df = read_probability_distribution(pi_2_path) if pi_2_path is not None else None
    Q_2_df = read_rate_matrix(Q_2_path) if Q_2_path is not None else None
    pairs_of_amino_acids = [a + b for a in amino_acids for b in amino_acids]
    if list(pi_1_df.index) != amino_acids:
        raise Exception(f"pi_1 index is:\n{list(pi_1_df.index)}\nbut expected amino acids:\n{amino_acids}")
    if pi_2_df is not None and list(pi_2_df.index) != pairs_of_amino_acids:
        raise Exception(
            f"pi_2 index is:\n{list(pi_2_df.index)}\nbut expected pairs of amino acids:\n{pairs_of_amino_acids}")
    if list(Q_1_df.index) != amino_acids:
        raise Exception(f"Q_1 index is:\n{list(Q_1_df.index)}\n\nbut expected amino acids:\n{amino_acids}")
    if list(Q_1_df.columns) != amino_acids:
        raise Exception(f"Q_1 columns are:\n{list(Q_1_df.columns)}\n\nbut expected amino acids:\n{amino_acids}")
    if Q_2_df is not None and list(Q_2_df.index) != pairs_of_amino_acids:
        raise Exception(
            f"Q_2 index is:\n{list(Q_2_df.index)}\n\nbut expected pairs of amino acids:\n{pairs_of_amino_acids}")
    if Q_2_df is not None and list(Q_2_df.columns) != pairs_of_amino_acids:
        raise Exception(
            f"Q_1 columns are:\n{list(Q_2_df.columns)}\n\nbut expected pairs of amino acids:\n{pairs_of_amino_acids}")
    for family in families:
        contact_map = (read_contact_map(os.path.join(contact_map_dir, family + ".txt"))
                       if contact_map_dir is not None else None)
        ll, lls = dp_likelihood_computation(
            tree=read_tree(os.path.join(tree_dir, family + ".txt")),
            msa=read_msa(os.path.join(msa_dir, family + ".txt")),
            contact_map=contact_map,
            site_rates=read_site_rates(os.path.join(site_rates_dir, family + ".txt")),
            amino_acids=amino_acids,
            pi_1=pi_1_df.to_numpy(), Q_1=Q_1_df.to_numpy(), reversible_1=reversible_1, device_1=device_1,
            pi_2=pi_2_df.to_numpy() if pi_2_df is not None else None,
            Q_2=Q_2_df.to_numpy() if Q_2_df is not None else None, reversible_2=reversible_2, device_2=device_2,
            output_profiling_path=os.path.join(output_likelihood_dir, family + ".profiling"),
        )
        write_log_likelihood((ll, lls), os.path.join(output_likelihood_dir, family + ".txt"))
    if write_total:
        with open(os.path.join(output_likelihood_dir, "profiling.txt"), "w") as f:
            f.write(f"Total time: {time.time() - st}\n")
