"""``cherryml_public_api`` and the two end-to-end pipelines:
[FastCherries trees ->] count -> JTT-IPW -> fit -> rate matrix file.

Same names, keyword arguments and defaults as the reference's
``cherryml/_cherryml_public_api.py:36-252`` and
``cherryml/estimation_end_to_end/_cherry.py:209-445, 449-584``.  Of the reference's three tree
estimators only FastCherries is on the path this package implements (SURVEY.md section 8, row
f3): with ``tree_estimator_name="FastCherries"`` trees and site rates are estimated on the GPU,
iterated ``num_iterations`` times for the LG model like the reference does; FastTree / PhyML
(external programs) raise ``NotImplementedError`` unless ``tree_dir`` (and ``site_rates_dir``
for the LG model) are given.
"""
import logging
import os
import tempfile
import time
from functools import partial
from typing import Dict, List, Optional

import numpy as np

from . import caching
from .counting import count_co_transitions, count_transitions, device_result
from .estimation import jtt_ipw, quantized_transitions_mle
from .io import (read_contact_map, read_msa, read_rate_matrix, read_site_rates, read_sites_subset, write_contact_map,
                 write_msa, write_rate_matrix, write_site_rates)
from .markov_chain import get_equ_path, get_lg_path
from .phylogeny_estimation import fast_cherries
from .utils import get_amino_acids, get_families

logger = logging.getLogger(__name__)
CHERRYML_TYPE = "cherry++"


def _runtime_from_profiling_file(path: str) -> float:
    """Third whitespace token of the first line (reference ``_cherry.py:148-155``)."""
    with open(path) as f:
        return float(f.readline().split()[2])


def _quantization_points(center: float, step: float, num_steps: int) -> List[str]:
    return [("%.8f" % (center * step**i)) for i in range(-num_steps, num_steps + 1, 1)]


def maximal_matching_pairs(contact_map: np.ndarray, minimum_distance: int):
    """Greedy maximal matching over the contacting pairs in row-major order -- what
    ``networkx.maximal_matching`` returns for the graph the reference builds
    (``evaluation/_maximal_matching.py:69-93``: nodes 0..L-1, edges added in sorted order)."""
    ii, jj = np.where(contact_map == 1)
    used = np.zeros(contact_map.shape[0], dtype=bool)
    match = []
    for i, j in zip(ii.tolist(), jj.tolist()):
        if i < j and j - i >= minimum_distance and not used[i] and not used[j]:
            used[i] = used[j] = True
            match.append((i, j))
    return match


@caching.cached_parallel_computation(
    parallel_arg="families",
    exclude_args=["num_processes"],
    output_dirs=["o_contact_map_dir"],
    write_extra_log_files=True,
)
def create_maximal_matching_contact_map(
    i_contact_map_dir: str,
    families: List[str],
    minimum_distance_for_nontrivial_contact: int,
    num_processes: int,
    o_contact_map_dir: Optional[str] = None,
) -> None:
    """Per family, the contact map reduced to a maximal matching of its non-trivial contacts, so
    that every site is in at most one pair (reference ``evaluation/_maximal_matching.py:31-115``;
    cached per family like the reference's stage)."""
    for family in families:
        cmap = read_contact_map(os.path.join(i_contact_map_dir, family + ".txt"))
        res = np.zeros(cmap.shape)
        for u, v in maximal_matching_pairs(cmap, minimum_distance_for_nontrivial_contact):
            res[u, v] = res[v, u] = 1
        write_contact_map(res, os.path.join(o_contact_map_dir, family + ".txt"))


@caching.cached_parallel_computation(
    exclude_args=["num_processes"],
    parallel_arg="families",
    output_dirs=["output_msa_dir", "output_site_rates_dir"],
    write_extra_log_files=True,
)
def _subset_data_to_sites_subset(
    sites_subset_dir: str,
    msa_dir: str,
    site_rates_dir: str,
    families: List[str],
    num_processes: int = 1,
    output_msa_dir: Optional[str] = None,
    output_site_rates_dir: Optional[str] = None,
):
    """MSAs and site rates restricted to the sites listed in ``<sites_subset_dir>/<family>.txt``
    (reference ``estimation_end_to_end/_cherry.py:41-147``)."""
    for family in families:
        sites = read_sites_subset(os.path.join(sites_subset_dir, family + ".txt"))
        msa = read_msa(os.path.join(msa_dir, family + ".txt"))
        rates = read_site_rates(os.path.join(site_rates_dir, family + ".txt"))
        write_msa({name: "".join(seq[i] for i in sites) for name, seq in msa.items()},
                  os.path.join(output_msa_dir, family + ".txt"))
        write_site_rates([rates[i] for i in sites], os.path.join(output_site_rates_dir, family + ".txt"))


def _no_tree_estimator(what: str):
    raise NotImplementedError(
        f"{what}: of the reference's tree estimators only FastCherries is implemented here "
        "(tree_estimator_name=\"FastCherries\"); FastTree and PhyML are external programs -- run "
        "them with the reference and pass tree_dir (and site_rates_dir for the LG model)."
    )


def _tree_estimation_runtime(tree_estimator_output_dirs: Dict, families: List[str], attribute: str) -> float:
    """Sum over families of the `<attribute>_time:` line of ``<family>.profiling`` (reference
    ``_cherry.py:157-191``)."""
    total = 0.0
    for family in families:
        path = os.path.join(tree_estimator_output_dirs["output_tree_dir"], family + ".profiling")
        if not os.path.exists(path):
            continue
        for line in open(path):
            if line.startswith(attribute + "_time"):
                total += float(line.split()[1])
    return total


def lg_end_to_end_with_cherryml_optimizer(
    msa_dir: str,
    families: List[str],
    tree_estimator=None,
    initial_tree_estimator_rate_matrix_path: Optional[str] = None,
    num_iterations: Optional[int] = 1,
    quantization_grid_center: float = 0.03,
    quantization_grid_step: float = 1.1,
    quantization_grid_num_steps: int = 64,
    use_cpp_counting_implementation: bool = True,
    optimizer_device: str = "cpu",
    learning_rate: float = 1e-1,
    num_epochs: int = 2000,
    do_adam: bool = True,
    edge_or_cherry: str = CHERRYML_TYPE,
    cpp_counting_command_line_prefix: str = "",
    cpp_counting_command_line_suffix: str = "",
    num_processes_tree_estimation: int = 8,
    num_processes_counting: int = 8,
    num_processes_optimization: int = 2,
    optimizer_initialization: str = "jtt-ipw",
    sites_subset_dir: Optional[str] = None,
    tree_dir: Optional[str] = None,
    site_rates_dir: Optional[str] = None,
    alphabet: List[str] = get_amino_acids(),
) -> Dict:
    if (tree_dir is None) != (site_rates_dir is None):
        raise ValueError(
            "tree_dir and site_rates_dir must be either both provided or none "
            f"provided. You provided: tree_dir={tree_dir} ; site_rates_dir={site_rates_dir}"
        )
    if tree_estimator is None and (tree_dir is None or num_iterations != 1):
        _no_tree_estimator("lg_end_to_end_with_cherryml_optimizer")
    res: Dict = {}
    quantization_points = _quantization_points(
        quantization_grid_center, quantization_grid_step, quantization_grid_num_steps)
    res["quantization_points"] = quantization_points
    times = dict(tree=0.0, pairing=0.0, ble=0.0, counting=0.0, jtt_ipw=0.0, optimization=0.0)
    is_a_pairer = False
    current_estimate_rate_matrix_path = initial_tree_estimator_rate_matrix_path
    tree_estimator_output_dirs: Dict = {}
    for iteration in range(num_iterations):
        if iteration == 0 and tree_dir is not None and site_rates_dir is not None:
            tree_estimator_output_dirs = {"output_tree_dir": tree_dir, "output_site_rates_dir": site_rates_dir}
        else:
            tree_estimator_output_dirs = tree_estimator(
                msa_dir=msa_dir, families=families, rate_matrix_path=current_estimate_rate_matrix_path,
                num_processes=num_processes_tree_estimation,
            )
            is_a_pairer = True
            times["tree"] += _tree_estimation_runtime(tree_estimator_output_dirs, families, "total")
            times["pairing"] += _tree_estimation_runtime(tree_estimator_output_dirs, families, "pairing")
            times["ble"] += _tree_estimation_runtime(tree_estimator_output_dirs, families, "ble")
        res[f"tree_estimator_output_dirs_{iteration}"] = tree_estimator_output_dirs
        if sites_subset_dir is not None:
            sub = _subset_data_to_sites_subset(
                sites_subset_dir=sites_subset_dir, msa_dir=msa_dir,
                site_rates_dir=tree_estimator_output_dirs["output_site_rates_dir"], families=families,
                num_processes=num_processes_counting,
            )
            msa_dir = sub["output_msa_dir"]
            tree_estimator_output_dirs = dict(tree_estimator_output_dirs,
                                              output_site_rates_dir=sub["output_site_rates_dir"])
        count_matrices_dir = count_transitions(
            tree_dir=tree_estimator_output_dirs["output_tree_dir"], msa_dir=msa_dir,
            site_rates_dir=tree_estimator_output_dirs["output_site_rates_dir"], families=families,
            amino_acids=alphabet[:], quantization_points=quantization_points, edge_or_cherry=edge_or_cherry,
            num_processes=num_processes_counting, use_cpp_implementation=use_cpp_counting_implementation,
            cpp_command_line_prefix=cpp_counting_command_line_prefix,
            cpp_command_line_suffix=cpp_counting_command_line_suffix,
        )["output_count_matrices_dir"]
        res[f"count_matrices_dir_{iteration}"] = count_matrices_dir
        times["counting"] += _runtime_from_profiling_file(os.path.join(count_matrices_dir, "profiling.txt"))
        jtt_ipw_dir = jtt_ipw(
            count_matrices_path=os.path.join(count_matrices_dir, "result.txt"), mask_path=None, use_ipw=True,
            normalize=False,
        )["output_rate_matrix_dir"]
        res[f"jtt_ipw_dir_{iteration}"] = jtt_ipw_dir
        times["jtt_ipw"] += _runtime_from_profiling_file(os.path.join(jtt_ipw_dir, "profiling.txt"))
        if optimizer_initialization == "jtt-ipw":
            initialization_path = os.path.join(jtt_ipw_dir, "result.txt")
        elif optimizer_initialization == "equ":
            initialization_path = get_equ_path()
        elif optimizer_initialization == "random":
            initialization_path = None
        else:
            raise ValueError(f"Unknown optimizer_initialization = {optimizer_initialization}")
        rate_matrix_dir = quantized_transitions_mle(
            count_matrices_path=os.path.join(count_matrices_dir, "result.txt"),
            initialization_path=initialization_path, mask_path=None, stationary_distribution_path=None,
            rate_matrix_parameterization="pande_reversible", device=optimizer_device,
            learning_rate=learning_rate, num_epochs=num_epochs, do_adam=do_adam,
            OMP_NUM_THREADS=num_processes_optimization, OPENBLAS_NUM_THREADS=num_processes_optimization,
        )["output_rate_matrix_dir"]
        res[f"rate_matrix_dir_{iteration}"] = rate_matrix_dir
        times["optimization"] += _runtime_from_profiling_file(os.path.join(rate_matrix_dir, "profiling.txt"))
        current_estimate_rate_matrix_path = os.path.join(rate_matrix_dir, "result.txt")
    res["learned_rate_matrix_path"] = current_estimate_rate_matrix_path
    res["all_site_rates"] = [
        read_site_rates(os.path.join(tree_estimator_output_dirs["output_site_rates_dir"], family + ".txt"))
        for family in sorted(families)
    ]
    res["time_tree_estimation"] = times["tree"]
    if is_a_pairer:
        res["time_pairing"] = times["pairing"]
        res["time_ble"] = times["ble"]
    res["time_counting"] = times["counting"]
    res["time_jtt_ipw"] = times["jtt_ipw"]
    res["time_optimization"] = times["optimization"]
    res["total_cpu_time"] = times["tree"] + times["counting"] + times["jtt_ipw"] + times["optimization"]
    res["profiling_str"] = (
        "CherryML runtimes:\n"
        f"time_tree_estimation (without parallelization): {res['time_tree_estimation']}\n"
        f"time_counting: {res['time_counting']}\n"
        f"time_jtt_ipw: {res['time_jtt_ipw']}\n"
        f"time_optimization: {res['time_optimization']}\n"
        f"total_cpu_time: {res['total_cpu_time']}\n"
    )
    if is_a_pairer:
        res["profiling_str"] += f"time_pairing {res['time_pairing']}\ntime_ble {res['time_ble']}"
    return res


def coevolution_end_to_end_with_cherryml_optimizer(
    msa_dir: str,
    contact_map_dir: str,
    minimum_distance_for_nontrivial_contact: int,
    coevolution_mask_path: Optional[str],
    families: List[str],
    tree_estimator=None,
    initial_tree_estimator_rate_matrix_path: Optional[str] = None,
    quantization_grid_center: float = 0.03,
    quantization_grid_step: float = 1.1,
    quantization_grid_num_steps: int = 64,
    use_cpp_counting_implementation: bool = True,
    optimizer_device: str = "cpu",
    learning_rate: float = 1e-1,
    num_epochs: int = 500,
    do_adam: bool = True,
    edge_or_cherry: str = CHERRYML_TYPE,
    cpp_counting_command_line_prefix: str = "",
    cpp_counting_command_line_suffix: str = "",
    num_processes_tree_estimation: int = 8,
    num_processes_counting: int = 8,
    num_processes_optimization: int = 8,
    optimizer_initialization: str = "jtt-ipw",
    use_maximal_matching: bool = True,
    tree_dir: Optional[str] = None,
    alphabet: List[str] = get_amino_acids(),
) -> Dict:
    if tree_dir is None and tree_estimator is None:
        _no_tree_estimator("coevolution_end_to_end_with_cherryml_optimizer")
    res: Dict = {}
    quantization_points = _quantization_points(
        quantization_grid_center, quantization_grid_step, quantization_grid_num_steps)
    res["quantization_points"] = quantization_points
    if tree_dir is None:
        tree_dir = tree_estimator(
            msa_dir=msa_dir, families=families, rate_matrix_path=initial_tree_estimator_rate_matrix_path,
            num_processes=num_processes_tree_estimation,
        )["output_tree_dir"]
    res["tree_estimator_output_dirs_0"] = {"output_tree_dir": tree_dir}
    mdnc = minimum_distance_for_nontrivial_contact
    if use_maximal_matching:
        contact_map_dir = create_maximal_matching_contact_map(
            i_contact_map_dir=contact_map_dir, families=families,
            minimum_distance_for_nontrivial_contact=mdnc, num_processes=num_processes_counting,
        )["o_contact_map_dir"]
    count_matrices_dir = count_co_transitions(
        tree_dir=tree_dir, msa_dir=msa_dir, contact_map_dir=contact_map_dir, families=families,
        amino_acids=alphabet[:], quantization_points=quantization_points, edge_or_cherry=edge_or_cherry,
        minimum_distance_for_nontrivial_contact=mdnc, num_processes=num_processes_counting,
        use_cpp_implementation=use_cpp_counting_implementation,
        cpp_command_line_prefix=cpp_counting_command_line_prefix,
        cpp_command_line_suffix=cpp_counting_command_line_suffix,
    )["output_count_matrices_dir"]
    res["count_matrices_dir_0"] = count_matrices_dir
    jtt_ipw_dir = jtt_ipw(
        count_matrices_path=os.path.join(count_matrices_dir, "result.txt"), mask_path=coevolution_mask_path,
        use_ipw=True, normalize=False,
    )["output_rate_matrix_dir"]
    res["jtt_ipw_dir_0"] = jtt_ipw_dir
    if optimizer_initialization == "jtt-ipw":
        initialization_path = os.path.join(jtt_ipw_dir, "result.txt")
    elif optimizer_initialization == "equ_x_equ":  # reference _cherry.py: the product of two uniform chains
        from .markov_chain import get_equ_x_equ_path

        initialization_path = get_equ_x_equ_path()
    elif optimizer_initialization == "random":
        initialization_path = None
    else:
        raise ValueError(f"Unknown optimizer_initialization = {optimizer_initialization}")
    rate_matrix_dir = quantized_transitions_mle(
        count_matrices_path=os.path.join(count_matrices_dir, "result.txt"),
        initialization_path=initialization_path, mask_path=coevolution_mask_path,
        stationary_distribution_path=None, rate_matrix_parameterization="pande_reversible",
        device=optimizer_device, learning_rate=learning_rate, num_epochs=num_epochs, do_adam=do_adam,
        OMP_NUM_THREADS=num_processes_optimization, OPENBLAS_NUM_THREADS=num_processes_optimization,
    )["output_rate_matrix_dir"]
    res["rate_matrix_dir_0"] = rate_matrix_dir
    res["learned_rate_matrix_path"] = os.path.join(rate_matrix_dir, "result.txt")
    return res


def cherryml_public_api(
    output_path: str,
    model_name: str,
    msa_dir: str,
    contact_map_dir: Optional[str] = None,
    tree_dir: Optional[str] = None,
    site_rates_dir: Optional[str] = None,
    cache_dir: Optional[str] = None,
    num_processes_tree_estimation: int = 32,
    num_processes_counting: int = 8,
    num_processes_optimization: int = 2,
    num_rate_categories: int = 20,
    initial_tree_estimator_rate_matrix_path: Optional[str] = None,
    num_iterations: int = 1,
    quantization_grid_center: float = 0.03,
    quantization_grid_step: float = 1.1,
    quantization_grid_num_steps: int = 64,
    use_cpp_counting_implementation: bool = True,
    optimizer_device: str = "cpu",
    learning_rate: float = 1e-1,
    num_epochs: int = 500,
    minimum_distance_for_nontrivial_contact: int = 7,
    do_adam: bool = True,
    cherryml_type: str = "cherry++",
    cpp_counting_command_line_prefix: str = "",
    cpp_counting_command_line_suffix: str = "",
    optimizer_initialization: str = "jtt-ipw",
    sites_subset_dir: Optional[str] = None,
    coevolution_mask_path: Optional[str] = None,
    use_maximal_matching: bool = True,
    families: Optional[List[str]] = None,
    tree_estimator_name: str = "FastTree",
) -> None:
    """Estimate a rate matrix (LG 20x20 or co-evolution 400x400) and write it to
    ``output_path``.  See the module docstring for what is (not) supported."""
    if model_name not in ["LG", "co-evolution"]:
        raise ValueError('model_name should be either "LG" or "co-evolution".')
    if cache_dir is None:
        # the reference stores a TemporaryDirectory object here and then fails; use a real path
        cache_dir = tempfile.mkdtemp(prefix="cherryml_b200_cache_")
        logger.info(f"Cache directory not provided. Will use temporary directory {cache_dir}.")
    caching.set_cache_dir(cache_dir)
    if families is None:
        families = get_families(msa_dir)
    if initial_tree_estimator_rate_matrix_path is None or (
        # the reference's default is this path relative to ITS checkout; callers that pass the
        # literal along (its CLI does) mean "the LG matrix"
        initial_tree_estimator_rate_matrix_path == "data/rate_matrices/lg.txt"
        and not os.path.exists(initial_tree_estimator_rate_matrix_path)
    ):
        initial_tree_estimator_rate_matrix_path = get_lg_path()
    if tree_estimator_name == "FastCherries":
        # `verbose` stays at its default (True) as in the reference's partial: it is part of the cache key
        tree_estimator = partial(fast_cherries, max_iters=50, num_rate_categories=num_rate_categories)
    elif tree_estimator_name in ("FastTree", "PhyML"):
        tree_estimator = None  # external programs: only usable here with tree_dir given
    else:
        raise ValueError(f"Unknown tree_estimator_name: {tree_estimator_name}")
    if model_name == "LG":
        outputs = lg_end_to_end_with_cherryml_optimizer(
            msa_dir=msa_dir, families=families, tree_estimator=tree_estimator,
            initial_tree_estimator_rate_matrix_path=initial_tree_estimator_rate_matrix_path,
            num_iterations=num_iterations, quantization_grid_center=quantization_grid_center,
            quantization_grid_step=quantization_grid_step,
            quantization_grid_num_steps=quantization_grid_num_steps,
            use_cpp_counting_implementation=use_cpp_counting_implementation,
            optimizer_device=optimizer_device, learning_rate=learning_rate, num_epochs=num_epochs,
            do_adam=do_adam, edge_or_cherry=cherryml_type,
            cpp_counting_command_line_prefix=cpp_counting_command_line_prefix,
            cpp_counting_command_line_suffix=cpp_counting_command_line_suffix,
            num_processes_tree_estimation=num_processes_tree_estimation,
            num_processes_counting=num_processes_counting,
            num_processes_optimization=num_processes_optimization,
            optimizer_initialization=optimizer_initialization, sites_subset_dir=sites_subset_dir,
            tree_dir=tree_dir, site_rates_dir=site_rates_dir,
        )
    else:
        if num_iterations > 1:
            raise ValueError(
                "Iteration is not used for learning a coevolution model. "
                f"You provided: num_iterations={num_iterations}. Set this argument to 1 and retry."
            )
        outputs = coevolution_end_to_end_with_cherryml_optimizer(
            msa_dir=msa_dir, contact_map_dir=contact_map_dir,
            minimum_distance_for_nontrivial_contact=minimum_distance_for_nontrivial_contact,
            coevolution_mask_path=coevolution_mask_path, families=families, tree_estimator=tree_estimator,
            initial_tree_estimator_rate_matrix_path=initial_tree_estimator_rate_matrix_path,
            quantization_grid_center=quantization_grid_center,
            quantization_grid_step=quantization_grid_step,
            quantization_grid_num_steps=quantization_grid_num_steps,
            use_cpp_counting_implementation=use_cpp_counting_implementation,
            optimizer_device=optimizer_device, learning_rate=learning_rate, num_epochs=num_epochs,
            do_adam=do_adam, edge_or_cherry=cherryml_type,
            cpp_counting_command_line_prefix=cpp_counting_command_line_prefix,
            cpp_counting_command_line_suffix=cpp_counting_command_line_suffix,
            num_processes_tree_estimation=num_processes_tree_estimation,
            num_processes_counting=num_processes_counting,
            num_processes_optimization=num_processes_optimization,
            optimizer_initialization=optimizer_initialization,
            use_maximal_matching=use_maximal_matching, tree_dir=tree_dir,
        )
    learned = read_rate_matrix(outputs["learned_rate_matrix_path"])
    write_rate_matrix(learned.to_numpy(), list(learned.columns), output_path)
