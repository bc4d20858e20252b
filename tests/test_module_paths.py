"""``import cherryml_b200 as cherryml`` serves the reference's module paths for everything on or
next to the hot path (reference cherryml/__init__.py and the sub-packages' __init__.py)."""
import importlib

import pytest

PATHS = [
    "cherryml_public_api", "learn_site_specific_rate_matrices", "count_transitions", "count_co_transitions",
    "jtt_ipw", "quantized_transitions_mle", "lg_end_to_end_with_cherryml_optimizer",
    "coevolution_end_to_end_with_cherryml_optimizer", "compute_log_likelihoods", "PhylogenyEstimatorType", "caching",
    "counting.count_transitions", "counting.count_co_transitions", "estimation.quantized_transitions_mle",
    "estimation.jtt_ipw", "estimation_end_to_end.lg_end_to_end_with_cherryml_optimizer",
    "estimation_end_to_end.coevolution_end_to_end_with_cherryml_optimizer", "estimation_end_to_end.CHERRYML_TYPE",
    "phylogeny_estimation.fast_cherries", "phylogeny_estimation.gt_tree_estimator",
    "evaluation.compute_log_likelihoods", "evaluation.create_maximal_matching_contact_map", "evaluation.l_infty_norm",
    "evaluation.rmse", "evaluation.mre", "evaluation.relative_errors", "evaluation.mean_relative_error",
    "markov_chain.matrix_exponential", "markov_chain.matrix_exponential_reversible", "markov_chain.chain_product",
    "markov_chain.compute_stationary_distribution", "markov_chain.compute_mutation_rate", "markov_chain.normalized",
    "markov_chain.FactorizedReversibleModel", "markov_chain.get_lg_path", "markov_chain.get_lg_stationary_path",
    "markov_chain.get_lg_x_lg_path", "markov_chain.get_lg_x_lg_stationary_path", "markov_chain.get_equ_path",
    "markov_chain.get_equ_x_equ_path", "markov_chain.get_wag_path", "markov_chain.get_wag_stationary_path",
    "markov_chain.equ_matrix", "markov_chain.wag_matrix", "markov_chain.wag_stationary_distribution",
    "caching.cached_computation", "caching.cached_parallel_computation", "caching.secure_parallel_output",
    "caching.set_cache_dir", "caching.set_dir_levels", "caching.set_hash_len", "caching.set_log_level",
    "caching.set_read_only", "caching.set_use_hash", "utils.amino_acids", "utils.get_amino_acids",
    "utils.quantization_idx", "utils.get_process_args", "utils.pushd", "utils.get_families",
    "types.PhylogenyEstimatorType", "io.Tree", "io.read_tree", "io.write_tree", "io.read_msa", "io.write_msa",
    "io.get_msa_num_sites", "io.get_msa_num_sequences", "io.get_msa_num_residues", "io.read_site_rates",
    "io.write_site_rates", "io.read_contact_map", "io.write_contact_map", "io.read_count_matrices",
    "io.write_count_matrices", "io.read_rate_matrix", "io.write_rate_matrix", "io.read_mask_matrix",
    "io.read_probability_distribution", "io.write_probability_distribution", "io.read_sites_subset",
    "io.write_sites_subset", "io.read_log_likelihood", "io.write_log_likelihood",
    "io.read_computed_cherries_from_file", "io.read_transitions", "io.write_transitions",
    "io.read_transitions_log_likelihood", "io.write_transitions_log_likelihood",
    "io.read_transitions_log_likelihood_per_site", "io.write_transitions_log_likelihood_per_site",
    "io.read_pickle", "io.write_pickle", "io.read_str", "io.write_str", "io.TransitionsType",
]


@pytest.mark.parametrize("path", PATHS)
def test_path_resolves(path):
    parts = path.split(".")
    obj = importlib.import_module("cherryml_b200")
    if len(parts) > 1:
        obj = importlib.import_module("cherryml_b200." + parts[0])
        parts = parts[1:]
    for a in parts:
        obj = getattr(obj, a)
    assert obj is not None
