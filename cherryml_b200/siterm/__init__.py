from ._vectorized import quantized_transitions_mle_vectorized_over_sites, solve_stationary_dist_fast  # noqa: F401
