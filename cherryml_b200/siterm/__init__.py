from ._counting import get_raw_count_matrices, get_raw_count_matrices_device  # noqa: F401
from ._vectorized import quantized_transitions_mle_vectorized_over_sites, solve_stationary_dist_fast  # noqa: F401
