from ._counting import get_raw_count_matrices, get_raw_count_matrices_device  # noqa: F401
from ._vectorized import quantized_transitions_mle_vectorized_over_sites, solve_stationary_dist_fast  # noqa: F401
from ._site_specific import (  # noqa: F401
    count_prior_probability_matrices,
    estimate_site_specific_rate_matrices_given_tree_and_site_rates,
    get_cherry_transitions,
    get_edge_transitions,
)
from ._public_api import (  # noqa: F401
    estimate_site_rates,
    get_standard_site_rate_grid,
    get_standard_site_rate_prior,
    learn_site_rate_matrices,
    learn_site_specific_rate_matrices,
)
