"""cherryml_b200: B200-native (sm_100a) implementation of CherryML's hot path.

Transition counting into quantized-time count tensors and the batched matrix-exponential
composite-likelihood fit of the rate matrix, behind CherryML's own Python API.  All compute
runs in hand-written CUDA kernels reached through a C ABI (``include/cherryml_b200.h``);
there is no CPU fallback.
"""
__version__ = "0.1.0"

from . import caching  # noqa: F401
from .counting import count_co_transitions, count_transitions  # noqa: F401
from .estimation import jtt_ipw, quantized_transitions_mle  # noqa: F401
from ._public_api import (  # noqa: F401
    cherryml_public_api,
    coevolution_end_to_end_with_cherryml_optimizer,
    lg_end_to_end_with_cherryml_optimizer,
)
from .evaluation import compute_log_likelihoods  # noqa: F401
from .phylogeny_estimation import fast_cherries, gt_tree_estimator  # noqa: F401
from .siterm import learn_site_specific_rate_matrices  # noqa: F401
from .types import PhylogenyEstimatorType  # noqa: F401
