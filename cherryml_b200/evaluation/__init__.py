"""Tree log-likelihoods on the GPU: drop-in for ``cherryml/evaluation/_likelihood.py``."""
from ._likelihood import compute_log_likelihoods, dp_likelihood_computation  # noqa: F401
