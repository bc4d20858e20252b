"""Type of a tree estimator stage, as the end-to-end drivers take it (reference cherryml/types.py).

An estimator receives the MSA directory, the family names and a rate matrix path (plus its own
keyword arguments bound beforehand) and returns the output directories by name:
``output_tree_dir``, ``output_site_rates_dir``, ``output_likelihood_dir``.
"""
from typing import Callable, Dict, List

PhylogenyEstimatorType = Callable[[str, List[str], str], Dict[str, str]]
