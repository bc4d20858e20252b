"""The "ground truth" tree estimator: a stage with the tree estimators' interface that hands
back trees, site rates and likelihoods that already exist (simulations, or trees estimated
elsewhere), so that they can flow through ``lg_end_to_end_with_cherryml_optimizer`` like any
estimator's output.  Reference ``cherryml/phylogeny_estimation/_gt_tree_estimator.py:35-120``."""
import os
from typing import List, Optional

from ..caching import cached_parallel_computation
from ..io import read_log_likelihood, read_site_rates, read_tree, write_log_likelihood, write_site_rates, write_tree


@cached_parallel_computation(
    parallel_arg="families",
    exclude_args=["num_processes"],
    output_dirs=["output_tree_dir", "output_site_rates_dir", "output_likelihood_dir"],
    write_extra_log_files=True,
)
def gt_tree_estimator(
    gt_tree_dir: str,
    gt_site_rates_dir: str,
    gt_likelihood_dir: str,
    msa_dir: str,
    families: List[str],
    rate_matrix_path: str,
    num_rate_categories: int,
    num_processes: int,
    output_tree_dir: Optional[str] = None,
    output_site_rates_dir: Optional[str] = None,
    output_likelihood_dir: Optional[str] = None,
) -> None:
    """Per family: the given tree, site rates and log-likelihood re-written into the output
    directories (parsed and formatted again, as the reference does) and ``<family>.profiling``.
    ``msa_dir``, ``rate_matrix_path``, ``num_rate_categories`` only enter the cache key;
    ``num_processes`` is accepted and ignored (the work is a few small files per family)."""
    for family in families:
        name = family + ".txt"
        write_tree(read_tree(os.path.join(gt_tree_dir, name)), os.path.join(output_tree_dir, name))
        write_site_rates(read_site_rates(os.path.join(gt_site_rates_dir, name)),
                         os.path.join(output_site_rates_dir, name))
        write_log_likelihood(read_log_likelihood(os.path.join(gt_likelihood_dir, name)),
                             os.path.join(output_likelihood_dir, name))
        with open(os.path.join(output_tree_dir, family + ".profiling"), "w") as f:
            f.write(f"time_gt_tree_estimator: {0}")
