"""Regenerate tests/golden/caching/dirs.json: the cache directories the UNMODIFIED reference's
caching decorators choose for calls of its own stage functions (imported from /root/reference
with the stubs of make_golden.import_reference; build container only).

    python tests/golden/make_golden_caching.py

For every case: the stage, the keyword arguments of the call, and the directory relative to the
cache root, with hashing on (the default, 128 hex digits in 3 directory levels) and off.  The
directory depends on the function's name, its parameter names, order and defaults, and on which
arguments the decorator leaves out -- everything a drop-in must reproduce for a cache written
by one implementation to be found by the other.
"""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests", "golden"))
OUT = os.path.join(REPO, "tests/golden/caching/dirs.json")

AA = list("ACDEFGHIKLMNPQRSTVWY")
GRID = [0.03 * 1.1 ** i for i in range(-3, 4)]


def cases():
    """(stage key, kwargs).  Paths are plain strings: nothing is read."""
    count = dict(tree_dir="/d/trees", msa_dir="/d/msas", site_rates_dir="/d/rates", families=["f1", "f0"],
                 amino_acids=AA, quantization_points=GRID, edge_or_cherry="cherry++", num_processes=4)
    yield "count_transitions", count
    yield "count_transitions", {**count, "edge_or_cherry": "edge", "use_cpp_implementation": False, "cpp_command_line_prefix": "x",
                                "cpp_command_line_suffix": "y"}
    co = dict(tree_dir="/d/trees", msa_dir="/d/msas", contact_map_dir="/d/cm", families=["f1", "f0"],
              amino_acids=AA, quantization_points=[str(q) for q in GRID], edge_or_cherry="cherry", minimum_distance_for_nontrivial_contact=7,
              num_processes=2)
    yield "count_co_transitions", co
    mle = dict(count_matrices_path="/d/counts/result.txt", initialization_path=None, mask_path=None,
               stationary_distribution_path=None, rate_matrix_parameterization="pande_reversible", device="cpu",
               learning_rate=1e-1, num_epochs=500, do_adam=True)
    yield "quantized_transitions_mle", mle
    yield "quantized_transitions_mle", {**mle, "initialization_path": "/d/jtt/result.txt", "device": "cuda",
                                        "OMP_NUM_THREADS": 8, "OPENBLAS_NUM_THREADS": 8, "num_epochs": 2000,
                                        "loss_normalization": True}
    yield "jtt_ipw", dict(count_matrices_path="/d/counts/result.txt", mask_path=None, use_ipw=True, normalize=False)
    fc = dict(msa_dir="/d/msas", families=["b", "a"], rate_matrix_path="/d/lg.txt", num_rate_categories=20,
              max_iters=50, num_processes=32)
    yield "fast_cherries", fc
    yield "fast_cherries", {**fc, "num_rate_categories": 4, "verbose": False, "seed": 7,
                            "quantization_grid_num_steps": 32, "remake": True}
    ll = dict(tree_dir="/d/trees", msa_dir="/d/msas", site_rates_dir="/d/rates", contact_map_dir=None,
              families=["f0"], amino_acids=AA, pi_1_path="/d/pi.txt", Q_1_path="/d/q.txt", reversible_1=True,
              device_1="cpu", pi_2_path=None, Q_2_path=None, reversible_2=None, device_2=None, num_processes=1)
    yield "compute_log_likelihoods", ll
    yield "compute_log_likelihoods", {**ll, "contact_map_dir": "/d/cm", "pi_2_path": "/d/pi2.txt",
                                      "Q_2_path": "/d/q2.txt", "reversible_2": True, "device_2": "cuda",
                                      "device_1": "cuda", "num_processes": 8}


    yield "create_maximal_matching_contact_map", dict(i_contact_map_dir="/d/cm", families=["f1", "f0"],
                                                      minimum_distance_for_nontrivial_contact=7, num_processes=4)
    yield "gt_tree_estimator", dict(gt_tree_dir="/d/t", gt_site_rates_dir="/d/r", gt_likelihood_dir="/d/l",
                                    msa_dir="/d/msas", families=["f0"], rate_matrix_path="/d/lg.txt",
                                    num_rate_categories=4, num_processes=2)


def reference_stage(cherryml, key):
    import cherryml.counting
    import cherryml.estimation
    import cherryml.evaluation
    import cherryml.phylogeny_estimation
    import cherryml.phylogeny_estimation._fast_cherries as fc

    return {"count_transitions": cherryml.counting.count_transitions,
            "count_co_transitions": cherryml.counting.count_co_transitions,
            "quantized_transitions_mle": cherryml.estimation.quantized_transitions_mle,
            "jtt_ipw": cherryml.estimation.jtt_ipw,
            "fast_cherries": fc.fast_cherries,
            "compute_log_likelihoods": cherryml.evaluation.compute_log_likelihoods,
            "create_maximal_matching_contact_map": cherryml.evaluation.create_maximal_matching_contact_map,
            "gt_tree_estimator": cherryml.phylogeny_estimation.gt_tree_estimator}[key]


def reference_dir(stage, kwargs, use_hash):
    """Through the reference's own helpers, with the decorator's lists read from the wrapper's closure."""
    from cherryml.caching._cached_computation import _get_func_caching_dir_aux
    from cherryml.caching._cached_parallel_computation import _get_parallel_func_caching_dir_aux

    cells = dict(zip(stage.__code__.co_freevars, (c.cell_contents for c in stage.__closure__)))
    func = cells["func"]
    common = dict(func=func, exclude_args=list(cells["exclude_args"]),
                  exclude_args_if_default=list(cells["exclude_args_if_default"]),
                  output_dirs=list(cells["output_dirs"]), args=[], kwargs=dict(kwargs), cache_dir="CACHE",
                  use_hash=use_hash)
    if "parallel_arg" in cells:
        return _get_parallel_func_caching_dir_aux(parallel_arg=cells["parallel_arg"], **common)
    return _get_func_caching_dir_aux(**common)


def main():
    from make_golden import import_reference

    cherryml = import_reference()
    out = []
    for key, kwargs in cases():
        stage = reference_stage(cherryml, key)
        out.append({"stage": key, "kwargs": kwargs,
                    "hashed": os.path.relpath(reference_dir(stage, kwargs, True), "CACHE"),
                    "plain": os.path.relpath(reference_dir(stage, kwargs, False), "CACHE")})
        print(key, out[-1]["hashed"][:24], "...", out[-1]["plain"][:100])
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", OUT, len(out), "cases")


if __name__ == "__main__":
    main()
