"""A cache written by the reference must be found by this package and vice versa: for calls of
the stage functions, the cache directory (function name, parameter names / order / defaults,
arguments the decorator leaves out, hashing) equals the one the UNMODIFIED reference's
decorators choose (tests/golden/caching/dirs.json, made by tests/golden/make_golden_caching.py)."""
import json
import os

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = json.load(open(os.path.join(HERE, "golden", "caching", "dirs.json")))


def _stage(key):
    import cherryml_b200 as pkg
    import cherryml_b200.evaluation  # noqa: F401

    return {"count_transitions": pkg.count_transitions, "count_co_transitions": pkg.count_co_transitions,
            "quantized_transitions_mle": pkg.quantized_transitions_mle, "jtt_ipw": pkg.jtt_ipw,
            "fast_cherries": pkg.fast_cherries, "compute_log_likelihoods": pkg.compute_log_likelihoods,
            "create_maximal_matching_contact_map": pkg.evaluation.create_maximal_matching_contact_map,
            "gt_tree_estimator": pkg.gt_tree_estimator}[key]


@pytest.mark.parametrize("i", range(len(CASES)))
def test_cache_directory_equals_the_reference(i):
    case = CASES[i]
    stage = _stage(case["stage"])
    assert os.path.relpath(stage.caching_dir("CACHE", use_hash=True, **case["kwargs"]), "CACHE") == case["hashed"]
    assert os.path.relpath(stage.caching_dir("CACHE", use_hash=False, **case["kwargs"]), "CACHE") == case["plain"]


def test_arguments_this_package_adds_do_not_change_the_directory():
    """device / process_group / streaming knobs exist only here; they must stay out of the key."""
    import cherryml_b200 as pkg

    case = next(c for c in CASES if c["stage"] == "fast_cherries")
    d = pkg.fast_cherries.caching_dir("CACHE", use_hash=True, device="cuda:3", process_group=object(), **case["kwargs"])
    assert os.path.relpath(d, "CACHE") == case["hashed"]
    case = next(c for c in CASES if c["stage"] == "compute_log_likelihoods")
    d = pkg.compute_log_likelihoods.caching_dir("CACHE", use_hash=True, process_group=object(), **case["kwargs"])
    assert os.path.relpath(d, "CACHE") == case["hashed"]
    for key in ("count_transitions", "count_co_transitions"):
        case = next(c for c in CASES if c["stage"] == key)
        d = getattr(pkg, key).caching_dir("CACHE", use_hash=True, device="cuda:1", process_group=object(),
                                          ingest="python", families_per_batch=7, **case["kwargs"])
        assert os.path.relpath(d, "CACHE") == case["hashed"]
