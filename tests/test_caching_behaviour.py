"""Behaviour of the caching decorators (reference caching/_cached_computation.py:150-369,
_cached_parallel_computation.py:162-440): what is computed when, where outputs go, tokens, modes."""
import os
import stat

import pytest

from cherryml_b200 import caching
from cherryml_b200.caching import CacheUsageError, cached_computation, cached_parallel_computation


@pytest.fixture()
def cache(tmp_path):
    caching.set_cache_dir(str(tmp_path / "cache"))
    try:
        yield str(tmp_path / "cache")
    finally:
        caching.set_cache_dir(None)
        caching.set_read_only(False)


def _mode(path):
    return stat.S_IMODE(os.stat(path).st_mode)


def test_cached_computation(cache, tmp_path):
    calls = []

    @cached_computation(exclude_args=["verbose"], exclude_args_if_default=["extra"], output_dirs=["out_dir"],
                        write_extra_log_files=True)
    def square(x, out_dir=None, verbose=False, extra=0):
        calls.append((x, extra))
        with open(os.path.join(out_dir, "result.txt"), "w") as f:
            f.write(str(x * x + extra))

    with pytest.raises(CacheUsageError, match="keyword arguments only"):
        square(3)
    r1 = square(x=3)
    out = r1["out_dir"]
    assert out.startswith(os.path.join(cache, "square")) and out.endswith("out_dir")
    assert open(os.path.join(out, "result.txt")).read() == "9"
    assert open(os.path.join(out, "result.success")).read() == "SUCCESS\n"
    assert _mode(os.path.join(out, "result.txt")) == 0o444
    assert os.path.exists(os.path.join(out, "_function_binding.log"))
    assert "x_3" in open(os.path.join(out, "_unhashed_output_dir.log")).read()
    assert square(x=3, verbose=True) == r1 and square(x=3, extra=0) == r1  # excluded / default: same entry
    assert calls == [(3, 0)]
    r2 = square(x=3, extra=1)  # a non-default value enters the key
    assert r2 != r1 and calls == [(3, 0), (3, 1)]
    given = str(tmp_path / "mine")
    assert square(x=4, out_dir=given) == {"out_dir": given}
    assert open(os.path.join(given, "result.txt")).read() == "16"
    caching.set_read_only(True)
    assert square(x=3) == r1
    with pytest.raises(CacheUsageError, match="read only"):
        square(x=5)
    caching.set_read_only(False)

    @cached_computation(output_dirs=["out_dir"])
    def forgets(x, out_dir=None):
        pass

    with pytest.raises(CacheUsageError, match="should have created"):
        forgets(x=1)
    with pytest.raises(CacheUsageError, match="is not an argument"):
        cached_computation(exclude_args=["nope"], output_dirs=["out_dir"])(forgets.__wrapped__)
    with pytest.raises(CacheUsageError, match="should be distinct"):
        cached_computation(exclude_args=["x", "x"], output_dirs=["out_dir"])(forgets.__wrapped__)


def test_cached_parallel_computation(cache, tmp_path):
    calls = []

    @cached_parallel_computation(parallel_arg="items", exclude_args=["workers"], output_dirs=["a_dir", "b_dir"])
    def work(items, scale, a_dir=None, b_dir=None, workers=1):
        calls.append(list(items))
        for v in items:
            for d, k in ((a_dir, 1), (b_dir, 2)):
                with open(os.path.join(d, v + ".txt"), "w") as f:
                    f.write(f"{v}:{scale * k}")

    r = work(items=["b", "a", "b"], scale=2)
    assert calls == [["a", "b"]]  # sorted, de-duplicated
    assert set(r) == {"a_dir", "b_dir"} and os.path.dirname(r["a_dir"]) == os.path.dirname(r["b_dir"])
    assert open(os.path.join(r["b_dir"], "a.txt")).read() == "a:4"
    assert _mode(os.path.join(r["a_dir"], "b.txt")) == 0o444
    assert os.path.exists(os.path.join(r["a_dir"], "b.success"))
    assert work(items=["c", "a"], scale=2, workers=9) == r  # same entry; only the new item is computed
    assert calls == [["a", "b"], ["c"]]
    assert work(items=["a", "b", "c"], scale=2) == r and len(calls) == 2
    os.chmod(os.path.join(r["b_dir"], "c.success"), 0o666)
    os.remove(os.path.join(r["b_dir"], "c.success"))  # a missing token in ANY output dir -> recomputed everywhere
    work(items=["c"], scale=2)
    assert calls[-1] == ["c"] and len(calls) == 3
    assert work(items=["a"], scale=3) != r and calls[-1] == ["a"]
    caching.set_read_only(True)
    with pytest.raises(CacheUsageError, match="read only"):
        work(items=["z"], scale=2)
    assert work(items=["a"], scale=2) == r


def test_without_a_cache_dir_the_function_just_runs(tmp_path):
    caching.set_cache_dir(None)
    seen = []

    @cached_computation(output_dirs=["out_dir"])
    def f(x, out_dir=None):
        seen.append(out_dir)

    f(x=1, out_dir=str(tmp_path))
    assert seen == [str(tmp_path)]


def test_gt_tree_estimator_stage(cache, tmp_path):
    """Reference phylogeny_estimation/_gt_tree_estimator.py: the given files come back through the
    tree estimators' interface, cached per family."""
    from cherryml_b200 import io
    from cherryml_b200.phylogeny_estimation import gt_tree_estimator

    for f in ("f0", "f1"):
        t = io.Tree()
        t.add_nodes(["r", "a", "b"])
        t.add_edges([("r", "a", 0.1), ("r", "b", 0.25)])
        io.write_tree(t, str(tmp_path / "gt_tree" / (f + ".txt")))
        io.write_site_rates([1.0, 0.5], str(tmp_path / "gt_rates" / (f + ".txt")))
        io.write_log_likelihood((-3.0, [-1.0, -2.0]), str(tmp_path / "gt_ll" / (f + ".txt")))
    kw = dict(gt_tree_dir=str(tmp_path / "gt_tree"), gt_site_rates_dir=str(tmp_path / "gt_rates"),
              gt_likelihood_dir=str(tmp_path / "gt_ll"), msa_dir="unused", rate_matrix_path="unused",
              num_rate_categories=2, num_processes=1)
    out = gt_tree_estimator(families=["f1", "f0"], **kw)
    assert set(out) == {"output_tree_dir", "output_site_rates_dir", "output_likelihood_dir"}
    for f in ("f0", "f1"):
        for d, src in (("output_tree_dir", "gt_tree"), ("output_site_rates_dir", "gt_rates"),
                       ("output_likelihood_dir", "gt_ll")):
            assert open(os.path.join(out[d], f + ".txt")).read() == open(tmp_path / src / (f + ".txt")).read()
            assert os.path.exists(os.path.join(out[d], f + ".success"))
        assert open(os.path.join(out["output_tree_dir"], f + ".profiling")).read() == "time_gt_tree_estimator: 0"
    os.remove(tmp_path / "gt_tree" / "f0.txt")  # cached: the inputs are not read again
    assert gt_tree_estimator(families=["f0"], **kw) == out


def test_secure_parallel_output(tmp_path):
    p = tmp_path / "x.txt"
    p.write_text("1")
    caching.secure_parallel_output(str(tmp_path), "x")
    assert _mode(str(p)) == 0o444 and (tmp_path / "x.success").read_text() == "SUCCESS\n"
