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
