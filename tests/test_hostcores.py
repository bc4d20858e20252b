"""bench helper: the core count the CPU arms size themselves with (benchlib/hostcores.py)."""
import os

from benchlib import hostcores


def test_usable_cores_is_bounded_by_the_machine_and_the_affinity_mask():
    n = hostcores.usable_cores()
    assert 1 <= n <= (os.cpu_count() or 1)
    if hasattr(os, "sched_getaffinity"):
        assert n <= len(os.sched_getaffinity(0))
    d = hostcores.describe()
    assert d["usable"] == n and d["cpu_count"] == os.cpu_count()


def test_cgroup_quota_caps_the_count(monkeypatch):
    monkeypatch.setattr(hostcores, "_cgroup_quota", lambda: 2.5)
    assert hostcores.usable_cores() == min(2, os.cpu_count() or 1)
    monkeypatch.setattr(hostcores, "_cgroup_quota", lambda: 0.3)
    assert hostcores.usable_cores() == 1
    monkeypatch.setattr(hostcores, "_cgroup_quota", lambda: None)
    assert hostcores.usable_cores() >= 1
