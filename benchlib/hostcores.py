"""How many host cores this process can actually run on.

``os.cpu_count()`` is the machine's core count; inside a container the scheduler affinity mask and the
cgroup CPU quota can both be smaller.  Starting one OpenMP thread per MACHINE core under a smaller
quota makes the spin-waiting threads throttle each other (measured: the reference's 20x20 fit arm went
from 1.8 s to 75 s on a 2-GPU box slice), so every CPU arm of the bench sizes itself with this."""
import math
import os


def _cgroup_quota():
    # cgroup v2: "max 100000" or "<quota> <period>"; v1: cpu.cfs_quota_us / cpu.cfs_period_us
    try:
        with open("/sys/fs/cgroup/cpu.max") as f:
            q, p = f.read().split()[:2]
        if q != "max" and float(p) > 0:
            return float(q) / float(p)
    except (OSError, ValueError):
        pass
    try:
        with open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us") as f:
            q = float(f.read())
        with open("/sys/fs/cgroup/cpu/cpu.cfs_period_us") as f:
            p = float(f.read())
        if q > 0 and p > 0:
            return q / p
    except (OSError, ValueError):
        pass
    return None


def usable_cores() -> int:
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except (AttributeError, OSError):
        pass
    quota = _cgroup_quota()
    if quota is not None:
        n = min(n, max(1, math.floor(quota + 1e-9)))
    return max(1, n)


def describe() -> dict:
    try:
        aff = len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        aff = None
    try:
        load = os.getloadavg()[0]
    except OSError:
        load = None
    return {"cpu_count": os.cpu_count(), "affinity": aff, "cgroup_quota": _cgroup_quota(),
            "usable": usable_cores(), "loadavg_1m": load}
