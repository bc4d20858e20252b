"""Round 2: why the reference CPU fit arm was 40x slower on a 2-GPU box slice.  Prints what the box
gives this process (cpu_count / affinity / cgroup quota / load) and a torch CPU matmul loop timed at
several thread counts."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from benchlib.hostcores import describe

print(json.dumps(describe()))
for p in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu.stat", "/sys/fs/cgroup/cpuset.cpus.effective"):
    try:
        print(p, open(p).read().strip().replace("\n", " | ")[:300])
    except OSError as e:
        print(p, "unreadable", e)
import torch

x = torch.randn(64, 20, 20, dtype=torch.float64)
for t in (os.cpu_count(), describe()["usable"], 8, 4, 1):
    torch.set_num_threads(int(t))
    for _ in range(20):
        torch.matrix_exp(x)
    t0 = time.perf_counter()
    for _ in range(200):
        torch.matrix_exp(x)
    print("threads", t, "matrix_exp 64x20x20 x200:", round(time.perf_counter() - t0, 4), "s")
