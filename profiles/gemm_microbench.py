"""DMMA GEMM micro-benchmark: our batched 400^3 kernel vs cuBLAS (torch.matmul), CUDA events.

    python profiles/gemm_microbench.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from cherryml_b200.estimation._gemm import gemm_f64_batched


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


for n, batch in [(400, 1), (400, 8), (400, 30), (400, 100), (400, 148), (400, 296), (800, 37), (4000, 1)]:
    A = torch.randn(batch, n, n, dtype=torch.float64, device="cuda")
    B = torch.randn(batch, n, n, dtype=torch.float64, device="cuda")
    flop = 2.0 * batch * n**3
    t_ours = timeit(lambda: gemm_f64_batched(A, B))
    t_cublas = timeit(lambda: torch.matmul(A, B))
    line = f"n={n:5d} batch={batch:4d}  ours {t_ours*1e3:9.1f} us {flop/t_ours/1e9:7.2f} TF/s | cuBLAS {t_cublas*1e3:9.1f} us {flop/t_cublas/1e9:7.2f} TF/s"
    if batch == 1 and n == 400:
        t5 = timeit(lambda: gemm_f64_batched(A, B, ksplit=5))
        line += f" | ours ksplit=5 {t5*1e3:7.1f} us"
    print(line)
