"""Fit timings reported by bench.py under the "fit" key (BASELINE.json: "end-to-end fit
seconds (LG 20x20, coevo 400x400)")."""
import ctypes
import time
from typing import Dict, Optional

import numpy as np
import torch

from .. import _lib
from ._engine import FitEngine, theta_from_initialization
from ._jtt_ipw import jtt_ipw_from_counts


def measure_fp64_gemm_peak(device, n: int = 4096, reps: int = 5) -> float:
    """cuBLAS DGEMM throughput (TFLOP/s) on this GPU: the denominator for the 400x400 fit's
    roofline, measured here because MEASURED_PEAKS.json has no FP64 figure."""
    a = torch.randn(n, n, dtype=torch.float64, device=device)
    b = torch.randn(n, n, dtype=torch.float64, device=device)
    for _ in range(2):
        torch.matmul(a, b)
    torch.cuda.synchronize(device)
    best = float("inf")
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize(device)
        best = min(best, e0.elapsed_time(e1))
    return 2.0 * n**3 / (best * 1e-3) / 1e12


def timed_fit(times, counts: torch.Tensor, num_epochs: int, mask=None) -> Dict:
    """JTT-IPW initialisation + `num_epochs` Adam epochs + results back on the host."""
    device = counts.device
    torch.cuda.synchronize(device)
    t0 = time.perf_counter()
    init = jtt_ipw_from_counts(times, counts, mask=mask)
    S = counts.shape[-1]
    theta0 = theta_from_initialization(init, np.ones((S, S)) if mask is None else mask)
    eng = FitEngine(np.asarray(times), counts, theta0, mask=mask, num_epochs=num_epochs, device=device)
    torch.cuda.synchronize(device)
    t_setup = time.perf_counter()
    _lib.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.run()
    e1.record()
    res = eng.results()
    t1 = time.perf_counter()
    out = {
        "S": int(S), "K": int(len(times)), "num_epochs": int(num_epochs),
        "seconds_end_to_end": t1 - t0, "seconds_setup_init": t_setup - t0,
        "seconds_device_epochs": e0.elapsed_time(e1) * 1e-3,
        "ms_per_epoch": e0.elapsed_time(e1) / max(1, num_epochs),
        "loss_first": float(res["loss"][0]), "loss_last": float(res["loss"][-1]),
        "gpu_launches": _lib.launch_count(),
    }
    if S > 32:
        s = (ctypes.c_int * len(times))()
        mu, deg = ctypes.c_double(0), ctypes.c_int(0)
        _lib.check(_lib.load().cherry_fit_schedule(ctypes.byref(eng.args), s, ctypes.byref(mu), ctypes.byref(deg)),
                   "cherry_fit_schedule")
        sq = int(sum(s))
        products_fwd = (deg.value - 1) + sq
        flops_epoch = 3.0 * products_fwd * 2.0 * S**3
        out.update({
            "taylor_degree": deg.value, "squarings_total": sq, "squarings_max": int(max(s)),
            "matrix_products_per_epoch": 3 * products_fwd, "flop_per_epoch": flops_epoch,
            "tflops_executed": flops_epoch * num_epochs / out["seconds_device_epochs"] / 1e12,
            "mu": mu.value,
        })
    return out


def bench_fit(device, lg_times=None, lg_counts: Optional[torch.Tensor] = None, num_epochs: int = 500,
              co_families: int = 256) -> Dict:
    from ..counting._device import count_raw, sorted_grid, symmetrize
    from ..synthetic import as_device_batch, quantization_grid, synthetic_co, synthetic_lg

    grid = quantization_grid()
    K = len(grid)
    out: Dict = {"metric": "end-to-end fit seconds", "init": "jtt-ipw", "optimizer": "Adam lr=0.1",
                 "dtype": "f64"}
    if lg_counts is None:
        dev = as_device_batch(synthetic_lg(512, 1024, 300, 4, seed=7, device=device), device)
        gd = torch.from_numpy(sorted_grid(grid)).to(device)
        lg_counts = symmetrize(count_raw(dev, gd, K, 20), "lg", K, 20, False)
        lg_times = grid
    timed_fit(lg_times, lg_counts, 64)  # warm-up (module load, graph instantiation paths)
    out["lg_20x20"] = timed_fit(lg_times, lg_counts, num_epochs)
    # co-evolution counts from synthetic contact-map families (BASELINE config 4 shape)
    dev = as_device_batch(synthetic_co(co_families, 1024, 300, seed=11, device=device), device)
    gd = torch.from_numpy(sorted_grid(grid)).to(device)
    torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    raw = count_raw(dev, gd, K, 20)
    co_counts = symmetrize(raw, "co", K, 20, False)
    e1.record()
    torch.cuda.synchronize(device)
    out["co_counting"] = {"families": co_families, "items_examined": dev.n_sites_examined,
                          "ms": e0.elapsed_time(e1),
                          "items_per_s": dev.n_sites_examined / (e0.elapsed_time(e1) * 1e-3)}
    del raw, dev
    peak = measure_fp64_gemm_peak(device)
    timed_fit(grid, co_counts, 4)
    co = timed_fit(grid, co_counts, num_epochs)
    co["roofline"] = {"bound": "tensor", "unit": "TFLOP/s", "achieved": co["tflops_executed"], "peak": peak,
                      "frac": co["tflops_executed"] / peak,
                      "peak_source": "cuBLAS DGEMM 4096^3 measured in this run (no FP64 figure in MEASURED_PEAKS.json)"}
    out["coevo_400x400"] = co
    return out
