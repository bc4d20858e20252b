"""Phase timeline + chain-kernel phase profile + graph-replayed ms/epoch of the 400x400 fit.

    CHERRY_FIT_TIMELINE=1 python profiles/fit_timeline_driver.py [--families 4096] [--epochs 200]
"""
import argparse
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from cherryml_b200 import _lib
from cherryml_b200.counting._device import count_raw, sorted_grid, symmetrize
from cherryml_b200.estimation import FitEngine, jtt_ipw_from_counts, theta_from_initialization
from cherryml_b200.synthetic import as_device_batch, quantization_grid, synthetic_co

ap = argparse.ArgumentParser()
ap.add_argument("--families", type=int, default=4096)
ap.add_argument("--epochs", type=int, default=192)
args = ap.parse_args()
device = torch.device("cuda", 0)
grid = quantization_grid()
K = len(grid)
dev = as_device_batch(synthetic_co(args.families, 1024, 300, seed=11, device=device), device)
gd = torch.from_numpy(sorted_grid(grid)).to(device)
c = symmetrize(count_raw(dev, gd, K, 20), "co", K, 20, False)
del dev
init = jtt_ipw_from_counts(grid, c)
theta0 = theta_from_initialization(init, np.ones((400, 400)))
eng = FitEngine(np.asarray(grid), c, theta0, num_epochs=0, device=device)
eng.loss_and_grad()  # with CHERRY_FIT_TIMELINE set: prints the phase timeline to stderr
sarr = (ctypes.c_int * K)()
mu = ctypes.c_double(0)
deg = ctypes.c_int(0)
_lib.check(_lib.load().cherry_fit_schedule(ctypes.byref(eng.args), sarr, ctypes.byref(mu), ctypes.byref(deg)), "cherry_fit_schedule")
print("schedule: mu %.4f degree %d squarings total %d max %d" % (mu.value, deg.value, sum(sarr), max(sarr)))
os.environ.pop("CHERRY_FIT_TIMELINE", None)
eng2 = FitEngine(np.asarray(grid), c, theta0, num_epochs=args.epochs, device=device)
eng2.run(64)  # warm-up (graph instantiation)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
eng2.run(args.epochs - 64)
e1.record()
torch.cuda.synchronize()
print("graph-replayed: %.4f ms per epoch (symmetric form: %s)" % (e0.elapsed_time(e1) / (args.epochs - 64), eng2.symmetric_form))
