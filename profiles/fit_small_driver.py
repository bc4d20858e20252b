"""LG-shaped 20x20 fit (K = 100) for profiling the small-fit kernel: python profiles/fit_small_driver.py [epochs]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from cherryml_b200.counting._device import count_raw, sorted_grid, symmetrize
from cherryml_b200.estimation import FitEngine, jtt_ipw_from_counts, theta_from_initialization
from cherryml_b200.synthetic import as_device_batch, quantization_grid, synthetic_lg

epochs = int(sys.argv[1]) if len(sys.argv) > 1 else 200
device = torch.device("cuda", 0)
grid = quantization_grid()
K = len(grid)
dev = as_device_batch(synthetic_lg(512, 1024, 300, 4, seed=7, device=device), device)
gd = torch.from_numpy(sorted_grid(grid)).to(device)
c = symmetrize(count_raw(dev, gd, K, 20), "lg", K, 20, False)
init = jtt_ipw_from_counts(grid, c)
theta0 = theta_from_initialization(init, np.ones((20, 20)))
eng = FitEngine(np.asarray(grid), c, theta0, num_epochs=epochs + 64, device=device)
eng.run(64)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
eng.run(epochs)
e1.record()
torch.cuda.synchronize()
print("ms per epoch %.4f" % (e0.elapsed_time(e1) / epochs))
