import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, json, numpy as np
from cherryml_b200.estimation import FitEngine, jtt_ipw_from_counts, theta_from_initialization
from cherryml_b200.synthetic import *
from cherryml_b200.counting._device import *
device=torch.device("cuda",0)
grid=quantization_grid(); K=len(grid)
dev = as_device_batch(synthetic_co(256, 1024, 300, seed=11, device=device), device)
gd = torch.from_numpy(sorted_grid(grid)).to(device)
c = symmetrize(count_raw(dev, gd, K, 20), "co", K, 20, False)
init = jtt_ipw_from_counts(grid, c)
eng = FitEngine(np.asarray(grid), c, theta_from_initialization(init, np.ones((400,400))), num_epochs=0, device=device)
eng.loss_and_grad()
