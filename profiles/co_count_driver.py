"""Driver used under ncu: co-transition counting on synthetic contact-map families."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from cherryml_b200.counting._device import build_bucket_table, count_raw, sorted_grid
from cherryml_b200.synthetic import as_device_batch, quantization_grid, synthetic_co

fams = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
device = torch.device("cuda", 0)
grid = quantization_grid()
K = len(grid)
gd = torch.from_numpy(sorted_grid(grid)).to(device)
dev = as_device_batch(synthetic_co(fams, 1024, 300, seed=11, device=device), device)
tab = build_bucket_table(dev, gd, K)
for _ in range(3):
    raw = count_raw(dev, gd, K, 20, tab=tab)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
raw = count_raw(dev, gd, K, 20, tab=tab)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"families {fams} items {dev.n_sites_examined} ms {ms:.3f} items/s {dev.n_sites_examined / ms * 1e3:.3e} "
      f"residue GB/s {dev.msa.numel() / ms / 1e6:.1f} counted {int(raw.sum().item())}")
