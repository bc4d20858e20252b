"""A/B of the LG counting kernels on the bench workload (BASELINE config 3 per GPU).

    [CHERRY_LG_BUCKET_PAD=p] python profiles/count_lg_ab_driver.py [--families 16384] [--mode table|fused]

mode table = cherry_build_bucket_table (one thread per pair) + cherry_count_lg, mode fused =
cherry_count_lg_fused (table built per tile); CHERRY_LG_BUCKET_PAD = padding cells per bucket of the
shared-memory histogram (default 5).  Prints ms per step (CUDA events, 10 steps after 3 warm-ups)
and a checksum of the raw histogram.
"""
import argparse
import os
import sys

ap = argparse.ArgumentParser()
ap.add_argument("--families", type=int, default=16384)
ap.add_argument("--mode", default="fused")
ap.add_argument("--cats", type=int, default=4)
ap.add_argument("--K", type=int, default=100)
args = ap.parse_args()
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from cherryml_b200.counting._device import build_bucket_table, count_raw, sorted_grid
from cherryml_b200.synthetic import as_device_batch, quantization_grid, synthetic_lg

device = torch.device("cuda", 0)
grid = quantization_grid(lo=-(args.K // 2), hi=args.K - args.K // 2 - 1)
K = len(grid)
syn = synthetic_lg(args.families, 1024, 300, args.cats, seed=1000, device=device)
dev = as_device_batch(syn, device)
gd = torch.from_numpy(sorted_grid(grid)).to(device)
raw = torch.zeros((K, 20, 20), dtype=torch.int64, device=device)


def step():
    raw.zero_()
    if args.mode == "fused":
        count_raw(dev, gd, K, 20, out=raw)
    else:
        tab = build_bucket_table(dev, gd, K)
        count_raw(dev, gd, K, 20, tab=tab, out=raw)


for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
w = torch.arange(raw.numel(), device=device, dtype=torch.int64) % 1000003
print(f"mode {args.mode:6s} K {K} cats {args.cats}: {ms:.4f} ms per step, {syn['n_sites_examined'] / ms / 1e9:.1f} G transitions/s, "
      f"sum {int(raw.sum())} checksum {int((raw.view(-1) * w).sum())}")
