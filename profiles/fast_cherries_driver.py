"""Times the two FastCherries kernels on a synthetic Pfam-shaped batch (families x 1024 x 300,
BASELINE config 3's shape); `python profiles/fast_cherries_driver.py [families] [R]`."""
import math
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from cherryml_b200 import _lib
from cherryml_b200.phylogeny_estimation import _fast_cherries as fc
from cherryml_b200.synthetic import synthetic_fc


def main():
    n_fams = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    R = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    msa, fams = synthetic_fc(n_fams, 1024, 300, seed=0)
    from cherryml_b200.io import read_rate_matrix
    from cherryml_b200.markov_chain import get_lg_path

    Q = read_rate_matrix(get_lg_path()).to_numpy()
    grid = fc.quantization_grid(0.03, 1.1, 64)
    cats = fc.ble_rate_categories(R)
    weights = fc.initial_rate_weights(cats)
    priors = np.array([2 * math.log(r) - 3 * r for r in cats])
    table = fc.log_transition_table(Q, grid, cats, "cuda:0")
    for rep in range(3):
        t = time.time()
        out = fc.fast_cherries_device(msa, fams, 20, table, priors, weights, 1234, 50, "cuda:0")
        print(f"rep {rep}: families {n_fams} R {R} pair {out['pair_ms']:.2f} ms ble {out['ble_ms']:.2f} ms "
              f"wall {time.time() - t:.3f} s iters mean {out['iters'].mean():.1f} max {out['iters'].max()}",
              flush=True)


if __name__ == "__main__":
    main()
