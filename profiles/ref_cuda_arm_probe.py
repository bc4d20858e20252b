"""Probe of the reference's device="cuda" fit arm (bench infrastructure): small LG-shaped counts through
oracle/run_reference_fit.py with a traceback dump if it stalls."""
import os
import subprocess
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from cherryml_b200.estimation._jtt_ipw import jtt_ipw_from_counts
from cherryml_b200.io import write_count_matrices_array, write_rate_matrix
from cherryml_b200.synthetic import quantization_grid
from cherryml_b200.utils import amino_acids

grid = quantization_grid()
rng = np.random.default_rng(0)
c = rng.integers(0, 50, (len(grid), 20, 20)).astype(np.float64)
c = c + c.transpose(0, 2, 1)
d = tempfile.mkdtemp()
write_count_matrices_array(list(grid), list(amino_acids), c, d + "/c.txt", "python")
write_rate_matrix(jtt_ipw_from_counts(grid, torch.from_numpy(c)), list(amino_acids), d + "/i.txt")
env = dict(os.environ, CHERRY_REF_FIT_DUMP_AFTER="40")
r = subprocess.run([sys.executable, "oracle/run_reference_fit.py", "--counts", d + "/c.txt", "--init", d + "/i.txt",
                    "--device", sys.argv[1] if len(sys.argv) > 1 else "cuda", "--epochs", "5", "--threads", "8",
                    "--out", d + "/o"], capture_output=True, text=True, timeout=100, env=env)
print("rc", r.returncode)
print(r.stdout[-2000:])
print(r.stderr[-6000:])
