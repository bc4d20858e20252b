"""Where one warmed learn_site_specific_rate_matrices(tree=None) call spends its host time:
every time_* key of the result, then cProfile's top cumulative entries."""
import cProfile
import io
import pstats
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from benchlib.siterm import GRID_STEPS, NUM_EPOCHS, _plant_family  # noqa: E402
from cherryml_b200.io import read_rate_matrix  # noqa: E402
from cherryml_b200.markov_chain import get_lg_path  # noqa: E402
from cherryml_b200.siterm import learn_site_specific_rate_matrices  # noqa: E402
from cherryml_b200.utils import amino_acids  # noqa: E402

rng = np.random.default_rng(0)
lg = read_rate_matrix(get_lg_path())
msas = [_plant_family(rng) for _ in range(3)]
kw = dict(tree=None, alphabet=list(amino_acids), regularization_rate_matrix=lg, regularization_strength=0.5,
          device="cuda:0", num_epochs=NUM_EPOCHS, quantization_grid_num_steps=GRID_STEPS)
learn_site_specific_rate_matrices(msa=msas[0], **kw)
t0 = time.perf_counter()
r = learn_site_specific_rate_matrices(msa=msas[1], **kw)
print(f"wall {time.perf_counter() - t0:.4f} s")
for k, v in r.items():
    if k.startswith("time_"):
        print(f"  {k:50s} {1e3 * v:8.2f} ms")
pr = cProfile.Profile()
pr.enable()
learn_site_specific_rate_matrices(msa=msas[2], **kw)
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
print(s.getvalue())
