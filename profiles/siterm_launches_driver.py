"""One warmed learn_site_specific_rate_matrices(tree=None) call at 20 epochs, for an ncu launch list."""
import sys

import numpy as np

sys.path.insert(0, ".")
from benchlib.siterm import GRID_STEPS, _plant_family  # noqa: E402
from cherryml_b200.io import read_rate_matrix  # noqa: E402
from cherryml_b200.markov_chain import get_lg_path  # noqa: E402
from cherryml_b200.siterm import learn_site_specific_rate_matrices  # noqa: E402
from cherryml_b200.utils import amino_acids  # noqa: E402

rng = np.random.default_rng(0)
lg = read_rate_matrix(get_lg_path())
learn_site_specific_rate_matrices(
    tree=None, msa=_plant_family(rng), alphabet=list(amino_acids), regularization_rate_matrix=lg,
    regularization_strength=0.5, device="cuda:0", num_epochs=20, quantization_grid_num_steps=GRID_STEPS)
