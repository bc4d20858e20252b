"""JTT-IPW initialiser against the reference's hand-verified goldens
(tests/estimation_tests/jtt_ipw_test.py:11-74; goldens copied to tests/golden/jtt)."""
import os

import numpy as np
import pytest
import torch

from cherryml_b200.estimation._jtt_ipw import jtt_ipw_from_counts
from cherryml_b200.io import read_count_matrices_array, read_mask_matrix
from tests.conftest import GOLDEN

INP = os.path.join(GOLDEN, "fit", "inputs")


@pytest.mark.parametrize("use_ipw", [True, False])
@pytest.mark.parametrize("masked", [False, True])
def test_jtt_ipw_toy_goldens(use_ipw, masked):
    q, _, counts = read_count_matrices_array(os.path.join(INP, "matrices_toy.txt"))
    mask = read_mask_matrix(os.path.join(INP, "3x3_mask.txt")).to_numpy() if masked else None
    got = jtt_ipw_from_counts(q, torch.from_numpy(counts), mask=mask, use_ipw=use_ipw)
    name = f"Q1_JTT{'-IPW' if use_ipw else ''}_on_toy_matrix{'_mask' if masked else ''}.txt"
    np.testing.assert_almost_equal(got, np.loadtxt(os.path.join(GOLDEN, "jtt", name)))


def test_stage_reads_counts_from_a_file_and_writes_the_reference_files(tmp_path):
    """The stage on counts that come from a file (host-side closed form, no device involved)."""
    from cherryml_b200.estimation import jtt_ipw
    from cherryml_b200.io import read_rate_matrix

    out = str(tmp_path / "jtt")
    jtt_ipw(count_matrices_path=os.path.join(INP, "matrices_toy.txt"), mask_path=None, use_ipw=True,
            output_rate_matrix_dir=out, normalize=False)
    got = read_rate_matrix(os.path.join(out, "result.txt"))
    np.testing.assert_almost_equal(got.to_numpy(), np.loadtxt(os.path.join(GOLDEN, "jtt", "Q1_JTT-IPW_on_toy_matrix.txt")))
    assert open(os.path.join(out, "profiling.txt")).read().startswith("Total time: ")
    out_n = str(tmp_path / "jtt_n")
    jtt_ipw(count_matrices_path=os.path.join(INP, "matrices_toy.txt"), mask_path=None, use_ipw=True,
            output_rate_matrix_dir=out_n, normalize=True)
    qn = read_rate_matrix(os.path.join(out_n, "result.txt")).to_numpy()
    from cherryml_b200.markov_chain import compute_mutation_rate

    assert abs(compute_mutation_rate(qn) - 1.0) < 1e-9
