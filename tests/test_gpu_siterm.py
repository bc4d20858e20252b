"""Batched per-site fit against goldens produced by the UNMODIFIED reference function
(tests/golden/make_golden_siterm.py) and against the reference's own recovery KATs
(_siterm/_site_specific_rate_matrix.py:1905-1920, 1984-1998)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from cherryml_b200.siterm import quantized_transitions_mle_vectorized_over_sites
from tests.conftest import GOLDEN

G = os.path.join(GOLDEN, "siterm")


def test_aa_with_initialisation_fp64_path():
    """With an initialisation the reference runs in fp64: tolerance 1e-6 relative."""
    g = np.load(os.path.join(G, "aa_init.npz"))
    r = quantized_transitions_mle_vectorized_over_sites(g["counts"], g["times"], num_epochs=30, initialization=g["init"])
    assert r["res"].shape == g["res"].shape
    ref_l, got_l = g["loss_per_epoch_per_site"], r["loss_per_epoch_per_site"]
    assert np.max(np.abs(got_l - ref_l) / np.abs(ref_l)) < 1e-6
    assert np.max(np.abs(r["loss_per_epoch"] - g["loss_per_epoch"]) / np.abs(g["loss_per_epoch"])) < 1e-6
    assert np.max(np.abs(r["res"] - g["res"])) < 1e-6 * np.max(np.abs(g["res"]))


def test_dna_without_initialisation_fp32_reference():
    """Without an initialisation the reference's parameters are fp32: tolerance 1e-4."""
    g = np.load(os.path.join(G, "dna_noinit.npz"))
    r = quantized_transitions_mle_vectorized_over_sites(g["counts"], g["times"], num_epochs=40, initialization=None)
    ref_l, got_l = g["loss_per_epoch_per_site"], r["loss_per_epoch_per_site"]
    assert np.max(np.abs(got_l - ref_l) / np.abs(ref_l)) < 1e-4
    assert np.max(np.abs(r["res"] - g["res"])) < 1e-3 * np.max(np.abs(g["res"]))


def test_zero_epochs_returns_the_initialisation():
    g = np.load(os.path.join(G, "dna_init_0epochs.npz"))
    r = quantized_transitions_mle_vectorized_over_sites(g["counts"], g["times"], num_epochs=0, initialization=g["init"])
    assert np.max(np.abs(r["res"] - g["res"])) < 1e-6
    assert np.mean((r["res"] - g["init"]) ** 2) < 1e-6  # the reference's own assertion


def test_recovers_the_true_rate_matrices():
    """The reference's KAT: counts := expm(t Q_true); 100 epochs from a random start recover
    Q_true with mean squared error < 1e-3."""
    g = np.load(os.path.join(G, "dna_noinit.npz"))
    r = quantized_transitions_mle_vectorized_over_sites(g["counts"], g["times"], num_epochs=100, initialization=None)
    assert np.mean((g["Q_true"] - r["res"]) ** 2) < 1e-3
