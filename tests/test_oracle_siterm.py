"""The SiteRM fit oracle against outputs of the unmodified reference function
(tests/golden/siterm, made by tests/golden/make_golden_siterm.py)."""
import os

import numpy as np

from oracle.siterm_oracle import fit_sites
from tests.conftest import GOLDEN

G = os.path.join(GOLDEN, "siterm")


def test_with_initialisation_matches_reference_run():
    g = np.load(os.path.join(G, "aa_init.npz"))
    r = fit_sites(g["counts"], g["times"], 30, g["init"])
    assert np.max(np.abs(r["loss_per_epoch_per_site"] - g["loss_per_epoch_per_site"])
                  / np.abs(g["loss_per_epoch_per_site"])) < 1e-9
    assert np.max(np.abs(r["res"] - g["res"])) < 1e-8 * np.max(np.abs(g["res"]))


def test_without_initialisation_matches_reference_run():
    g = np.load(os.path.join(G, "dna_noinit.npz"))
    r = fit_sites(g["counts"], g["times"], 40, None)
    assert np.max(np.abs(r["loss_per_epoch_per_site"] - g["loss_per_epoch_per_site"])
                  / np.abs(g["loss_per_epoch_per_site"])) < 1e-5
    assert np.max(np.abs(r["res"] - g["res"])) < 1e-4 * np.max(np.abs(g["res"]))


def test_quantization_idx_array_equals_scalar_definition():
    from cherryml_b200.utils import quantization_idx, quantization_idx_array

    rng = np.random.default_rng(0)
    q = [0.03 * 1.1 ** i for i in range(-64, 65)]
    t = np.exp(rng.uniform(np.log(1e-5), np.log(30), 5000))
    t[:129] = q                                                   # exactly on grid points
    t[129:257] = np.sqrt(np.array(q[:-1]) * np.array(q[1:]))      # near the decision boundaries
    got = quantization_idx_array(t, q)
    exp = [quantization_idx(float(x), q) for x in t]
    assert got.tolist() == [-1 if e is None else e for e in exp]
