"""The bench's reference fit arm (benchlib/fit.py reference_fit_arms -> oracle/run_reference_fit.py): the UNMODIFIED
reference quantized_transitions_mle, from the packed copy under oracle/_ref, on count matrices and an initialisation
written by this package's writers.  CPU arm only here (no GPU); the cuda arm runs in bench.py on the GPU box."""
import numpy as np
import pytest
import torch

from oracle.ref_package import reference_available


@pytest.mark.skipif(not reference_available(), reason="oracle/_ref/reference_package.tar.gz not built (needs /root/reference)")
def test_reference_cpu_arm_trains_on_our_files(tmp_path):
    from benchlib.fit import reference_fit_arms

    rng = np.random.default_rng(5)
    S, K, epochs = 20, 3, 6
    times = [0.05, 0.4, 2.0]
    half = rng.integers(1, 60, size=(K, S, S)).astype(np.float64)
    counts = half + half.transpose(0, 2, 1)
    for k in range(K):  # more mass on the diagonal for short times, like real cherries
        counts[k] += np.diag(rng.integers(500, 900, size=S) / (1 + 3 * k)).round()
    out = reference_fit_arms(times, torch.from_numpy(counts), num_epochs_full=epochs, cpu_epochs=epochs, cuda_epochs=0,
                             workdir=str(tmp_path), timeout_s=170)
    assert "cpu" in out and "error" not in out["cpu"], out
    arm = out["cpu"]
    assert arm["kind"] == "reference" and arm["device"] == "cpu" and arm["epochs_timed"] == epochs
    assert not arm["extrapolated"] and arm["seconds_end_to_end"] > 0
    assert arm["loss_last"] < arm["loss_first"]
