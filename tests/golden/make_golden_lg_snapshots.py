"""Snapshots of the UNMODIFIED reference's LG demo fit (build container only):

    python tests/golden/make_golden_lg_snapshots.py

Runs the reference's ``quantized_transitions_mle`` on the count matrices and the JTT-IPW initialisation of
tests/golden/e2e/lg.npz (what its public API feeds the stage for BASELINE config 1: 500 Adam epochs, fp32) and
stores every rate matrix file the stage writes besides ``result.txt``: ``Q_last.txt`` and the power-of-two
snapshots ``Q_<2^j>.txt``, plus the epoch whose loss is the smallest (the iterate ``result.txt`` holds).
tests/test_gpu_e2e.py compares our fit with these AT THE SAME EPOCHS, which separates "the trajectories agree to
fp32 tolerance" from "the two arg-mins fall on different iterates of a flat loss".
"""
import os
import sys
import tempfile

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests", "golden"))
OUT = os.path.join(REPO, "tests/golden/e2e/lg_snapshots.npz")


def main():
    from make_golden_fit import import_reference

    import_reference()
    import pandas as pd
    from cherryml.estimation import quantized_transitions_mle
    from cherryml.io import read_rate_matrix, write_count_matrices, write_rate_matrix
    from cherryml.utils import get_amino_acids

    g = np.load(os.path.join(REPO, "tests/golden/e2e/lg.npz"))
    states = get_amino_acids()
    with tempfile.TemporaryDirectory() as tmp:
        cpath, ipath, odir = os.path.join(tmp, "c.txt"), os.path.join(tmp, "i.txt"), os.path.join(tmp, "out")
        write_count_matrices([[float(q), pd.DataFrame(c, index=states, columns=states)] for q, c in zip(g["q"], g["counts"])], cpath)
        write_rate_matrix(g["jtt_ipw"], states, ipath)
        quantized_transitions_mle(
            count_matrices_path=cpath, initialization_path=ipath, mask_path=None, output_rate_matrix_dir=odir,
            stationary_distribution_path=None, rate_matrix_parameterization="pande_reversible", device="cpu",
            learning_rate=1e-1, num_epochs=500, do_adam=True)
        loss = pd.read_csv(os.path.join(odir, "df_res.txt"))["loss"].to_numpy()
        out = {"loss": loss, "best_epoch": int(np.argmin(loss)),
               "result": read_rate_matrix(os.path.join(odir, "result.txt")).to_numpy()}
        for f in sorted(os.listdir(odir)):
            if f.startswith("Q_") and f.endswith(".txt"):
                out[f[:-4]] = read_rate_matrix(os.path.join(odir, f)).to_numpy()
        assert np.max(np.abs(loss - g["loss"])) <= 1e-6 * np.max(np.abs(g["loss"])), "not the public API's trajectory"
        # The same fit through the oracle's restatement of the reference's arithmetic: in fp32 it IS the run above,
        # in fp64 it is what the reference would compute without its two float casts.  Stored: the fp64 last
        # iterate and loss trace (ground truth for the fp64 CUDA path) and how far fp32 drifts from it.
        import torch

        from oracle.fit_oracle import fit_oracle

        o32 = fit_oracle(list(g["q"]), g["counts"], None, g["jtt_ipw"], 0.1, 500, dtype=torch.float32)
        drift32 = float(np.max(np.abs(o32["Q_last"] - out["Q_last"])))
        assert drift32 <= 1e-6 * np.max(np.abs(out["Q_last"])), f"the fp32 oracle is not the reference run: {drift32}"
        o64 = fit_oracle(list(g["q"]), g["counts"], None, g["jtt_ipw"], 0.1, 500, dtype=torch.float64)
        out["fp64_Q_last"] = o64["Q_last"]
        out["fp64_loss"] = np.asarray(o64["loss"])
        out["fp32_vs_fp64_Q_last"] = np.array(np.max(np.abs(o64["Q_last"] - out["Q_last"])))
        # how well conditioned is the iterate after 500 Adam steps?  The same fp64 fit from an initialisation
        # perturbed by 1e-14 relative: its distance is the floor for ANY two fp64 implementations.
        rng = np.random.default_rng(0)
        pert = g["jtt_ipw"] * (1.0 + 1e-14 * rng.standard_normal(g["jtt_ipw"].shape))
        o64p = fit_oracle(list(g["q"]), g["counts"], None, pert, 0.1, 500, dtype=torch.float64)
        out["fp64_sensitivity_Q_last"] = np.array(np.max(np.abs(o64p["Q_last"] - o64["Q_last"])))
        print("fp64 last iterate under a 1e-14 perturbation of the initialisation moves by", out["fp64_sensitivity_Q_last"])
        print("fp32 reference vs fp64 arithmetic at the last epoch:", out["fp32_vs_fp64_Q_last"])
        np.savez_compressed(OUT, **out)
        print({k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})


if __name__ == "__main__":
    main()
