"""Regenerate tests/golden/fit by RUNNING THE UNMODIFIED REFERENCE (build container only).

    python tests/golden/make_golden_fit.py

For every case below the reference's own ``quantized_transitions_mle`` (CPU, fp32 expm as
shipped: cherryml/estimation/_quantized_transitions_mle.py:40-122 -> ratelearner.py:66-152
-> trainer.py:118-243) is run on committed inputs and its per-epoch losses (df_res.txt),
Q snapshots and result are stored as one compressed .npz per case, next to copies of the
small input files it consumed.  The reference is imported from /root/reference with stubs
for packages that are absent here (ete3, matplotlib, ...) and a pandas-3 shim for the
removed ``delim_whitespace`` keyword; none of the hot-path arithmetic is touched.
"""
import os
import shutil
import sys
import tempfile

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests", "golden"))
REF = "/root/reference"
DST = os.path.join(REPO, "tests/golden/fit")
TID = os.path.join(REF, "tests/test_input_data")


def import_reference():
    from make_golden import import_reference as base_import
    import pandas as pd

    _read_csv = pd.read_csv

    def read_csv(*args, **kwargs):
        if kwargs.pop("delim_whitespace", False):
            kwargs["sep"] = r"\s+"
        return _read_csv(*args, **kwargs)

    pd.read_csv = read_csv
    base_import()
    import matplotlib.pyplot as plt
    from unittest import mock

    plt.subplots = lambda *a, **k: (mock.MagicMock(), mock.MagicMock())
    # Series.plot needs a plotting backend: give pandas a no-op one
    import types

    backend = types.ModuleType("cherry_noop_backend")
    backend.plot = lambda *a, **k: None
    sys.modules["cherry_noop_backend"] = backend
    pd.options.plotting.backend = "cherry_noop_backend"


def write_counts_txt(path, q, states, counts):
    from cherryml_b200.io import write_count_matrices_array

    write_count_matrices_array(list(q), states, counts, path, "python")


def run_case(name, counts_path, init_path, mask_path, num_epochs, lr=0.1, do_adam=True, keep_q=True):
    from cherryml.estimation import quantized_transitions_mle
    from cherryml.io import read_rate_matrix

    case_dir = os.path.join(DST, name)
    os.makedirs(case_dir, exist_ok=True)
    with tempfile.TemporaryDirectory() as out:
        quantized_transitions_mle(
            count_matrices_path=counts_path, initialization_path=init_path, mask_path=mask_path,
            output_rate_matrix_dir=out, stationary_distribution_path=None,
            rate_matrix_parameterization="pande_reversible", device="cpu", learning_rate=lr,
            num_epochs=num_epochs, do_adam=do_adam, OMP_NUM_THREADS=8, OPENBLAS_NUM_THREADS=8,
        )
        import pandas as pd

        df = pd.read_csv(os.path.join(out, "df_res.txt"))
        arrays = {"loss": df["loss"].to_numpy(), "num_epochs": np.array(num_epochs), "lr": np.array(lr)}
        for f in sorted(os.listdir(out)):
            if f.startswith("Q_") or f == "result.txt":
                if keep_q or f in ("result.txt", "Q_last.txt", "Q_1.txt"):
                    arrays[f[:-4]] = read_rate_matrix(os.path.join(out, f)).to_numpy().astype(np.float32)  # printed from fp32
        np.savez_compressed(os.path.join(case_dir, "reference_run.npz"), **arrays)
    print(name, "loss[0], loss[-1] =", arrays["loss"][0], arrays["loss"][-1])


def main():
    import_reference()
    os.makedirs(DST, exist_ok=True)
    inputs = os.path.join(DST, "inputs")
    os.makedirs(inputs, exist_ok=True)
    for f in ("matrices_toy.txt", "3x3_pande_reversible_initialization.txt",
              "3x3_pande_reversible_initialization_mask.txt", "3x3_mask.txt", "20x20_random_mask.txt"):
        shutil.copyfile(os.path.join(TID, f), os.path.join(inputs, f))
    shutil.copyfile(os.path.join(REF, "data/rate_matrices/lg.txt"), os.path.join(inputs, "lg.txt"))
    shutil.copyfile(os.path.join(REF, "data/rate_matrices/equ.txt"), os.path.join(inputs, "equ.txt"))
    # LG-shaped counts: the reference C++ program's output on three real families (102 buckets)
    lg_counts = os.path.join(REPO, "tests/golden/counting/medium3/refcpp_count_matrices_dir_cherries_plus_plus/result.txt")
    # 400x400 counts: reference C++ co-transition output on the same families (10 buckets)
    z = np.load(os.path.join(REPO, "tests/golden/counting/medium3/refcpp_count_co_matrices_dir_cherries_plus_plus/result.npz"))
    from cherryml_b200.utils import amino_acids

    pair_states = [a + b for a in amino_acids for b in amino_acids]
    co_counts = os.path.join(tempfile.mkdtemp(), "co_counts.txt")
    write_counts_txt(co_counts, z["q"], pair_states, z["counts"])

    # the two 400x400 inputs of the co-evolution cases, kept as one compressed archive
    from cherryml_b200.io import read_mask_matrix, read_rate_matrix

    np.savez_compressed(
        os.path.join(inputs, "co400_inputs.npz"),
        coevolution=read_rate_matrix(os.path.join(REF, "data/rate_matrices/coevolution/coevolution.txt")).to_numpy(),
        aa_coevolution_mask=read_mask_matrix(os.path.join(REF, "data/mask_matrices/aa_coevolution_mask.txt")).to_numpy().astype(np.int8),
    )
    run_case("toy3_init", f"{inputs}/matrices_toy.txt", f"{inputs}/3x3_pande_reversible_initialization.txt", None, 60)
    run_case("toy3_init_mask", f"{inputs}/matrices_toy.txt", f"{inputs}/3x3_pande_reversible_initialization_mask.txt",
             f"{inputs}/3x3_mask.txt", 60)
    run_case("toy3_noinit", f"{inputs}/matrices_toy.txt", None, None, 40)
    run_case("toy3_sgd", f"{inputs}/matrices_toy.txt", f"{inputs}/3x3_pande_reversible_initialization.txt", None, 20,
             lr=0.01, do_adam=False)
    run_case("lg20_init_equ", lg_counts, f"{inputs}/equ.txt", None, 200)
    run_case("lg20_init_lg", lg_counts, f"{inputs}/lg.txt", None, 100)
    run_case("lg20_noinit_mask", lg_counts, None, f"{inputs}/20x20_random_mask.txt", 50)
    run_case("co400_init", co_counts, os.path.join(REF, "data/rate_matrices/coevolution/coevolution.txt"), None, 6,
             keep_q=False)
    run_case("co400_noinit_mask", co_counts, None, os.path.join(REF, "data/mask_matrices/aa_coevolution_mask.txt"), 4,
             keep_q=False)


if __name__ == "__main__":
    main()
