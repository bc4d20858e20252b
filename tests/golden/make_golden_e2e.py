"""Regenerate the end-to-end goldens (BASELINE.json configs 1 and 2) by RUNNING THE UNMODIFIED
REFERENCE through its own public API on its own demo data (build container only).

    python tests/golden/make_golden_e2e.py

Writes
* tests/golden/demo_data.tar.xz       -- the reference's demo_data/{msas,trees,site_rates,contact_maps}
                                         (32 Pfam-shaped families; data, not code), and
* tests/golden/e2e/{lg,coevolution}.npz -- what ``cherryml_public_api`` produced: the count
  tensor, the JTT-IPW initialisation, the per-epoch losses and the learned rate matrix.
The reference runs with its Python counting implementation (its C++ program needs mpirun and a
writable package directory) and one counting process (its Pool uses the spawn start method,
which would re-import the package without the import stubs).
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests", "golden"))
REF = "/root/reference"
DEMO = os.path.join(REF, "demo_data")
OUT = os.path.join(REPO, "tests/golden/e2e")
CO_EPOCHS = 12


def main():
    from make_golden_fit import import_reference

    import_reference()
    import pandas as pd
    from cherryml import cherryml_public_api
    from cherryml.io import read_count_matrices, read_rate_matrix

    os.makedirs(OUT, exist_ok=True)
    subprocess.run(["tar", "-cJf", os.path.join(REPO, "tests/golden/demo_data.tar.xz"), "-C", DEMO,
                    "msas", "trees", "site_rates", "contact_maps"], check=True)

    def collect(cache_dir, out_file, output_path):
        def only(sub):
            d = os.path.join(cache_dir, sub)
            (h,) = os.listdir(d)
            return os.path.join(d, h)

        count_fn = "count_transitions" if "lg" in out_file else "count_co_transitions"
        counts = read_count_matrices(os.path.join(only(count_fn), "output_count_matrices_dir", "result.txt"))
        q = np.array([x[0] for x in counts])
        c = np.stack([x[1].to_numpy() for x in counts])
        jtt = read_rate_matrix(os.path.join(only("jtt_ipw"), "output_rate_matrix_dir", "result.txt")).to_numpy()
        mle = os.path.join(only("quantized_transitions_mle"), "output_rate_matrix_dir")
        loss = pd.read_csv(os.path.join(mle, "df_res.txt"))["loss"].to_numpy()
        learned = read_rate_matrix(output_path).to_numpy()
        if c.shape[-1] == 400:  # sparse storage for the 129 x 400 x 400 tensor
            idx = np.nonzero(c)
            np.savez_compressed(out_file, q=q, counts_idx=np.stack(idx).astype(np.int32), counts_val=c[idx],
                                counts_shape=np.array(c.shape), jtt_ipw=jtt, loss=loss,
                                learned=learned.astype(np.float32))
        else:
            np.savez_compressed(out_file, q=q, counts=c, jtt_ipw=jtt, loss=loss, learned=learned.astype(np.float32))
        print(out_file, "loss", loss[0], "->", loss[-1])

    with tempfile.TemporaryDirectory() as tmp:
        out = os.path.join(tmp, "lg.txt")
        cherryml_public_api(
            output_path=out, model_name="LG", msa_dir=f"{DEMO}/msas", tree_dir=f"{DEMO}/trees",
            site_rates_dir=f"{DEMO}/site_rates", cache_dir=os.path.join(tmp, "cache"), num_processes_counting=1,
            use_cpp_counting_implementation=False, num_epochs=500,
        )
        collect(os.path.join(tmp, "cache"), os.path.join(OUT, "lg.npz"), out)
    with tempfile.TemporaryDirectory() as tmp:
        out = os.path.join(tmp, "co.txt")
        cherryml_public_api(
            output_path=out, model_name="co-evolution", msa_dir=f"{DEMO}/msas", contact_map_dir=f"{DEMO}/contact_maps",
            tree_dir=f"{DEMO}/trees", cache_dir=os.path.join(tmp, "cache"), num_processes_counting=1,
            num_processes_optimization=8, use_cpp_counting_implementation=False, num_epochs=CO_EPOCHS,
        )
        collect(os.path.join(tmp, "cache"), os.path.join(OUT, "coevolution.npz"), out)


if __name__ == "__main__":
    main()
