"""Regenerate tests/golden/counting from the reference checkout (run in the build container).

    python tests/golden/make_golden.py

* copies the reference's own counting fixtures and their golden count matrices
  (tests/counting_tests/test_input_data/{tiny,tiny_2,tiny_3,tiny_4}) verbatim -- data, not code;
* copies three small families of its ``medium`` fixture (real Pfam-shaped data, ~1000
  sequences) and produces golden count matrices for them by RUNNING THE REFERENCE ITSELF:
  the unmodified C++ binaries compiled into oracle/_ref (oracle/build_ref.py) and the
  unmodified Python implementation imported from /root/reference (with import stubs for
  packages that are not installed here).
The GPU box has no /root/reference; the tests only read what this script wrote.
"""
import os
import shutil
import subprocess
import sys
import tempfile
import types
from unittest import mock

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"
SRC = os.path.join(REF, "tests/counting_tests/test_input_data")
DST = os.path.join(REPO, "tests/golden/counting")
MEDIUM3 = ["1a92_1_A", "1a4p_1_A", "1a64_1_A"]
AMINO = list("ARNDCQEGHILKMFPSTWYV")
GRID_LG = [0.06 * 1.1**i for i in range(-51, 51, 1)]
GRID_CO = [0.06 * 2.0**i for i in range(-5, 5, 1)]


def copy_tiny():
    for name in ("tiny", "tiny_2", "tiny_3", "tiny_4"):
        dst = os.path.join(DST, name)
        if os.path.exists(dst):
            shutil.rmtree(dst)
        shutil.copytree(os.path.join(SRC, name), dst)
        for root, _, files in os.walk(dst):
            os.chmod(root, 0o755)
            for f in files:
                os.chmod(os.path.join(root, f), 0o644)


def copy_medium3():
    for sub in ("tree_dir", "msa_dir", "msa_with_anc_dir", "site_rates_dir", "contact_map_dir"):
        d = os.path.join(DST, "medium3", sub)
        os.makedirs(d, exist_ok=True)
        for fam in MEDIUM3:
            shutil.copyfile(os.path.join(SRC, "medium", sub, fam + ".txt"), os.path.join(d, fam + ".txt"))
            os.chmod(os.path.join(d, fam + ".txt"), 0o644)


def run_ref_binary(binary, third_dir, grid, mode, out_dir, min_dist=None, msa_sub="msa_dir"):
    m3 = os.path.join(DST, "medium3")
    os.makedirs(out_dir, exist_ok=True)
    with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False) as f:
        f.write(" ".join(MEDIUM3))
        fams_path = f.name
    cmd = [os.path.join(REPO, "oracle/_ref", binary), os.path.join(m3, "tree_dir"),
           os.path.join(m3, msa_sub), os.path.join(m3, third_dir), str(len(MEDIUM3)),
           str(len(AMINO)), str(len(grid)), fams_path, *AMINO, *[str(q) for q in grid], mode]
    if min_dist is not None:
        cmd.append(str(min_dist))
    cmd.append(out_dir)
    subprocess.run(cmd, check=True)
    os.remove(fams_path)
    for f in os.listdir(out_dir):
        if f != "result.txt":
            os.remove(os.path.join(out_dir, f))


def import_reference():
    """Import the unmodified reference package with stubs for absent third-party modules."""
    for name in ("ete3", "matplotlib", "matplotlib.pyplot", "matplotlib.patches", "seaborn",
                 "biotite", "biotite.structure", "biotite.structure.io", "biotite.structure.io.pdb",
                 "wget", "parameterized"):
        sys.modules.setdefault(name, mock.MagicMock())
    # the Cython extension is only needed by SiteRM; stub it so `import cherryml` works
    stub = types.ModuleType("cherryml._siterm.fast_site_rates")
    stub.compute_optimal_site_rates = None
    sys.modules.setdefault("cherryml._siterm.fast_site_rates", stub)
    sys.path.insert(0, REF)
    import cherryml  # noqa: F401

    return cherryml


def run_ref_python():
    import_reference()
    from cherryml.counting import _count_co_transitions as co
    from cherryml.counting import _count_transitions as lg
    from cherryml.io import write_count_matrices

    m3 = os.path.join(DST, "medium3")
    for mode, tag, msa_sub in (("cherry++", "cherries_plus_plus", "msa_dir"), ("cherry", "cherries", "msa_dir"),
                               ("edge", "edges", "msa_with_anc_dir")):
        res = lg._map_func([os.path.join(m3, "tree_dir"), os.path.join(m3, msa_sub),
                            os.path.join(m3, "site_rates_dir"), MEDIUM3, AMINO, GRID_LG, mode])
        write_count_matrices(res, os.path.join(m3, f"refpy_count_matrices_dir_{tag}", "result.txt"))
    res = co._map_func([os.path.join(m3, "tree_dir"), os.path.join(m3, "msa_dir"),
                        os.path.join(m3, "contact_map_dir"), MEDIUM3, AMINO, GRID_CO, "cherry++", 7])
    write_count_matrices(res, os.path.join(m3, "refpy_count_co_matrices_dir_cherries_plus_plus", "result.txt"))


def compress_co_goldens():
    """400x400 goldens are ~4 MB of text each: keep them as compressed .npz (q, counts)."""
    import numpy as np

    sys.path.insert(0, REPO)
    from oracle.counting_oracle import read_count_matrices_text

    m3 = os.path.join(DST, "medium3")
    for d in sorted(os.listdir(m3)):
        if "count_co_matrices" not in d:
            continue
        txt = os.path.join(m3, d, "result.txt")
        q, states, counts = read_count_matrices_text(txt)
        np.savez_compressed(os.path.join(m3, d, "result.npz"), q=q, counts=counts)
        os.remove(txt)


def main():
    os.makedirs(DST, exist_ok=True)
    copy_tiny()
    copy_medium3()
    m3 = os.path.join(DST, "medium3")
    for mode, tag, msa_sub in (("cherry++", "cherries_plus_plus", "msa_dir"), ("cherry", "cherries", "msa_dir"),
                               ("edge", "edges", "msa_with_anc_dir")):
        run_ref_binary("count_transitions", "site_rates_dir", GRID_LG, mode,
                       os.path.join(m3, f"refcpp_count_matrices_dir_{tag}"), msa_sub=msa_sub)
        run_ref_binary("count_co_transitions", "contact_map_dir", GRID_CO, mode,
                       os.path.join(m3, f"refcpp_count_co_matrices_dir_{tag}"), min_dist=7, msa_sub=msa_sub)
    run_ref_python()
    compress_co_goldens()
    print("golden fixtures written to", DST)


if __name__ == "__main__":
    main()
