"""Regenerate the FastCherries goldens by RUNNING THE UNMODIFIED REFERENCE PROGRAM
(oracle/_ref/fast_cherries, compiled by oracle/build_ref.py from the sources under
/root/reference; build container only).

    python tests/golden/make_golden_fast_cherries.py

Writes tests/golden/fast_cherries/cases.json.gz: for every case the MSA (inline text, or the name
of a family in tests/golden/demo_data.tar.xz), the labelled rate matrix text, the parameters,
and what the program wrote: the cherries file (names + '%.17f' distances) and the site-rates
file.  Cases: demo families (Pfam-shaped, ~1000 sequences), the reference's own
tests/phylogeny_estimation_tests/different_alphabet fixture, the MSA of the program's C++ tests,
and seeded small MSAs covering odd sizes, 2-3 sequences, duplicates, all-gap columns/rows
(a single-sequence MSA makes the reference program crash: no golden).
"""
import json
import os
import sys
import tempfile

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
REF = "/root/reference"
OUT = os.path.join(REPO, "tests/golden/fast_cherries")
AA = "ARNDCQEGHILKMFPSTWYV"


def synthetic_msa(rng, n, L, gap=0.15, mut=0.2, dup=False, gap_cols=(), gap_rows=()):
    root = rng.integers(0, 20, L)
    rows = []
    for i in range(n):
        r = rows[rng.integers(0, len(rows))].copy() if rows and rng.random() < 0.7 else root.copy()
        flip = rng.random(L) < mut
        r[flip] = rng.integers(0, 20, int(flip.sum()))
        rows.append(r)
    out = []
    for i, r in enumerate(rows):
        s = np.array(list(AA))[r]
        s[rng.random(L) < gap] = "-"
        for c in gap_cols:
            s[c] = "-"
        if i in gap_rows:
            s[:] = "-"
        out.append("".join(s))
    if dup and n >= 4:
        out[n - 1] = out[0]
        out[n - 2] = out[1]
    return "".join(f">seq{i}\n{s}\n" for i, s in enumerate(out))


def main():
    from oracle.fast_cherries_oracle import run_reference_binary
    from oracle.build_ref import build_reference_binaries

    assert build_reference_binaries()
    os.makedirs(OUT, exist_ok=True)
    lg = open(os.path.join(REF, "data/rate_matrices/lg.txt")).read()
    equ = open(os.path.join(REF, "data/rate_matrices/equ.txt")).read()
    weird = open(os.path.join(REF, "tests/phylogeny_estimation_tests/weird_rate_matrix.txt")).read()
    cases = []

    def add(name, msa_text, demo_family, Q_text, R, max_iters=50, seed=1234, num_steps=64):
        with tempfile.TemporaryDirectory() as tmp:
            if demo_family is not None:
                msa_path = os.path.join(REF, "demo_data/msas", demo_family + ".txt")
            else:
                msa_path = os.path.join(tmp, "msa.txt")
                open(msa_path, "w").write(msa_text)
            qp = os.path.join(tmp, "Qlab.txt")
            open(qp, "w").write(Q_text)
            (res,) = run_reference_binary([msa_path], qp, R, max_iters=max_iters, seed=seed, num_steps=num_steps)
        cases.append({
            "name": name, "msa_text": msa_text, "demo_family": demo_family, "rate_matrix_text": Q_text,
            "num_rate_categories": R, "max_iters": max_iters, "seed": seed, "num_steps": num_steps,
            "cherries_file": res["output_text"], "site_rates_file": res["site_rates_file_text"],
        })

    for fam, R in [("13gs_1_A", 20), ("13gs_1_A", 1), ("1a0b_1_A", 20), ("1a0b_1_A", 4), ("1a2t_1_A", 20),
                   ("1a12_1_A", 1)]:
        add(f"demo_{fam}_R{R}", None, fam, lg, R)
    add("demo_13gs_equ_R1_seed7", None, "13gs_1_A", equ, 1, seed=7)
    add("different_alphabet", open(os.path.join(REF, "tests/phylogeny_estimation_tests/different_alphabet/msa.txt")).read(),
        None, weird, 4)
    add("cpp_tests_msa", open(os.path.join(
        REF, "cherryml/phylogeny_estimation/FastCherries/tests/Aln0000_txt-gb_phyml.txt")).read(), None, lg, 4)
    rng = np.random.default_rng(0)
    for n, L, kw in [(2, 10, {}), (3, 17, {}), (4, 33, {}), (5, 16, {}), (7, 40, {"dup": True}),
                     (16, 48, {"gap_cols": (0, 5)}), (33, 100, {"gap_rows": (3,)}), (64, 65, {"gap": 0.5}),
                     (101, 130, {"mut": 0.02}), (257, 31, {"mut": 0.5})]:
        for R in (1, 4, 20):
            add(f"synthetic_n{n}_L{L}_R{R}", synthetic_msa(rng, n, L, **kw), None, lg, R)
    add("synthetic_max_iters_1", synthetic_msa(rng, 40, 60), None, lg, 20, max_iters=1)
    add("synthetic_max_iters_0", synthetic_msa(rng, 40, 60), None, lg, 20, max_iters=0)
    add("synthetic_grid_8", synthetic_msa(rng, 40, 60), None, lg, 4, num_steps=8)
    # other alphabets: DNA (4 states) and amino acids + a gap STATE (21 states)
    def labelled(states, M):
        return "\t" + "\t".join(states) + "\n" + "".join(
            s + "\t" + "\t".join(repr(float(v)) for v in M[i]) + "\n" for i, s in enumerate(states))

    hky = np.array([[0, 1, 4, 1], [1, 0, 1, 4], [4, 1, 0, 1], [1, 4, 1, 0]], dtype=float) * np.array([0.3, 0.2, 0.2, 0.3])
    np.fill_diagonal(hky, -hky.sum(axis=1))
    hky /= -(np.array([0.3, 0.2, 0.2, 0.3]) * np.diag(hky)).sum()
    dna = synthetic_msa(rng, 30, 200, gap=0.1, mut=0.15)
    dna = "\n".join(ln if ln.startswith(">") else ln.translate(str.maketrans(AA, "ACGT" * 5)) for ln in dna.split("\n"))
    add("dna_hky_n30_L200_R4", dna, None, labelled(list("ACGT"), hky), 4)
    lg_rows = [[float(v) for v in ln.split()[1:]] for ln in lg.strip().split("\n")[1:]]
    Q21 = np.zeros((21, 21))
    Q21[:20, :20] = np.array(lg_rows)
    np.fill_diagonal(Q21, 0)
    Q21[:20, 20] = 0.03
    Q21[20, :20] = 0.4
    np.fill_diagonal(Q21, -Q21.sum(axis=1))
    add("aa_plus_gap_state_n25_L60_R20", synthetic_msa(rng, 25, 60, gap=0.25), None, labelled(list(AA) + ["-"], Q21), 20)
    add("long_n12_L1300_R4", synthetic_msa(rng, 12, 1300), None, lg, 4)
    add("many_n1500_L24_R1", synthetic_msa(rng, 1500, 24, mut=0.08), None, lg, 1)
    import gzip

    with gzip.open(os.path.join(OUT, "cases.json.gz"), "wt") as f:
        json.dump(cases, f)
    print(len(cases), "cases", os.path.getsize(os.path.join(OUT, "cases.json.gz")), "bytes")


if __name__ == "__main__":
    main()
