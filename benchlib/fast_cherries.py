"""bench.py's FastCherries section: both kernels on synthetic Pfam-shaped families (BASELINE
config 3's shape), the same through host buffers, and -- on rank 0 at N = 1 -- the unmodified
reference program (oracle/_ref/fast_cherries) on a text rendering of a sample of the same
families, one process per host core, with its outputs compared to ours."""
import math
import os
import shutil
import subprocess
import tempfile
import time
from typing import Dict, Optional

import numpy as np
import torch

from cherryml_b200.io import read_rate_matrix
from cherryml_b200.markov_chain import get_lg_path
from cherryml_b200.utils import amino_acids
from cherryml_b200.phylogeny_estimation import _fast_cherries as fc
from benchlib.hostcores import usable_cores

N_SEQS, N_SITES, N_RATE_CATS, MAX_ITERS, SEED = 1024, 300, 20, 50, 1234


def _render(msa: np.ndarray, fams: np.ndarray, out_dir: str, n: int):
    letters = np.frombuffer(("".join(amino_acids) + "-").encode(), dtype=np.uint8)
    paths = []
    for f in range(n):
        fam = fams[f]
        rows = msa[int(fam["msa_off"]): int(fam["msa_off"]) + int(fam["n_seqs"]) * int(fam["row_stride"])]
        rows = rows.reshape(int(fam["n_seqs"]), -1)[:, : int(fam["n_sites"])]
        text = letters[rows]
        path = os.path.join(out_dir, f"fam{f}.txt")
        with open(path, "w") as fh:
            fh.write("".join(f">seq{i}\n{text[i].tobytes().decode()}\n" for i in range(text.shape[0])))
        paths.append(path)
    return paths


def _reference_program(paths, tmp: str, cores: int) -> Optional[Dict]:
    """Runs the reference program like _fast_cherries.py:85-104 does, one process per core on the
    wrapper's own striping (get_process_args: paths[r::P])."""
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ref_bin = os.path.join(repo, "oracle", "_ref", "fast_cherries")
    if not (os.path.exists(ref_bin) and os.access(ref_bin, os.X_OK)):
        return None
    lines = open(get_lg_path()).read().strip().split("\n")
    with open(os.path.join(tmp, "Q.txt"), "w") as f:
        f.writelines([ln[1:] + "\n" for ln in lines[1:]])
    alphabet = lines[0].split()
    with open(os.path.join(tmp, "alphabet.txt"), "w") as f:
        f.write(str(len(alphabet)) + " " + " ".join(alphabet))

    def write_list(items, fn):
        with open(fn, "w") as f:
            f.write(str(len(items)) + "\n" + "\n".join(items))

    procs_n = min(cores, len(paths))
    cmds = []
    for r in range(procs_n):
        mine = paths[r::procs_n]
        for kind, ext in (("msas", ""), ("outs", ".output"), ("profs", ".profiling"), ("rates", ".rates")):
            write_list([p + ext for p in mine], os.path.join(tmp, f"{kind}_{r}.txt"))
        cmds.append([
            ref_bin, "-seed", str(SEED), "-quantization_grid_center", "0.03", "-quantization_grid_step", "1.1",
            "-quantization_grid_num_steps", "64", "-output_list_path", os.path.join(tmp, f"outs_{r}.txt"),
            "-rate_matrix_path", os.path.join(tmp, "Q.txt"), "-msa_list_path", os.path.join(tmp, f"msas_{r}.txt"),
            "-profiling_list_path", os.path.join(tmp, f"profs_{r}.txt"), "-site_rate_list_path",
            os.path.join(tmp, f"rates_{r}.txt"), "-num_rate_categories_ble", str(N_RATE_CATS), "-max_iters_ble",
            str(MAX_ITERS), "-alphabet_path", os.path.join(tmp, "alphabet.txt"),
        ])
    t0 = time.perf_counter()
    running = [subprocess.Popen(c, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for c in cmds]
    rcs = [p.wait() for p in running]
    seconds = time.perf_counter() - t0
    if any(rcs):
        return None
    return {"seconds": seconds, "processes": procs_n}


def bench_fast_cherries(device, families: int = 2048, reps: int = 3, cpu_baseline: bool = True,
                        cpu_families: int = 0, seed: int = 0) -> Dict:
    from cherryml_b200.synthetic import synthetic_fc

    msa, fams = synthetic_fc(families, N_SEQS, N_SITES, seed=seed)
    Q = read_rate_matrix(get_lg_path()).to_numpy(dtype=np.float64)
    grid = fc.quantization_grid(0.03, 1.1, 64)
    cats = fc.ble_rate_categories(N_RATE_CATS)
    weights = fc.initial_rate_weights(cats)
    priors = np.array([2 * math.log(r) - 3 * r for r in cats])
    pinned = torch.from_numpy(msa).pin_memory()
    msa_p = pinned.numpy()
    t0 = time.perf_counter()
    table = fc.log_transition_table(Q, grid, cats, device)
    torch.cuda.synchronize()
    table_s = time.perf_counter() - t0
    best = None
    for _ in range(reps + 1):  # first pass = warm-up
        t0 = time.perf_counter()
        out = fc.fast_cherries_device(msa_p, fams, 20, table, priors, weights, SEED, MAX_ITERS, device)
        wall = time.perf_counter() - t0
        if best is None or wall < best["wall"]:
            best = dict(out, wall=wall)
    residues = families * N_SEQS * N_SITES
    res = {
        "workload": f"{families} families x {N_SEQS} seqs x {N_SITES} sites per GPU, {N_RATE_CATS} rate categories, "
                    f"129 grid points, max_iters {MAX_ITERS}, LG rate matrix",
        "pair_kernel_ms": best["pair_ms"], "ble_kernel_ms": best["ble_ms"],
        "families_per_s_kernels": families / ((best["pair_ms"] + best["ble_ms"]) * 1e-3),
        "e2e_seconds": best["wall"], "families_per_s_e2e": families / best["wall"],
        "residues_per_s_e2e": residues / best["wall"],
        "e2e_note": "pinned host residue buffer -> H2D -> cherry_fc_pair + cherry_fc_ble -> D2H of cherries, "
                    "length indices and site categories",
        "h2d_bytes": int(msa.nbytes), "log_table_seconds": table_s,
        "ble_iterations_mean": float(best["iters"].mean()), "ble_iterations_max": int(best["iters"].max()),
        "gpu_launches": 3,
    }
    # MSAs -> cherries -> LG count tensor with the residues resident on the device (no text hand-off)
    from cherryml_b200.phylogeny_estimation._pipeline import fast_cherries_then_count_lg
    from cherryml_b200.synthetic import quantization_grid

    qp = quantization_grid()
    best_pipe = None
    for _ in range(3):
        t0 = time.perf_counter()
        counts, _ = fast_cherries_then_count_lg(msa_p, fams, amino_acids, Q, qp, num_rate_categories=N_RATE_CATS,
                                                max_iters=MAX_ITERS, seed=SEED, device=device)
        total = float(counts.sum().item())  # D2H of the result
        wall = time.perf_counter() - t0
        best_pipe = wall if best_pipe is None else min(best_pipe, wall)
    res["then_count_lg"] = {
        "seconds": best_pipe, "families_per_s": families / best_pipe, "transitions_counted": total,
        "note": "pinned host residues -> H2D -> FastCherries kernels -> cherry_fc_lengths_and_rates (host, exact "
                "text-file values) -> cherry_fc_relayout_lg -> bucket table + cherry_count_lg + symmetrise -> D2H; "
                "counts identical to the route through tree / site-rate files (tests)",
    }
    if cpu_baseline:
        cores = usable_cores()
        n = min(families, cpu_families or max(128, 16 * cores))
        tmp = tempfile.mkdtemp(prefix="cherry_fc_")
        try:
            paths = _render(msa, fams, tmp, n)
            ref = _reference_program(paths, tmp, cores)
            if ref is not None:
                same = True
                for f, p in enumerate(paths):
                    fam = fams[f]
                    c0, c1 = int(fam["cherry_off"]), int(fam["cherry_off"]) + int(fam["n_seqs"]) // 2
                    s0, s1 = int(fam["site_off"]), int(fam["site_off"]) + int(fam["n_sites"])
                    lengths, rates = fc.normalise_lengths_and_rates(best["len_idx"][c0:c1], best["site_cat"][s0:s1],
                                                                    grid, cats)
                    ours = "".join(f"seq{a}\nseq{b}\n{fc._fixed17(d)}\n" for a, b, d in
                                   zip(best["pair_a"][c0:c1], best["pair_b"][c0:c1], lengths))
                    ours_rates = f"{len(rates)} sites\n" + "".join(fc._fixed17(r) + " " for r in rates)
                    same = same and ours == open(p + ".output").read() and ours_rates == open(p + ".rates").read()
                res["cpu_baseline"] = {
                    "value": n / ref["seconds"], "unit": "families/s", "cores": ref["processes"], "kind": "reference",
                    "sample": f"{n} of the workload's families as MSA text files, unmodified reference FastCherries "
                              f"program, {ref['processes']} processes on the wrapper's family striping, "
                              "text parsing and table set-up included",
                    "seconds": ref["seconds"],
                }
                res["outputs_identical_to_reference_program_on_sample"] = bool(same)
            # our stage on the SAME text files: native read + encode -> H2D -> kernels -> D2H -> tree,
            # newick, site-rate, likelihood and profiling files (the reference program writes less: its
            # Python wrapper builds the tree files afterwards)
            out_dir = os.path.join(tmp, "ours")
            os.makedirs(out_dir, exist_ok=True)
            j = lambda ext: [os.path.join(out_dir, f"fam{f}{ext}") for f in range(n)]  # noqa: E731
            best_text = None
            for _ in range(3):
                t0 = time.perf_counter()
                with fc.NativeMsas(paths, amino_acids) as msas:
                    o = fc.fast_cherries_device(msas.msa, msas.fams, 20, table, priors, weights, SEED, MAX_ITERS, device)
                    t1 = time.perf_counter()
                    msas.write_outputs(o, grid, cats, j(".txt"), j(".newick"), j(".rates"), j(".ll"), j(".profiling"),
                                       np.zeros((n, 4)))
                wall = time.perf_counter() - t0
                if best_text is None or wall < best_text[0]:
                    best_text = (wall, t1 - t0)
            res["e2e_text"] = {
                "value": n / best_text[0], "unit": "families/s", "seconds": best_text[0],
                "seconds_read_and_device": best_text[1], "host_threads": cores,
                "sample": f"the same {n} MSA text files as cpu_baseline -> tree, newick, site-rate, likelihood and "
                          "profiling files (cherry_fc_read_msas -> cherry_fc_pair/_ble -> cherry_fc_write_outputs); "
                          "best of 3 passes, pooled page-locked staging buffer",
            }
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
    return res
