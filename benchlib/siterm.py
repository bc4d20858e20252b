"""bench.py's SiteRM section (BASELINE config 5): plant-dataset-shaped families (38 sequences x
331 sites), FastCherries trees, batched per-site fits -- ``learn_site_specific_rate_matrices``
with ``tree=None`` end to end per family; beside it, on rank 0 at N = 1, the torch-CPU oracle
port of the reference's batched fit on the count tensors of one family."""
import time
from typing import Dict

import numpy as np

from cherryml_b200.io import read_rate_matrix
from cherryml_b200.markov_chain import get_lg_path
from cherryml_b200.utils import amino_acids
from benchlib.hostcores import usable_cores

N_SEQS, N_SITES, NUM_EPOCHS, GRID_STEPS = 38, 331, 100, 8


def _plant_family(rng) -> Dict[str, str]:
    aa = np.array(amino_acids)
    root = rng.integers(0, 20, N_SITES)
    div = rng.uniform(0.02, 0.6, N_SITES)  # slow and fast sites
    msa = {}
    for k in range(N_SEQS // 2):
        anc = np.where(rng.random(N_SITES) < div, rng.integers(0, 20, N_SITES), root)
        for j in range(2):
            leaf = np.where(rng.random(N_SITES) < 0.4 * div, rng.integers(0, 20, N_SITES), anc)
            s = aa[leaf]
            s[rng.random(N_SITES) < 0.08] = "-"
            msa[f"seq{2 * k + j}"] = "".join(s)
    return msa


def bench_siterm(device, families: int = 8, cpu_baseline: bool = True, seed: int = 0) -> Dict:
    from cherryml_b200.siterm import learn_site_specific_rate_matrices

    rng = np.random.default_rng(seed)
    lg = read_rate_matrix(get_lg_path())
    msas = [_plant_family(rng) for _ in range(families)]
    kw = dict(tree=None, alphabet=list(amino_acids), regularization_rate_matrix=lg, regularization_strength=0.5,
              device=str(device), num_epochs=NUM_EPOCHS, quantization_grid_num_steps=GRID_STEPS)
    learn_site_specific_rate_matrices(msa=msas[0], **kw)  # warm-up
    t0 = time.perf_counter()
    results = [learn_site_specific_rate_matrices(msa=m, **kw) for m in msas]
    wall = time.perf_counter() - t0
    keys = ("time_estimate_tree", "time_get_raw_count_matrices", "time_get_pseudocount_matrices", "time_compute_loss")
    out = {
        "workload": f"{families} plant-shaped families x {N_SEQS} seqs x {N_SITES} sites, FastCherries trees (20 rate "
                    f"categories), lambda 0.5, Q0 = LG, {2 * GRID_STEPS + 1} grid points, {NUM_EPOCHS} Adam epochs, "
                    "one learn_site_specific_rate_matrices call per family",
        "seconds_per_family": wall / families, "sites_per_s": families * N_SITES / wall,
        "seconds_by_stage_per_family": {k: float(np.mean([r.get(k, 0.0) for r in results])) for k in keys},
    }
    if cpu_baseline:
        from oracle.siterm_oracle import fit_sites

        # the same regularised, compacted count tensors the GPU fit consumed, rebuilt for family 0
        import torch

        from cherryml_b200.siterm._site_specific import estimate_site_specific_rate_matrices_given_tree_and_site_rates as stage
        from cherryml_b200.siterm import _site_specific as mod

        captured = {}
        real = mod.quantized_transitions_mle_vectorized_over_sites

        def spy(**kwargs):
            captured.update(kwargs)
            return real(**kwargs)

        mod.quantized_transitions_mle_vectorized_over_sites = spy
        try:
            r0 = results[0]
            step = 1.1 ** (64 / GRID_STEPS)
            gpu = stage(tree=r0["learnt_tree"], site_rates=r0["learnt_site_rates"], msa=msas[0],
                        alphabet=list(amino_acids), regularization_strength=0.5,
                        regularization_rate_matrix=lg.to_numpy(),
                        quantization_points=[0.03 * step ** i for i in range(-GRID_STEPS, GRID_STEPS + 1)],
                        optimization_num_epochs=NUM_EPOCHS, vectorized_cherryml_implementation_device=str(device))
        finally:
            mod.quantized_transitions_mle_vectorized_over_sites = real
        import os

        cores = usable_cores()
        t0 = time.perf_counter()
        cpu = fit_sites(captured["counts"], captured["times"], NUM_EPOCHS, captured["initialization"],
                        num_threads=cores)
        cpu_s = time.perf_counter() - t0
        torch.set_num_threads(1)
        out["cpu_baseline"] = {
            "value": cpu_s, "unit": "s per family (batched fit only)", "cores": cores, "kind": "port",
            "sample": f"the {N_SITES} per-site problems of one family, torch-CPU restatement of the reference's "
                      "quantized_transitions_mle_vectorized_over_sites (oracle/siterm_oracle.py)",
            "gpu_seconds_same_stage": float(gpu.get("time_compute_loss", 0.0)),
        }
        out["matches_oracle_1e-6"] = bool(np.max(np.abs(cpu["res"] - gpu["res"])) < 1e-6 * np.max(np.abs(cpu["res"])))
    return out


def bench_siterm_sharded(device, process_group, sites: int = 16 * N_SITES, buckets: int = 4, seed: int = 5) -> Dict:
    """SURVEY 8e, third stage row: the batched per-site fit with the sites sharded over the ranks of
    ``process_group`` (contiguous blocks, no exchange during training, one gather of the results).  A fixed
    total of ``sites`` synthetic per-site problems (LG-like counts in ``buckets`` time buckets, strong scaling):
    every rank times the same call; returns the maximum over the ranks and, on every rank, the full result's
    checksum so that the gather is exercised."""
    import torch
    import torch.distributed as dist

    from cherryml_b200.siterm._vectorized import quantized_transitions_mle_vectorized_over_sites

    rng = np.random.default_rng(seed)
    lg = read_rate_matrix(get_lg_path()).to_numpy()
    counts = rng.poisson(3.0, size=(sites, buckets, 20, 20)).astype(np.float64)
    counts = counts + counts.transpose(0, 1, 3, 2) + 20.0 * np.eye(20)[None, None]
    times = np.tile(0.05 * 2.5 ** np.arange(buckets), (sites, 1))
    init = np.tile(lg[None], (sites, 1, 1))
    kw = dict(counts=counts, times=times, num_epochs=NUM_EPOCHS, initialization=init, device=str(device))
    quantized_transitions_mle_vectorized_over_sites(**dict(kw, num_epochs=4), process_group=process_group)  # warm-up
    if process_group is not None:
        dist.barrier(group=process_group)
    torch.cuda.synchronize(device)
    t0 = time.perf_counter()
    res = quantized_transitions_mle_vectorized_over_sites(**kw, process_group=process_group)
    torch.cuda.synchronize(device)
    secs = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=device)
    world = 1
    if process_group is not None:
        dist.all_reduce(secs, op=dist.ReduceOp.MAX, group=process_group)
        world = dist.get_world_size(process_group)
    return {"sites_total": sites, "buckets": buckets, "num_epochs": NUM_EPOCHS, "n_gpus": world,
            "seconds": float(secs[0]), "sites_per_s": sites / float(secs[0]),
            "result_checksum": float(np.abs(res["res"]).sum()),
            "sharding": f"contiguous blocks of sites over {world} ranks, no exchange during training, one gather"}
