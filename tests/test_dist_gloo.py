"""world_size-2 test of the multi-rank counting logic on CPU (gloo): the reference's family
striping over ranks plus ONE all-reduce of the raw integer histogram reproduces the
single-process result bit for bit.  The per-rank kernel is replaced by its numpy emulation
(tests/_emulate.py) -- this test is about the sharding and the reduction, not the kernel."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cherryml_b200.counting._ingest import build_co_batch, build_lg_batch
from cherryml_b200.utils import amino_acids, get_process_args
from oracle.native import count_batch_oracle
from tests._emulate import emulate_count, emulate_symmetrize
from tests.conftest import GOLDEN
from tests.test_oracle_counting import GRID_CO, GRID_LG, MEDIUM3

M3 = os.path.join(GOLDEN, "counting", "medium3")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, kind, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fams = get_process_args(rank, world, MEDIUM3)  # the reference's MPI striping
    if kind == "lg":
        batch = build_lg_batch(f"{M3}/tree_dir", f"{M3}/msa_dir", f"{M3}/site_rates_dir", fams, amino_acids,
                               "cherry++", True)
        raw = emulate_count(batch, sorted(GRID_LG), 20)
    else:
        batch = build_co_batch(f"{M3}/tree_dir", f"{M3}/msa_dir", f"{M3}/contact_map_dir", fams, amino_acids,
                               "cherry++", 7, True)
        raw = emulate_count(batch, sorted(GRID_CO), 20)
    t = torch.from_numpy(raw)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    dist.barrier()
    if rank == 0:
        np.save(out_path, emulate_symmetrize(t.numpy(), kind, 20, False))
    dist.destroy_process_group()


@pytest.mark.parametrize("kind", ["lg", "co"])
def test_two_rank_counting_equals_single_process(tmp_path, kind):
    out = str(tmp_path / "counts.npy")
    mp.spawn(_worker, args=(2, _free_port(), kind, out), nprocs=2, join=True)
    got = np.load(out)
    if kind == "lg":
        batch = build_lg_batch(f"{M3}/tree_dir", f"{M3}/msa_dir", f"{M3}/site_rates_dir", MEDIUM3, amino_acids,
                               "cherry++", True)
        exp = count_batch_oracle(batch, GRID_LG, 20, False)
    else:
        batch = build_co_batch(f"{M3}/tree_dir", f"{M3}/msa_dir", f"{M3}/contact_map_dir", MEDIUM3, amino_acids,
                               "cherry++", 7, True)
        exp = count_batch_oracle(batch, GRID_CO, 20, False)
    assert np.array_equal(got, exp)


def test_striping_is_a_partition():
    fams = [f"f{i}" for i in range(11)]
    for world in (1, 2, 3, 8, 16):
        parts = [get_process_args(r, world, fams) for r in range(world)]
        assert sorted(sum(parts, [])) == sorted(fams)
        assert parts[0] == fams[0::world]


# ---------------------------------------------------------------- fit: bucket sharding
def _fit_worker(rank, world, port, out_path):
    """Each rank evaluates loss and dL/dQ on ITS buckets (fit oracle, CPU), one all-reduce of
    [dL/dQ | loss] gives every rank the totals -- the exchange step of the sharded fit."""
    from cherryml_b200.estimation._engine import assign_buckets
    from cherryml_b200.io import read_count_matrices_array, read_rate_matrix
    from oracle.fit_oracle import loss_and_grad_oracle
    from tests.test_oracle_fit import INP, LG_COUNTS

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    q, _, counts = read_count_matrices_array(LG_COUNTS)
    Q = read_rate_matrix(os.path.join(INP, "lg.txt")).to_numpy()
    mine = assign_buckets(q, world, rate_scale=float(np.max(-np.diag(Q))))[rank]
    loss, grad = loss_and_grad_oracle(Q, np.asarray(q)[mine], counts[mine])
    w = counts[mine].sum()  # the oracle normalises by its own counts: undo, the kernels exchange raw sums
    packed = torch.from_numpy(np.concatenate([grad.reshape(-1) * w, [loss * w]]))
    dist.all_reduce(packed, op=dist.ReduceOp.SUM)
    dist.barrier()
    if rank == 0:
        np.save(out_path, packed.numpy())
    dist.destroy_process_group()


def test_two_rank_bucket_sharded_loss_and_gradient(tmp_path):
    from cherryml_b200.io import read_count_matrices_array, read_rate_matrix
    from oracle.fit_oracle import loss_and_grad_oracle
    from tests.test_oracle_fit import INP, LG_COUNTS

    out = str(tmp_path / "packed.npy")
    mp.spawn(_fit_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    q, _, counts = read_count_matrices_array(LG_COUNTS)
    got = np.load(out) / counts.sum()
    Q = read_rate_matrix(os.path.join(INP, "lg.txt")).to_numpy()
    loss, grad = loss_and_grad_oracle(Q, q, counts)
    assert np.allclose(got[-1], loss, rtol=1e-12)
    assert np.allclose(got[:-1].reshape(20, 20), grad, rtol=1e-10, atol=1e-10 * np.abs(grad).max())


def test_assign_buckets_is_a_balanced_partition():
    from cherryml_b200.estimation._engine import assign_buckets
    from cherryml_b200.synthetic import quantization_grid

    grid = quantization_grid()
    for world in (1, 2, 3, 8, 16):
        parts = assign_buckets(grid, world, rate_scale=1.1)
        assert sorted(np.concatenate(parts).tolist()) == list(range(len(grid)))
        cost = 1.0 + np.maximum(0.0, np.ceil(np.log2(np.asarray(grid) * 1.1 / 1.09)))
        loads = [cost[p].sum() for p in parts]
        assert max(loads) - min(loads) <= cost.max()
        again = assign_buckets(grid, world, rate_scale=1.1)  # deterministic: every rank gets the same answer
        assert all(np.array_equal(a, b) for a, b in zip(parts, again))


def _fc_worker(rank, world, port, msa_dir, out_root, families, cache_dir=None):
    """Two ranks run the sharded FastCherries stage with the per-rank GPU work replaced by its
    oracle (this test is about the striping, the files and the barrier, not the kernels)."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cherryml_b200.phylogeny_estimation import _fast_cherries as fc

    seen = []

    def fake_local(msa_dir, fams, rate_matrix_path, R, max_iters, tree_dir, rates_dir, ll_dir, *rest):
        seen.extend(fams)
        for f in fams:
            for d in (tree_dir, rates_dir, ll_dir):
                with open(os.path.join(d, f + ".txt"), "w") as fh:
                    fh.write(f"rank {rank}\n")

    fc._fast_cherries_local = fake_local
    dirs = [os.path.join(out_root, k) for k in ("tree", "rates", "ll")]
    if cache_dir is not None:  # the caching wrapper is active, and rank 1 enters it late
        import time

        from cherryml_b200 import caching

        caching.set_cache_dir(cache_dir)
        if rank == 1:
            time.sleep(1.0)
    fc.fast_cherries(msa_dir=msa_dir, families=families, rate_matrix_path="unused", num_rate_categories=4,
                     max_iters=50, num_processes=1, output_tree_dir=dirs[0], output_site_rates_dir=dirs[1],
                     output_likelihood_dir=dirs[2], process_group=dist.group.WORLD)
    # after the call returns on ANY rank, every family's files exist (the barrier)
    ok = all(os.path.exists(os.path.join(d, f + ".txt")) for d in dirs for f in families)
    with open(os.path.join(out_root, f"seen_{rank}.txt"), "w") as fh:
        fh.write(("ok" if ok else "missing") + "\n" + " ".join(seen))
    dist.destroy_process_group()


def test_two_rank_fast_cherries_stage_stripes_families(tmp_path):
    families = [f"fam{i}" for i in range(7)]
    mp.spawn(_fc_worker, args=(2, _free_port(), str(tmp_path), str(tmp_path), families), nprocs=2, join=True)
    seen = []
    for rank in range(2):
        status, fams = open(tmp_path / f"seen_{rank}.txt").read().split("\n")
        assert status == "ok"
        assert fams.split() == sorted(families)[rank::2]
        seen += fams.split()
    assert sorted(seen) == sorted(families)
    for f in families:
        owner = sorted(families).index(f) % 2
        assert open(tmp_path / "tree" / f"{f}.txt").read() == f"rank {owner}\n"


def test_two_rank_fast_cherries_stage_through_the_cache(tmp_path):
    """The same stage with a cache directory: every rank runs the caching wrapper; rank 0 alone cleans,
    verifies and writes the success tokens (a late rank used to delete the files of a faster one)."""
    families = [f"fam{i}" for i in range(5)]
    out = tmp_path / "out"
    out.mkdir()
    for k in ("tree", "rates", "ll"):
        (out / k).mkdir()
    mp.spawn(_fc_worker, args=(2, _free_port(), str(tmp_path), str(out), families, str(tmp_path / "cache")),
             nprocs=2, join=True)
    for rank in range(2):
        status, fams = open(out / f"seen_{rank}.txt").read().split("\n")
        assert status == "ok" and fams.split() == sorted(families)[rank::2]
    for k in ("tree", "rates", "ll"):
        for f in families:
            assert (out / k / f"{f}.txt").exists() and (out / k / f"{f}.success").exists()


def _siterm_worker(rank, world, port, out_root):
    """Two ranks fit disjoint blocks of sites (the per-rank GPU fit is replaced by a deterministic
    stand-in: this test is about the blocks and the gather) and both end up with the full result."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cherryml_b200.siterm import _vectorized as v

    L, B, N = 7, 3, 4
    counts = np.arange(L * B * N * N, dtype=np.float64).reshape(L, B, N, N)
    times = np.ones((L, B))
    real = v.quantized_transitions_mle_vectorized_over_sites

    def fake(counts, times, num_epochs, initialization=None, num_cores=1, device="cpu", process_group=None):
        if process_group is not None:
            return real(counts, times, num_epochs, initialization, num_cores, device, process_group)
        n = counts.shape[0]
        return {"res": counts.sum(axis=1), "loss_per_epoch_per_site": np.tile(counts.sum(axis=(1, 2, 3)), (2, 1)),
                "loss_per_epoch": np.zeros(2), "time_compute_loss": float(n)}

    v.quantized_transitions_mle_vectorized_over_sites = fake
    out = v._fit_sharded_over_sites(counts, times, 2, None, 1, "cpu", dist.group.WORLD)
    np.savez(os.path.join(out_root, f"siterm_{rank}.npz"), res=out["res"], per_site=out["loss_per_epoch_per_site"],
             mine=np.array([out["time_compute_loss"]]))
    dist.destroy_process_group()


def test_two_rank_siterm_fit_shards_sites(tmp_path):
    from cherryml_b200.siterm._vectorized import site_blocks

    assert [list(b) for b in site_blocks(7, 2)] == [[0, 1, 2, 3], [4, 5, 6]]
    assert [len(b) for b in site_blocks(3, 8)] == [1, 1, 1, 0, 0, 0, 0, 0]
    mp.spawn(_siterm_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    counts = np.arange(7 * 3 * 4 * 4, dtype=np.float64).reshape(7, 3, 4, 4)
    for rank, n_mine in ((0, 4.0), (1, 3.0)):
        g = np.load(tmp_path / f"siterm_{rank}.npz")
        assert np.array_equal(g["res"], counts.sum(axis=1))
        assert np.array_equal(g["per_site"], np.tile(counts.sum(axis=(1, 2, 3)), (2, 1)))
        assert g["mine"][0] == n_mine


def _ll_worker(rank, world, port, root, families):
    """Two ranks run the sharded likelihood stage with the per-family device function replaced
    by a deterministic stand-in (this test is about the striping, the files and the barrier)."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cherryml_b200.evaluation import _likelihood as ev
    from cherryml_b200.markov_chain import get_lg_path
    from cherryml_b200.utils import amino_acids

    seen = []

    def fake_dp(tree, msa, contact_map, site_rates, **kw):
        seen.append(sorted(msa.keys())[0])
        lls = [-(rank + 1.0)] * len(site_rates)
        return sum(lls), lls

    ev.dp_likelihood_computation = fake_dp
    out = os.path.join(root, "ll")
    ev.compute_log_likelihoods(
        tree_dir=os.path.join(root, "tree"), msa_dir=os.path.join(root, "msa"),
        site_rates_dir=os.path.join(root, "rates"), contact_map_dir=None, families=families,
        amino_acids=amino_acids, pi_1_path=os.path.join(root, "pi.txt"), Q_1_path=get_lg_path(), reversible_1=True,
        device_1="cuda", pi_2_path=None, Q_2_path=None, reversible_2=None, device_2=None,
        output_likelihood_dir=out, num_processes=1, process_group=dist.group.WORLD)
    ok = all(os.path.exists(os.path.join(out, f + ".txt")) for f in families)
    with open(os.path.join(root, f"ll_seen_{rank}.txt"), "w") as fh:
        fh.write(("ok" if ok else "missing") + "\n" + " ".join(seen))
    dist.destroy_process_group()


def test_two_rank_likelihood_stage_stripes_families(tmp_path):
    import numpy as np

    from cherryml_b200 import io
    from cherryml_b200.markov_chain import compute_stationary_distribution, get_lg_path
    from cherryml_b200.utils import amino_acids

    families = [f"fam{i}" for i in range(5)]
    for f in families:
        tree = io.Tree()
        tree.add_nodes(["r", f + "_a", f + "_b"])
        tree.add_edge("r", f + "_a", 0.1)
        tree.add_edge("r", f + "_b", 0.2)
        io.write_tree(tree, str(tmp_path / "tree" / (f + ".txt")))
        io.write_msa({f + "_a": "ACD", f + "_b": "ACE"}, str(tmp_path / "msa" / (f + ".txt")))
        io.write_site_rates([1.0, 0.5, 2.0], str(tmp_path / "rates" / (f + ".txt")))
    Q = io.read_rate_matrix(get_lg_path()).to_numpy(dtype=np.float64)
    io.write_probability_distribution(compute_stationary_distribution(Q), amino_acids, str(tmp_path / "pi.txt"))
    mp.spawn(_ll_worker, args=(2, _free_port(), str(tmp_path), families), nprocs=2, join=True)
    for rank in range(2):
        status, fams = open(tmp_path / f"ll_seen_{rank}.txt").read().split("\n")
        assert status == "ok"
        assert fams.split() == [f + "_a" for f in families[rank::2]]
    for k, f in enumerate(families):
        ll, lls = io.read_log_likelihood(str(tmp_path / "ll" / (f + ".txt")))
        assert lls == [-(k % 2 + 1.0)] * 3 and ll == 3 * lls[0]
    assert os.path.exists(tmp_path / "ll" / "profiling.txt")
