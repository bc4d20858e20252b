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
