"""FastCherries on the GPU (cherry_fc_pair / cherry_fc_ble through the C ABI) against
(1) outputs of the UNMODIFIED reference program on the golden cases -- cherries, '%.17f'
distances and site rates must be identical text -- and (2) the oracle on seeded batches of
several families per launch, fed the SAME device-computed table: pairs, length indices, rate
categories and iteration counts bit-exact."""
import math
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from cherryml_b200 import _lib
from cherryml_b200.io import read_site_rates, read_tree
from cherryml_b200.phylogeny_estimation import _fast_cherries as fc
from tests._fc_cases import expected_outputs, load_cases, msa_text, parse_rate_matrix

CASES = load_cases()
AA = "ARNDCQEGHILKMFPSTWYV"


def _run_files(tmp_path, texts, alphabet, Q, R, max_iters, seed, num_steps):
    paths = []
    for i, t in enumerate(texts):
        p = tmp_path / f"fam{i}.txt"
        p.write_text(t)
        paths.append(str(p))
    grid = fc.quantization_grid(0.03, 1.1, num_steps)
    cats = fc.ble_rate_categories(R)
    weights = fc.initial_rate_weights(cats)
    priors = np.array([2 * math.log(r) - 3 * r for r in cats])
    names, msa, fams = fc.encode_families(paths, alphabet)
    table = fc.log_transition_table(Q, grid, cats, "cuda:0")
    before = _lib.launch_count()
    out = fc.fast_cherries_device(msa, fams, len(alphabet), table, priors, weights, seed, max_iters, "cuda:0")
    assert _lib.launch_count() - before == 3  # pairing, pair tables, coordinate ascent
    return names, msa, fams, table, grid, cats, weights, out


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_matches_reference_program(case, tmp_path):
    alphabet, Q = parse_rate_matrix(case["rate_matrix_text"])
    names, msa, fams, table, grid, cats, weights, out = _run_files(
        tmp_path, [msa_text(case)], alphabet, Q, case["num_rate_categories"], case["max_iters"], case["seed"],
        case["num_steps"])
    exp_cherries, exp_dist, exp_rates = expected_outputs(case)
    got = [(names[0][a], names[0][b]) for a, b in zip(out["pair_a"], out["pair_b"])]
    assert got == exp_cherries
    lengths, rates = fc.normalise_lengths_and_rates(out["len_idx"], out["site_cat"], grid, cats)
    assert ["%.17f" % x for x in lengths] == exp_dist
    assert ["%.17f" % x for x in rates] == exp_rates
    n = len(names[0])
    assert (out["unpaired"][0] >= 0) == (n % 2 == 1)


def _random_msa(rng, n, L, gap, mut):
    root = rng.integers(0, 20, L)
    rows = []
    for _ in range(n):
        r = rows[rng.integers(0, len(rows))].copy() if rows and rng.random() < 0.7 else root.copy()
        flip = rng.random(L) < mut
        r[flip] = rng.integers(0, 20, int(flip.sum()))
        rows.append(r)
    out = []
    for i, r in enumerate(rows):
        s = np.array(list(AA))[r]
        s[rng.random(L) < gap] = "-"
        out.append("".join(s))
    return "".join(f">s{i}\n{s}\n" for i, s in enumerate(out))


@pytest.mark.parametrize("R,max_iters", [(1, 50), (4, 50), (20, 50), (20, 2)])
def test_batch_matches_oracle_on_same_table(tmp_path, R, max_iters):
    from oracle import fast_cherries_oracle as fo

    rng = np.random.default_rng(100 + R)
    shapes = [(2, 16), (3, 15), (9, 33), (40, 100), (65, 64), (128, 301), (257, 48), (300, 17), (31, 500), (1, 20),
              (512, 90)]
    texts = [_random_msa(rng, n, L, gap=float(rng.uniform(0, 0.4)), mut=float(rng.uniform(0.02, 0.5)))
             for n, L in shapes]
    alphabet, Q = parse_rate_matrix(CASES[0]["rate_matrix_text"])
    names, msa, fams, table, grid, cats, weights, out = _run_files(tmp_path, texts, alphabet, Q, R, max_iters, 99, 64)
    sym = table.cpu().numpy()
    for f, (n, L) in enumerate(shapes):
        fam = fams[f]
        rows = msa[int(fam["msa_off"]): int(fam["msa_off"]) + n * int(fam["row_stride"])].reshape(n, -1)[:, :L]
        seqs = rows.astype(np.int64)
        seqs[seqs == 20] = -1
        c0, c1 = int(fam["cherry_off"]), int(fam["cherry_off"]) + n // 2
        s0, s1 = int(fam["site_off"]), int(fam["site_off"]) + L
        cherries = fo.divide_and_pair(seqs, 99)
        assert list(zip(out["pair_a"][c0:c1].tolist(), out["pair_b"][c0:c1].tolist())) == cherries, (n, L)
        paired = {v for c in cherries for v in c}
        left = [i for i in range(n) if i not in paired]
        assert int(out["unpaired"][f]) == (left[0] if left else -1)
        if n < 2:
            continue
        # the oracle's ble() adds T + T^T itself: hand it half of the (exactly symmetric) device table
        len_idx, site_cat, iters = _ble_on_sym(fo, seqs, cherries, sym, cats, weights, max_iters)
        assert np.array_equal(out["len_idx"][c0:c1], len_idx), (n, L)
        assert np.array_equal(out["site_cat"][s0:s1], site_cat), (n, L)
        assert int(out["iters"][f]) == iters


def _ble_on_sym(fo, seqs, cherries, sym, cats, weights, max_iters):
    a = np.array([c[0] for c in cherries])
    b = np.array([c[1] for c in cherries])
    xa, xb = seqs[a], seqs[b]
    site_cat = fo.initial_site_categories(seqs, weights, sym.shape[2])
    len_idx = fo.branch_length_indices(xa, xb, sym, site_cat)
    priors = np.array([2 * math.log(r) - 3 * r for r in cats])
    match, iters = False, 0
    while not match and max_iters:
        site_cat = fo.site_rate_indices(xa, xb, sym, len_idx, priors)
        new_len = fo.branch_length_indices(xa, xb, sym, site_cat)
        match = bool(np.array_equal(new_len, len_idx))
        len_idx = new_len
        max_iters -= 1
        iters += 1
    return len_idx, site_cat, iters


def test_device_table_close_to_scipy():
    from oracle import fast_cherries_oracle as fo

    alphabet, Q = parse_rate_matrix(CASES[0]["rate_matrix_text"])
    grid = fc.quantization_grid(0.03, 1.1, 64)
    cats = fc.ble_rate_categories(4)
    dev = fc.log_transition_table(Q, grid, cats, "cuda:0").cpu().numpy()
    T = fo.log_table_scipy(Q, grid, cats)
    ref = T + np.swapaxes(T, 2, 3)
    assert np.array_equal(dev, np.swapaxes(dev, 2, 3))  # exactly symmetric
    assert np.max(np.abs(dev - ref) / np.abs(ref)) < 1e-10


def test_stage_function_outputs_and_cache(tmp_path):
    from cherryml_b200 import caching
    from cherryml_b200.phylogeny_estimation import fast_cherries

    case = next(c for c in CASES if c["name"] == "synthetic_n33_L100_R4")
    msa_dir = tmp_path / "msas"
    msa_dir.mkdir()
    (msa_dir / "famA.txt").write_text(case["msa_text"])
    (msa_dir / "famB.txt").write_text(next(c for c in CASES if c["name"] == "synthetic_n16_L48_R4")["msa_text"])
    qp = tmp_path / "Q.txt"
    qp.write_text(case["rate_matrix_text"])
    caching.set_cache_dir(str(tmp_path / "cache"))
    try:
        kw = dict(msa_dir=str(msa_dir), families=["famB", "famA"], rate_matrix_path=str(qp), num_rate_categories=4,
                  max_iters=50, num_processes=3, verbose=False)
        res = fast_cherries(**kw)
        before = _lib.launch_count()
        res2 = fast_cherries(**kw)  # cached: nothing runs
        assert res2 == res and _lib.launch_count() == before
    finally:
        caching.set_cache_dir(None)
    exp_cherries, exp_dist, _ = expected_outputs(case)
    assert open(os.path.join(res["output_site_rates_dir"], "famA.txt")).read() == case["site_rates_file"]
    assert len(read_site_rates(os.path.join(res["output_site_rates_dir"], "famA.txt"))) == 100
    tree = read_tree(os.path.join(res["output_tree_dir"], "famA.txt"))
    inner = [v for v, _ in tree.children("root")]
    assert inner[:16] == [f"internal-{i}" for i in range(16)] and len(inner) == 17  # 33 leaves: one left over
    for i, (pair, d) in enumerate(zip(exp_cherries, exp_dist)):
        kids = tree.children(f"internal-{i}")
        assert tuple(k for k, _ in kids) == pair
        assert all(length == float(d) / 2.0 for _, length in kids)
    assert open(os.path.join(res["output_likelihood_dir"], "famA.txt")).read() == "0.0"
    for d in res.values():
        assert os.path.exists(os.path.join(d, "famA.success")) and os.path.exists(os.path.join(d, "famB.success"))
    prof = open(os.path.join(res["output_tree_dir"], "famA.profiling")).read().split("\n")
    assert [ln.split()[0] for ln in prof] == ["pairing_time:", "ble_time:", "cpp_time:", "total_time:"]


def test_tree_free_lg_pipeline_counts_match_reference_trees(tmp_path):
    """FastCherries -> counting -> fit through the public pipeline.  The count matrices must equal
    those counted on trees / site rates written from the REFERENCE program's golden outputs."""
    import tarfile
    from functools import partial

    from cherryml_b200 import caching
    from cherryml_b200._public_api import _quantization_points, lg_end_to_end_with_cherryml_optimizer
    from cherryml_b200.counting import count_transitions
    from cherryml_b200.io import Tree, read_rate_matrix, write_tree
    from cherryml_b200.markov_chain import get_lg_path
    from cherryml_b200.phylogeny_estimation import fast_cherries
    from cherryml_b200.utils import get_amino_acids
    from tests.conftest import GOLDEN

    fams = ["13gs_1_A", "1a0b_1_A", "1a2t_1_A"]
    with tarfile.open(os.path.join(GOLDEN, "demo_data.tar.xz")) as tf:
        tf.extractall(tmp_path, members=[tf.getmember(f"msas/{f}.txt") for f in fams])
    msa_dir = str(tmp_path / "msas")
    tree_dir, rates_dir = tmp_path / "ref_trees", tmp_path / "ref_rates"
    tree_dir.mkdir()
    rates_dir.mkdir()
    for f in fams:
        case = next(c for c in CASES if c["name"] == f"demo_{f}_R20")
        cherries, dist, _ = expected_outputs(case)
        tree = Tree()
        tree.add_node("root")
        for i, ((a, b), d) in enumerate(zip(cherries, dist)):
            tree.add_node(f"internal-{i}")
            tree.add_edge("root", f"internal-{i}", 1.0)
            for leaf in (a, b):
                tree.add_node(leaf)
                tree.add_edge(f"internal-{i}", leaf, float(d) / 2.0)
        write_tree(tree, str(tree_dir / f"{f}.txt"))
        (rates_dir / f"{f}.txt").write_text(case["site_rates_file"])
    caching.set_cache_dir(str(tmp_path / "cache"))
    try:
        res = lg_end_to_end_with_cherryml_optimizer(
            msa_dir=msa_dir, families=fams,
            tree_estimator=partial(fast_cherries, max_iters=50, num_rate_categories=20, verbose=False),
            initial_tree_estimator_rate_matrix_path=get_lg_path(), num_iterations=2, num_epochs=40,
            use_cpp_counting_implementation=False,
        )
        ref_counts = count_transitions(
            tree_dir=str(tree_dir), msa_dir=msa_dir, site_rates_dir=str(rates_dir), families=fams,
            amino_acids=get_amino_acids(), quantization_points=_quantization_points(0.03, 1.1, 64),
            edge_or_cherry="cherry++", num_processes=1, use_cpp_implementation=False,
        )["output_count_matrices_dir"]
    finally:
        caching.set_cache_dir(None)
    ours = open(os.path.join(res["count_matrices_dir_0"], "result.txt")).read()
    assert ours == open(os.path.join(ref_counts, "result.txt")).read()
    assert "time_pairing" in res and res["time_ble"] > 0
    assert res["tree_estimator_output_dirs_1"]["output_tree_dir"] != res["tree_estimator_output_dirs_0"]["output_tree_dir"]
    Q = read_rate_matrix(res["learned_rate_matrix_path"]).to_numpy()
    assert Q.shape == (20, 20) and np.allclose(Q.sum(axis=1), 0, atol=1e-5) and (Q - np.diag(np.diag(Q)) >= 0).all()


def test_stage_with_process_group_shards_families(tmp_path):
    """A one-rank NCCL group drives the sharded code path (stripe, barrier); outputs equal the
    plain call's."""
    import torch
    import torch.distributed as dist

    from cherryml_b200.phylogeny_estimation import fast_cherries

    msa_dir = tmp_path / "msas"
    msa_dir.mkdir()
    fams = []
    for c in CASES:
        if c["name"] in ("synthetic_n16_L48_R4", "synthetic_n33_L100_R4", "synthetic_n7_L40_R4"):
            (msa_dir / f"{c['name']}.txt").write_text(c["msa_text"])
            fams.append(c["name"])
    qp = tmp_path / "Q.txt"
    qp.write_text(CASES[0]["rate_matrix_text"])
    created = not dist.is_initialized()
    if created:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    try:
        kw = dict(msa_dir=str(msa_dir), families=fams, rate_matrix_path=str(qp), num_rate_categories=4, max_iters=50,
                  num_processes=1, verbose=False)
        dirs = {}
        for tag, pg in (("plain", None), ("group", dist.group.WORLD)):
            dirs[tag] = {k: str(tmp_path / f"{tag}_{k}") for k in ("tree", "rates", "ll")}
            fast_cherries(output_tree_dir=dirs[tag]["tree"], output_site_rates_dir=dirs[tag]["rates"],
                          output_likelihood_dir=dirs[tag]["ll"], process_group=pg, **kw)
    finally:
        if created:
            dist.destroy_process_group()
    for f in fams:
        for k in ("tree", "rates", "ll"):
            plain = open(os.path.join(dirs["plain"][k], f + ".txt")).read()
            assert plain == open(os.path.join(dirs["group"][k], f + ".txt")).read()
        case = next(c for c in CASES if c["name"] == f)
        assert open(os.path.join(dirs["group"]["rates"], f + ".txt")).read() == case["site_rates_file"]


def test_tree_free_coevolution_pipeline(tmp_path):
    """cherryml_public_api(model_name="co-evolution", tree_estimator_name="FastCherries") from MSAs
    and contact maps alone: trees come from the GPU FastCherries stage; the co-transition counts
    equal those counted on the trees the stage wrote."""
    import tarfile

    from cherryml_b200 import caching, cherryml_public_api
    from cherryml_b200.counting import count_co_transitions
    from cherryml_b200._public_api import _quantization_points, create_maximal_matching_contact_map
    from cherryml_b200.io import read_rate_matrix
    from cherryml_b200.utils import get_amino_acids
    from tests.conftest import GOLDEN

    fams = ["13gs_1_A", "1a0b_1_A"]
    with tarfile.open(os.path.join(GOLDEN, "demo_data.tar.xz")) as tf:
        tf.extractall(tmp_path, members=[tf.getmember(f"{d}/{f}.txt") for d in ("msas", "contact_maps") for f in fams])
    cache, out = str(tmp_path / "cache"), str(tmp_path / "learned.txt")
    try:
        cherryml_public_api(output_path=out, model_name="co-evolution", msa_dir=str(tmp_path / "msas"),
                            contact_map_dir=str(tmp_path / "contact_maps"), cache_dir=cache, num_epochs=6,
                            num_rate_categories=1, families=fams, tree_estimator_name="FastCherries",
                            use_cpp_counting_implementation=False)
        (h,) = os.listdir(os.path.join(cache, "fast_cherries"))
        tree_dir = os.path.join(cache, "fast_cherries", h, "output_tree_dir")
        (h2,) = os.listdir(os.path.join(cache, "count_co_transitions"))
        ours = open(os.path.join(cache, "count_co_transitions", h2, "output_count_matrices_dir", "result.txt")).read()
        caching.set_cache_dir(str(tmp_path / "cache2"))
        cm = create_maximal_matching_contact_map(i_contact_map_dir=str(tmp_path / "contact_maps"), families=fams,
                                                 minimum_distance_for_nontrivial_contact=7, num_processes=1)
        again = count_co_transitions(
            tree_dir=tree_dir, msa_dir=str(tmp_path / "msas"), contact_map_dir=cm["o_contact_map_dir"], families=fams,
            amino_acids=get_amino_acids(), quantization_points=_quantization_points(0.03, 1.1, 64),
            edge_or_cherry="cherry++", minimum_distance_for_nontrivial_contact=7, num_processes=1,
            use_cpp_implementation=False)["output_count_matrices_dir"]
    finally:
        caching.set_cache_dir(None)
    assert ours == open(os.path.join(again, "result.txt")).read()
    Q = read_rate_matrix(out).to_numpy()
    assert Q.shape == (400, 400) and np.allclose(Q.sum(axis=1), 0, atol=1e-4)


@pytest.mark.parametrize("cpp_personality", [True, False])
def test_in_memory_fast_cherries_to_counts_equals_the_text_route(tmp_path, cpp_personality):
    """MSAs -> FastCherries -> counts with the residues resident on the device
    (fast_cherries_then_count_lg) against the reference's route: fast_cherries stage writes trees
    and site rates, count_transitions reads them back.  Counts identical."""
    import tarfile

    from cherryml_b200.counting import count_transitions
    from cherryml_b200._public_api import _quantization_points
    from cherryml_b200.io import read_count_matrices_array, read_rate_matrix
    from cherryml_b200.markov_chain import get_lg_path
    from cherryml_b200.phylogeny_estimation import fast_cherries
    from cherryml_b200.phylogeny_estimation._pipeline import fast_cherries_then_count_lg
    from tests.conftest import GOLDEN

    fams = ["13gs_1_A", "1a0b_1_A", "1a2t_1_A", "1a12_1_A"]
    with tarfile.open(os.path.join(GOLDEN, "demo_data.tar.xz")) as tf:
        tf.extractall(tmp_path, members=[tf.getmember(f"msas/{f}.txt") for f in fams])
    msa_dir = str(tmp_path / "msas")
    small = next(c for c in CASES if c["name"] == "synthetic_n33_L100_R20")   # odd family: one row left over
    (tmp_path / "msas" / "odd.txt").write_text(small["msa_text"])
    fams = sorted(fams + ["odd"])
    dirs = {k: str(tmp_path / k) for k in ("tree", "rates", "ll", "counts")}
    fast_cherries(msa_dir=msa_dir, families=fams, rate_matrix_path=get_lg_path(), num_rate_categories=20,
                  max_iters=50, num_processes=1, verbose=False, output_tree_dir=dirs["tree"],
                  output_site_rates_dir=dirs["rates"], output_likelihood_dir=dirs["ll"])
    qp = _quantization_points(0.03, 1.1, 64)
    from cherryml_b200.phylogeny_estimation import _fast_cherries as fc_stage

    # (1) the stage pair as the pipeline runs it: count_transitions takes the resident FastCherries results
    assert "entry" in fc_stage._HANDOFF
    count_transitions(tree_dir=dirs["tree"], msa_dir=msa_dir, site_rates_dir=dirs["rates"], families=fams,
                      amino_acids=list(AA), quantization_points=qp, edge_or_cherry="cherry++", num_processes=1,
                      use_cpp_implementation=cpp_personality, output_count_matrices_dir=str(tmp_path / "counts_mem"))
    assert "entry" not in fc_stage._HANDOFF  # taken
    # (2) the reference's route: everything read back from the files (no hand-off left)
    count_transitions(tree_dir=dirs["tree"], msa_dir=msa_dir, site_rates_dir=dirs["rates"], families=fams,
                      amino_acids=list(AA), quantization_points=qp, edge_or_cherry="cherry++", num_processes=1,
                      use_cpp_implementation=cpp_personality, output_count_matrices_dir=dirs["counts"])
    assert (open(os.path.join(dirs["counts"], "result.txt"), "rb").read()
            == open(tmp_path / "counts_mem" / "result.txt", "rb").read())
    _, _, expected = read_count_matrices_array(os.path.join(dirs["counts"], "result.txt"))
    Q = read_rate_matrix(get_lg_path()).to_numpy()
    with fc.NativeMsas([os.path.join(msa_dir, f + ".txt") for f in fams], list(AA)) as msas:
        counts, out = fast_cherries_then_count_lg(msas.msa, msas.fams, list(AA), Q, [float(x) for x in qp],
                                                  float32_branch_lengths=cpp_personality, device="cuda:0")
        got = counts.cpu().numpy()
    assert got.sum() > 1e5
    if cpp_personality:  # the C++ layout of result.txt keeps 6 significant digits: compare what it can show
        assert np.array_equal(np.array([[float("%g" % v) for v in row] for row in got.reshape(-1, 20)]).reshape(
            got.shape), expected)
    else:
        assert np.array_equal(got, expected)


def test_full_size_properties():
    """2048 Pfam-shaped families (1024 x 300) in one launch, the bench workload: every sequence is in
    exactly one cherry, indices are in range, a second run is identical (the kernels are
    deterministic), a family's result does not depend on the batch it is in, and a sample of
    families equals the oracle."""
    from oracle import fast_cherries_oracle as fo
    from cherryml_b200.io import read_rate_matrix
    from cherryml_b200.markov_chain import get_lg_path
    from cherryml_b200.synthetic import synthetic_fc

    F, N, L, R = 2048, 1024, 300, 20
    msa, fams = synthetic_fc(F, N, L, seed=3)
    Q = read_rate_matrix(get_lg_path()).to_numpy()
    grid = fc.quantization_grid(0.03, 1.1, 64)
    cats = fc.ble_rate_categories(R)
    weights = fc.initial_rate_weights(cats)
    priors = np.array([2 * math.log(r) - 3 * r for r in cats])
    table = fc.log_transition_table(Q, grid, cats, "cuda:0")
    run = lambda m, f: fc.fast_cherries_device(m, f, 20, table, priors, weights, 1234, 50, "cuda:0")  # noqa: E731
    out = run(msa, fams)
    pa, pb = out["pair_a"].reshape(F, N // 2), out["pair_b"].reshape(F, N // 2)
    both = np.sort(np.concatenate([pa, pb], axis=1), axis=1)
    assert np.array_equal(both, np.broadcast_to(np.arange(N), (F, N)))          # a perfect matching per family
    assert (out["unpaired"] == -1).all()
    assert out["len_idx"].min() >= 0 and out["len_idx"].max() < len(grid)
    assert out["site_cat"].min() >= 0 and out["site_cat"].max() < R
    assert out["iters"].min() >= 1 and out["iters"].max() <= 50
    again = run(msa, fams)
    for k in ("pair_a", "pair_b", "len_idx", "site_cat", "iters"):
        assert np.array_equal(out[k], again[k]), k
    # families 5, 700 and 2047 on their own
    stride = int(fams[0]["row_stride"])
    for f in (5, 700, 2047):
        one = fams[f: f + 1].copy()
        one["msa_off"], one["cherry_off"], one["site_off"], one["seq_off"] = 0, 0, 0, 0
        rows = msa[f * N * stride: (f + 1) * N * stride]
        alone = run(rows.copy(), one)
        c0, s0 = f * (N // 2), f * L
        assert np.array_equal(alone["pair_a"], out["pair_a"][c0: c0 + N // 2])
        assert np.array_equal(alone["pair_b"], out["pair_b"][c0: c0 + N // 2])
        assert np.array_equal(alone["len_idx"], out["len_idx"][c0: c0 + N // 2])
        assert np.array_equal(alone["site_cat"], out["site_cat"][s0: s0 + L])
    sym = table.cpu().numpy()
    for f in (0, 1234):
        seqs = msa[f * N * stride: (f + 1) * N * stride].reshape(N, stride)[:, :L].astype(np.int64)
        seqs[seqs == 20] = -1
        cherries = fo.divide_and_pair(seqs, 1234)
        c0, s0 = f * (N // 2), f * L
        assert list(zip(out["pair_a"][c0: c0 + N // 2].tolist(), out["pair_b"][c0: c0 + N // 2].tolist())) == cherries
        len_idx, site_cat, iters = _ble_on_sym(fo, seqs, cherries, sym, cats, weights, 50)
        assert np.array_equal(out["len_idx"][c0: c0 + N // 2], len_idx)
        assert np.array_equal(out["site_cat"][s0: s0 + L], site_cat)
        assert int(out["iters"][f]) == iters
