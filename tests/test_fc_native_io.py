"""The FastCherries stage's native text I/O (cherry_fc_read_msas / cherry_fc_write_outputs, host
threads, no GPU) against the plain-Python implementations of the same formats, byte for byte."""
import numpy as np

from cherryml_b200.io import write_tree
from cherryml_b200.phylogeny_estimation import _fast_cherries as fc

AA = list("ARNDCQEGHILKMFPSTWYV")


def _write_msas(tmp_path, rng, shapes):
    paths = []
    for f, (n, L) in enumerate(shapes):
        rows = []
        for i in range(n):
            s = rng.choice(AA + ["-", "X", "a"], L)
            rows.append(f">fam{f}_s{i} extra\n{''.join(s)}\n")
        text = "".join(rows)
        if f % 3 == 1:
            text = text[:-1]          # no trailing newline
        if f % 3 == 2:
            text = "# comment\n" + text + ">dangling_name\n"   # ignored line, name without a sequence
        p = tmp_path / f"m{f}.txt"
        p.write_text(text)
        paths.append(str(p))
    return paths


def test_read_msas_matches_python_encoder(tmp_path):
    rng = np.random.default_rng(0)
    paths = _write_msas(tmp_path, rng, [(5, 33), (2, 16), (9, 1), (40, 100), (1, 7), (3, 15)])
    names, buf, fams = fc.encode_families(paths, AA)
    with fc.NativeMsas(paths, AA, n_threads=3, pinned=False) as m:
        assert np.array_equal(m.fams, fams)
        assert np.array_equal(m.msa, buf)
        for f in range(len(paths)):
            assert m.names(f) == names[f]


def test_write_outputs_matches_python_writers(tmp_path):
    rng = np.random.default_rng(1)
    shapes = [(7, 20), (2, 16), (12, 50), (33, 9)]
    paths = _write_msas(tmp_path, rng, shapes)
    grid = fc.quantization_grid(0.03, 1.1, 64)
    cats = fc.ble_rate_categories(20)
    with fc.NativeMsas(paths, AA, n_threads=2, pinned=False) as m:
        fams = m.fams.copy()
        n_ch = int((fams["n_seqs"] // 2).sum())
        pair_a = np.zeros(n_ch, dtype=np.int32)
        pair_b = np.zeros(n_ch, dtype=np.int32)
        unpaired = np.full(len(paths), -1, dtype=np.int32)
        for f, (n, _) in enumerate(shapes):
            perm = rng.permutation(n)
            c0 = int(fams[f]["cherry_off"])
            pair_a[c0: c0 + n // 2] = perm[0: 2 * (n // 2): 2]
            pair_b[c0: c0 + n // 2] = perm[1: 2 * (n // 2): 2]
            if n % 2:
                unpaired[f] = perm[-1]
        # grid indices over the whole range: exercises repr() in fixed and exponent notation
        len_idx = rng.integers(0, len(grid), n_ch).astype(np.int32)
        len_idx[:4] = [0, 1, len(grid) - 1, 64]
        site_cat = rng.integers(0, len(cats), int(fams["n_sites"].sum())).astype(np.int32)
        out = dict(pair_a=pair_a, pair_b=pair_b, unpaired=unpaired, len_idx=len_idx, site_cat=site_cat)
        prof = np.abs(rng.standard_normal((len(paths), 4))) * np.array([1e-7, 1e-3, 1.0, 1e17])
        prof[0] = [0.0, 1e-4, 1e-5, 123456789012345678.0]
        prof[1] = [1.0, 1e16, 1e15, 0.1]
        d = tmp_path / "out"
        d.mkdir()
        j = lambda ext: [str(d / f"f{f}{ext}") for f in range(len(paths))]  # noqa: E731
        m.write_outputs(out, grid, cats, j(".tree"), j(".newick"), j(".rates"), j(".ll"), j(".prof"), prof)
        for f, (n, L) in enumerate(shapes):
            c0, s0 = int(fams[f]["cherry_off"]), int(fams[f]["site_off"])
            names = m.names(f)
            pairs = list(zip(pair_a[c0: c0 + n // 2].tolist(), pair_b[c0: c0 + n // 2].tolist()))
            lengths, rates = fc.normalise_lengths_and_rates(len_idx[c0: c0 + n // 2], site_cat[s0: s0 + L], grid, cats)
            ref_tree = str(d / f"ref{f}.tree")
            write_tree(fc.cherries_tree(names, pairs, lengths, int(unpaired[f])), ref_tree)
            assert open(j(".tree")[f]).read() == open(ref_tree).read()
            assert open(j(".newick")[f]).read() == fc._newick(names, pairs, lengths, int(unpaired[f]))
            assert open(j(".rates")[f]).read() == f"{L} sites\n" + "".join(fc._fixed17(r) + " " for r in rates)
            assert open(j(".ll")[f]).read() == "0.0"
            p = prof[f]
            assert open(j(".prof")[f]).read() == (f"pairing_time: {p[0]}\nble_time: {p[1]}\ncpp_time: {p[2]}\n"
                                                   f"total_time: {p[3]}")


def test_count_matrix_files_native_equals_python(tmp_path):
    from cherryml_b200.io import (read_count_matrices_array, read_count_matrices_array_py,
                                  write_count_matrices_array, write_count_matrices_array_py)

    rng = np.random.default_rng(2)
    for S, states in ((3, ["A", "C", "G"]), (20, AA), (16, [a + b for a in "ACGT" for b in "ACGT"])):
        K = 5
        counts = rng.integers(0, 9, (K, S, S)) * 0.25 * (rng.random((K, S, S)) < 0.4)
        counts[0, 0, 0] = 1234567.25       # 6-significant-digit rounding in the C++ layout
        counts[1, 0, 1] = 1e-5 + 0.1
        counts[2, 1, 1] = 123456789012.5
        counts[3, 0, 2] = -0.0
        q = [6.729602379904665e-05, 0.0001, 0.03, 1.0, 13.373747053577777]
        for style in ("python", "cpp"):
            a, b = str(tmp_path / f"n_{S}_{style}.txt"), str(tmp_path / f"p_{S}_{style}.txt")
            write_count_matrices_array(q, states, counts, a, style)
            write_count_matrices_array_py(q, states, counts, b, style)
            assert open(a).read() == open(b).read()
            qn, sn, cn = read_count_matrices_array(a)
            qp, sp, cp = read_count_matrices_array_py(a)
            assert sn == sp == list(states)
            assert np.array_equal(qn, qp) and np.array_equal(cn, cp)
    with __import__("pytest").raises(Exception):
        read_count_matrices_array(str(tmp_path / "missing.txt"))


def test_rate_matrix_files_native_equals_python(tmp_path):
    from cherryml_b200.io import read_rate_matrix, write_rate_matrix, write_rate_matrix_py

    rng = np.random.default_rng(3)
    for dt in (np.float64, np.float32):
        for S, states in ((4, list("ACGT")), (20, AA)):
            m = (rng.standard_normal((S, S)) * np.exp(rng.uniform(-25, 25, (S, S)))).astype(dt)
            m[0, :4] = [0.0, -0.0, 1.0, -1.0]
            m[1, :4] = [1e-5, 1e-4, 1e15, 1e16]
            m[2, :3] = [np.inf, -np.inf, np.nan]
            m[3, :3] = [0.1, 123456.789, 5e-324 if dt is np.float64 else 1e-45]
            a, b = str(tmp_path / f"n{S}{dt.__name__}.txt"), str(tmp_path / f"p{S}{dt.__name__}.txt")
            write_rate_matrix(m, states, a)
            write_rate_matrix_py(m, states, b)
            assert open(a).read() == open(b).read()
    back = read_rate_matrix(a)
    assert list(back.columns) == AA


def test_in_memory_hand_off_equals_the_route_through_files(tmp_path):
    """FastCherries results -> counting batch WITHOUT the text files (count_layout +
    cherry_fc_lengths_and_rates; the device re-layout is emulated in numpy here) is array for array
    the batch cherry_ingest_lg builds from the files cherry_fc_write_outputs writes."""

    from cherryml_b200 import _lib
    from cherryml_b200.counting._ingest import build_lg_batch_native
    from cherryml_b200.phylogeny_estimation._pipeline import count_layout, count_layout_numpy

    rng = np.random.default_rng(5)
    shapes = [(9, 37), (2, 16), (40, 100), (13, 5), (64, 301)]
    msa_dir = tmp_path / "msas"
    msa_dir.mkdir()
    fam_names = [f"fam{i}" for i in range(len(shapes))]
    paths = []
    for name, (n, L) in zip(fam_names, shapes):
        text = "".join(f">{name}_s{i}\n{''.join(rng.choice(AA + ['-'], L))}\n" for i in range(n))
        (msa_dir / f"{name}.txt").write_text(text)
        paths.append(str(msa_dir / f"{name}.txt"))
    grid = fc.quantization_grid(0.03, 1.1, 64)
    cats = fc.ble_rate_categories(20)
    for f32 in (True, False):
        with fc.NativeMsas(paths, AA, n_threads=2, pinned=False) as m:
            fams = m.fams.copy()
            msa = m.msa.copy()
            n_ch = int((fams["n_seqs"] // 2).sum())
            pair_a = np.zeros(n_ch, dtype=np.int32)
            pair_b = np.zeros(n_ch, dtype=np.int32)
            unpaired = np.full(len(shapes), -1, dtype=np.int32)
            for f, (n, _) in enumerate(shapes):
                perm = rng.permutation(n)
                c0 = int(fams[f]["cherry_off"])
                pair_a[c0: c0 + n // 2] = perm[0: 2 * (n // 2): 2]
                pair_b[c0: c0 + n // 2] = perm[1: 2 * (n // 2): 2]
                if n % 2:
                    unpaired[f] = perm[-1]
            len_idx = rng.integers(0, len(grid), n_ch).astype(np.int32)
            site_cat = rng.integers(0, len(cats), int(fams["n_sites"].sum())).astype(np.int32)
            site_cat[int(fams[3]["site_off"]): int(fams[3]["site_off"]) + 5] = 7      # one category only
            out = dict(pair_a=pair_a, pair_b=pair_b, unpaired=unpaired, len_idx=len_idx, site_cat=site_cat)
            d = tmp_path / f"out{int(f32)}"
            for sub in ("trees", "rates"):
                (d / sub).mkdir(parents=True)
            j = lambda sub, ext: [str(d / sub / f"{nm}{ext}") for nm in fam_names]  # noqa: E731
            m.write_outputs(out, grid, cats, j("trees", ".txt"), [None] * 5, j("rates", ".txt"), [None] * 5,
                            [None] * 5, np.zeros((5, 4)))
        ref = build_lg_batch_native(str(d / "trees"), str(msa_dir), str(d / "rates"), fam_names, AA, "cherry++", f32,
                                    n_threads=2)
        lib = _lib.load()
        pair_t = np.zeros(n_ch)
        rate_table = np.zeros((len(shapes), len(cats)))
        _lib.check(lib.cherry_fc_lengths_and_rates(_lib.ptr(fams), len(shapes), _lib.ptr(len_idx), _lib.ptr(site_cat),
                                                   _lib.ptr(grid), len(grid), _lib.ptr(cats), len(cats), int(f32),
                                                   _lib.ptr(pair_t), _lib.ptr(rate_table), 2), "lengths_and_rates")
        lay = count_layout(fams, site_cat, rate_table, n_threads=2)
        ref_lay = count_layout_numpy(fams, site_cat, rate_table)
        for key in ("dest", "aux", "rate_vals", "fams", "tiles"):
            assert np.array_equal(lay[key], ref_lay[key]), key
        assert all(lay[k] == ref_lay[k] for k in ("r_pad", "msa_bytes", "examined"))
        assert np.array_equal(pair_t, ref.pair_t)
        assert np.array_equal(lay["rate_vals"], ref.rate_vals)
        assert np.array_equal(lay["aux"], ref.aux)
        assert np.array_equal(lay["fams"], ref.fams)
        assert np.array_equal(lay["tiles"], ref.tiles)
        assert lay["r_pad"] == ref.r_pad and lay["examined"] == ref.n_sites_examined
        local = np.concatenate([np.arange(n // 2) for n, _ in shapes])
        assert np.array_equal(ref.pair_a, 2 * local) and np.array_equal(ref.pair_b, 2 * local + 1)
        # the device re-layout, emulated
        msa_out = np.full(lay["msa_bytes"], 20, dtype=np.uint8)
        for f, (n, L) in enumerate(shapes):
            fin, fo = fams[f], lay["fams"][f]
            rows_in = msa[int(fin["msa_off"]): int(fin["msa_off"]) + n * int(fin["row_stride"])].reshape(n, -1)
            dest = lay["dest"][int(fin["site_off"]): int(fin["site_off"]) + L]
            c0 = int(fin["cherry_off"])
            for c in range(n // 2):
                for side, src in enumerate((pair_a[c0 + c], pair_b[c0 + c])):
                    base = int(fo["msa_off"]) + (2 * c + side) * int(fo["row_stride"])
                    msa_out[base + dest] = rows_in[src, :L]
        assert np.array_equal(msa_out, ref.msa[: len(msa_out)])


def test_empty_and_single_sequence_msas(tmp_path):
    """An empty MSA, a name without a sequence line and a single-sequence MSA (on which the reference
    program crashes) are read and written without cherries."""
    from cherryml_b200.io import read_site_rates, read_tree

    (tmp_path / "empty.txt").write_text("")
    (tmp_path / "one.txt").write_text(">only\nARND-\n")
    (tmp_path / "noseq.txt").write_text(">dangling\n")
    grid = fc.quantization_grid(0.03, 1.1, 64)
    cats = fc.ble_rate_categories(4)
    with fc.NativeMsas([str(tmp_path / f) for f in ("empty.txt", "one.txt", "noseq.txt")], AA, pinned=False) as m:
        assert m.fams["n_seqs"].tolist() == [0, 1, 0] and m.fams["n_sites"].tolist() == [0, 5, 0]
        assert m.names(1) == ["only"]
        out = dict(pair_a=np.zeros(0, np.int32), pair_b=np.zeros(0, np.int32),
                   unpaired=np.array([-1, 0, -1], np.int32), len_idx=np.zeros(0, np.int32),
                   site_cat=np.zeros(5, np.int32))
        j = lambda e: [str(tmp_path / f"o{i}{e}") for i in range(3)]  # noqa: E731
        m.write_outputs(out, grid, cats, j(".tree"), j(".nw"), j(".rates"), j(".ll"), j(".prof"), np.zeros((3, 4)))
    assert open(j(".tree")[0]).read() == "1 nodes\nroot\n0 edges\n"
    assert open(j(".rates")[0]).read() == "0 sites\n"
    tree = read_tree(j(".tree")[1])
    assert tree.nodes() == ["root", "only"] and tree.edges() == [("root", "only", 1.0)]
    assert read_site_rates(j(".rates")[1]) == [1.0] * 5
