"""``sites_subset_dir`` of the LG pipeline: MSAs and site rates restricted to listed sites
(reference estimation_end_to_end/_cherry.py:41-147, io/_sites_subset.py); host logic, no GPU."""
import os

import pytest

from cherryml_b200 import caching
from cherryml_b200._public_api import _subset_data_to_sites_subset
from cherryml_b200.io import (read_msa, read_site_rates, read_sites_subset, write_msa, write_site_rates,
                              write_sites_subset)


def test_sites_subset_files_round_trip(tmp_path):
    p = str(tmp_path / "s" / "fam.txt")
    write_sites_subset([4, 0, 2], p)
    assert open(p).read() == "3 sites\n4 0 2"
    assert read_sites_subset(p) == [4, 0, 2]
    write_sites_subset([], p)
    assert read_sites_subset(p) == []
    (tmp_path / "bad.txt").write_text("2 sites\n1")
    with pytest.raises(Exception, match="supposed to have 2 sites"):
        read_sites_subset(str(tmp_path / "bad.txt"))
    (tmp_path / "bad2.txt").write_text("2 columns\n1 2")
    with pytest.raises(Exception, match="should start with line"):
        read_sites_subset(str(tmp_path / "bad2.txt"))


def test_subset_stage_restricts_msa_and_rates(tmp_path):
    for d in ("msa", "rates", "subset"):
        (tmp_path / d).mkdir()
    write_msa({"a": "ARNDC", "b": "QEGHI"}, str(tmp_path / "msa" / "f.txt"))
    write_site_rates([0.5, 1.0, 1.5, 2.0, 2.5], str(tmp_path / "rates" / "f.txt"))
    write_sites_subset([3, 1], str(tmp_path / "subset" / "f.txt"))
    caching.set_cache_dir(str(tmp_path / "cache"))
    try:
        out = _subset_data_to_sites_subset(sites_subset_dir=str(tmp_path / "subset"), msa_dir=str(tmp_path / "msa"),
                                           site_rates_dir=str(tmp_path / "rates"), families=["f"], num_processes=3)
    finally:
        caching.set_cache_dir(None)
    assert read_msa(os.path.join(out["output_msa_dir"], "f.txt")) == {"a": "DR", "b": "HE"}
    assert read_site_rates(os.path.join(out["output_site_rates_dir"], "f.txt")) == [2.0, 1.0]
    assert os.path.exists(os.path.join(out["output_msa_dir"], "f.success"))


def test_log_likelihood_and_cherries_files(tmp_path):
    from cherryml_b200.io import read_computed_cherries_from_file, read_log_likelihood, write_log_likelihood

    p = str(tmp_path / "ll" / "f.txt")
    write_log_likelihood((-12.5, [-4.25, -8.25]), p)
    assert open(p).read() == "-12.5\n2 sites\n-4.25 -8.25"
    assert read_log_likelihood(p) == (-12.5, [-4.25, -8.25])
    write_log_likelihood((0.0, None), p)
    assert read_log_likelihood(p) == (0.0, None)
    (tmp_path / "c.txt").write_text("a\nb\n0.12500000000000000\nc\nd\n2.00000000000000000\n")
    assert read_computed_cherries_from_file(str(tmp_path / "c.txt")) == ([("a", "b"), ("c", "d")], [0.125, 2.0])


def test_transition_files_and_msa_sizes(tmp_path):
    import pytest

    from cherryml_b200 import io

    msa = {"a": "AC-D", "b": "A_.D", "c": "ACDE"}
    p = str(tmp_path / "m" / "fam.txt")
    io.write_msa(msa, p)
    assert io.get_msa_num_sites(p) == 4
    assert io.get_msa_num_sequences(p) == 3
    assert io.get_msa_num_residues(p, exclude_gaps=False) == 12
    assert io.get_msa_num_residues(p, exclude_gaps=True) == 9

    tr = [("A", "C", 0.5), ("AC", "DE", 1e-05), ("S", "S", 3.0)]
    p = str(tmp_path / "t" / "tr.txt")
    io.write_transitions(tr, p)
    assert open(p).read() == "3 transitions\nA C 0.5\nAC DE 1e-05\nS S 3.0\n"
    assert io.read_transitions(p) == tr
    with open(p, "w") as f:
        f.write("2 transitions\nA C 0.5\n")
    with pytest.raises(ValueError, match="Expected 2 transitions"):
        io.read_transitions(p)
    with open(p, "w") as f:
        f.write("1 transition\nA C 0.5\n")
    with pytest.raises(ValueError, match="should start with"):
        io.read_transitions(p)

    lls = [-1.25, -0.1, -3.0e-07]
    p = str(tmp_path / "l" / "ll.txt")
    io.write_transitions_log_likelihood(lls, p)
    assert io.read_transitions_log_likelihood(p) == lls
    per_site = [[-1.0, -2.0], [-0.5]]
    p = str(tmp_path / "ps" / "ll.pkl")
    io.write_transitions_log_likelihood_per_site(per_site, p)
    assert io.read_transitions_log_likelihood_per_site(p) == per_site
    io.write_str("x y", str(tmp_path / "s.txt"))
    assert io.read_str(str(tmp_path / "s.txt")) == "x y"
