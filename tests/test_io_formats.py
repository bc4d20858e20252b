"""Every text format of cherryml.io: this package's writers produce the bytes the UNMODIFIED
reference's writers produce for the same objects (tests/golden/io, made by make_golden_io.py),
and its readers give the objects back."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from _io_cases import AA, make_tree, objects, write  # noqa: E402

from cherryml_b200 import io  # noqa: E402

GOLD = os.path.join(HERE, "golden", "io")


@pytest.mark.parametrize("name", sorted(objects()))
def test_writer_bytes_equal_the_reference(name, tmp_path):
    path = str(tmp_path / "sub" / name)
    write(io, name, path)
    assert open(path, "rb").read() == open(os.path.join(GOLD, name), "rb").read()


def test_readers_give_the_objects_back():
    o = {k: v[1] for k, v in objects().items()}
    q = io.read_rate_matrix(os.path.join(GOLD, "rate_matrix.txt"))
    assert list(q.index) == AA and list(q.columns) == AA and np.array_equal(q.to_numpy(), o["rate_matrix.txt"][0])
    pi = io.read_probability_distribution(os.path.join(GOLD, "pi.txt"))
    assert list(pi.index) == AA and np.array_equal(pi.to_numpy().reshape(-1), o["pi.txt"][0])
    assert io.read_site_rates(os.path.join(GOLD, "site_rates.txt")) == [float(x) for x in o["site_rates.txt"][0]]
    assert np.array_equal(io.read_contact_map(os.path.join(GOLD, "contact_map.txt")), o["contact_map.txt"][0])
    assert io.read_msa(os.path.join(GOLD, "msa.txt")) == o["msa.txt"][0]
    assert io.read_sites_subset(os.path.join(GOLD, "sites_subset.txt")) == o["sites_subset.txt"][0]
    assert io.read_log_likelihood(os.path.join(GOLD, "ll.txt")) == o["ll.txt"][0]
    assert io.read_transitions(os.path.join(GOLD, "transitions.txt")) == o["transitions.txt"][0]
    assert io.read_transitions_log_likelihood(os.path.join(GOLD, "tll.txt")) == o["tll.txt"][0]
    t, want = io.read_tree(os.path.join(GOLD, "tree.txt")), make_tree(io)
    assert t.nodes() == want.nodes() and t.edges() == want.edges()
    cms = io.read_count_matrices(os.path.join(GOLD, "count_matrices.txt"))
    for (qv, df), (qw, m) in zip(cms, o["count_matrices.txt"][0]):
        assert qv == qw and list(df.index) == AA and np.array_equal(df.to_numpy(), m)


def test_write_tree_scaling_and_prefix(tmp_path):
    """Reference io/_tree.py:193-211; expected text written by the reference for the same tree."""
    path = str(tmp_path / "t.txt")
    io.write_tree(make_tree(io), path, scaling_factor=0.3, node_name_prefix="fam-")
    assert open(path).read() == ("5 nodes\nfam-r\nfam-x\nfam-a\nfam-b\nfam-c\n4 edges\nfam-r fam-x 0.03\n"
                                 "fam-r fam-c 3e-06\nfam-x fam-a 0.075\nfam-x fam-b 0.8999999999999999\n")


def test_keyword_names_of_the_reference():
    import inspect

    assert list(inspect.signature(io.read_probability_distribution).parameters) == ["probability_distribution_path"]
    assert list(inspect.signature(io.write_probability_distribution).parameters) == [
        "probability_distribution", "states", "probability_distribution_path"]
    assert list(inspect.signature(io.write_tree).parameters) == ["tree", "tree_path", "scaling_factor", "node_name_prefix"]


def test_tree_scaled():
    t = make_tree(io).scaled(0.3, "p_")
    assert t.nodes() == ["p_r", "p_x", "p_a", "p_b", "p_c"]
    assert t.edges() == [("p_r", "p_x", 0.03), ("p_r", "p_c", 3e-06), ("p_x", "p_a", 0.075),
                         ("p_x", "p_b", 0.8999999999999999)]
    assert make_tree(io).scaled(1.0).edges() == make_tree(io).edges()
