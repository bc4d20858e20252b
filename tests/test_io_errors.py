"""Malformed input files: the readers raise what the reference's readers raise (messages compared
side by side with the unmodified reference in the build container; the expected prefixes below
are the reference's)."""
import pytest

from cherryml_b200 import io

CASES = {
    "msa_odd": ("read_msa", ">a\nAC\n>b\n", "should have an even number of lines"),
    "msa_no_marker": ("read_msa", "a\nAC\n>b\nAC\n", "at line 0 expected '>[seq_name]' but found a"),
    "msa_empty": ("read_msa", "", "should have an even number of lines"),
    "tree_header": ("read_tree", "3 node\na\nb\nc\n2 edges\na b 0.1\na c 0.2\n", "should start with '[num_nodes] nodes'"),
    "tree_few_nodes": ("read_tree", "3 nodes\na\nb\n2 edges\na b 0.1\na c 0.2\n", "should have line '[num_edges] edges' at positi"),
    "tree_bad_edge": ("read_tree", "3 nodes\na\nb\nc\n2 edges\na b 0.1\na c\n", "should have line '[u] [v] [length]' at position"),
    "tree_unknown_node": ("read_tree", "3 nodes\na\nb\nc\n2 edges\na b 0.1\na d 0.2\n", "a and d should be nodes in the tree"),
    "tree_two_parents": ("read_tree", "3 nodes\na\nb\nc\n2 edges\na c 0.1\nb c 0.2\n",
                         "Node c already has a parent (a), cannot also have parent b - graph is not a tree."),
    "rates_header": ("read_site_rates", "3 site\n1.0 2.0 3.0\n", "should start with line '[num_sites] sites'"),
    "rates_count": ("read_site_rates", "3 sites\n1.0 2.0\n", "was supposed to have 3 sites, but it has 2"),
    "contacts_header": ("read_contact_map", "2 site\n10\n01\n",
                        "Contact map file should start with line '[num_sites] sites', but started with: 2 site"),
    "contacts_rows": ("read_contact_map", "2 sites\n10\n", "should have 2 rows, but has 1"),
    "contacts_columns": ("read_contact_map", "2 sites\n101\n01\n", ""),  # the reference fails inside numpy here
    "subset_count": ("read_sites_subset", "3 sites\n1 2\n", "was supposed to have 3 sites, but it ha"),
    "ll_count": ("read_log_likelihood", "-3.0\n3 sites\n-1.0 -2.0\n", "should have 3.0 values in line 3"),
    "ll_header": ("read_log_likelihood", "-3.0\n2 site\n-1.0 -2.0\n", "should have second line '[num_sites] site"),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_malformed_file_is_rejected_like_the_reference(name, tmp_path):
    reader, text, message = CASES[name]
    path = tmp_path / (name + ".txt")
    path.write_text(text)
    with pytest.raises(Exception) as err:
        getattr(io, reader)(str(path))
    assert message in str(err.value)


def test_ragged_msa_is_accepted_like_the_reference(tmp_path):
    path = tmp_path / "m.txt"
    path.write_text(">a\nAC\n>b\nACD\n")
    assert io.read_msa(str(path)) == {"a": "AC", "b": "ACD"}
