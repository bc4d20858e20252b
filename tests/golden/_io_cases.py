"""The objects tests/golden/make_golden_io.py writes with the reference's cherryml.io and
tests/test_io_formats.py writes with cherryml_b200.io: name -> (writer name, arguments)."""
import numpy as np

AA = list("ACDEFGHIKLMNPQRSTVWY")


def objects():
    rng = np.random.default_rng(0)
    Q = rng.uniform(1e-4, 2, (20, 20))
    np.fill_diagonal(Q, 0)
    np.fill_diagonal(Q, -Q.sum(1))
    pi = rng.dirichlet(np.ones(20))
    cm = (rng.random((7, 7)) < 0.3).astype(int)
    cm = np.maximum(cm, cm.T)
    counts = [(0.03, rng.integers(0, 50, (20, 20)).astype(float) + rng.choice([0, 0.5, 0.25], (20, 20))),
              (1.234e-05, np.zeros((20, 20)))]
    return {
        "rate_matrix.txt": ("write_rate_matrix", (Q, AA)),
        "pi.txt": ("write_probability_distribution", (pi, AA)),
        "site_rates.txt": ("write_site_rates", ([1.0, 0.5, 1e-5, 2.123456789012345, 3],)),
        "contact_map.txt": ("write_contact_map", (cm,)),
        "msa.txt": ("write_msa", ({"b": "AC-", "a": "DEF", "c": "GHI"},)),
        "sites_subset.txt": ("write_sites_subset", ([3, 1, 2],)),
        "ll.txt": ("write_log_likelihood", ((-3.5, [-1.0, -2.5, 1e-07]),)),
        "transitions.txt": ("write_transitions", ([("A", "C", 0.1), ("DE", "FG", 1e-05)],)),
        "tll.txt": ("write_transitions_log_likelihood", ([-1.5, -2e-08],)),
        "tree.txt": ("write_tree", (None,)),  # built with the io module's own Tree class, see make_tree
        "count_matrices.txt": ("write_count_matrices", (counts,)),
    }


def make_tree(io):
    t = io.Tree()
    t.add_nodes(["r", "x", "a", "b", "c"])
    t.add_edges([("r", "x", 0.1), ("r", "c", 1e-05), ("x", "a", 0.25), ("x", "b", 3.0)])
    return t


def write(io, name, path):
    import pandas as pd

    fn, args = objects()[name]
    if name == "tree.txt":
        args = (make_tree(io),)
    if name == "count_matrices.txt":
        args = ([(q, pd.DataFrame(m, index=AA, columns=AA)) for q, m in args[0]],)
    getattr(io, fn)(*args, path)
