"""The contact-map matching stage of the co-evolution pipeline against outputs of the UNMODIFIED
reference stage (networkx maximal matching) on seeded random contact maps
(tests/golden/maximal_matching, made by make_golden_maximal_matching.py)."""
import os

import numpy as np
import pytest

from cherryml_b200 import io
from cherryml_b200.evaluation import create_maximal_matching_contact_map

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "maximal_matching")
FAMILIES = [f"f{k}" for k in range(5)]


@pytest.mark.parametrize("sub,dist", [("ref", 3), ("ref7", 7)])
def test_outputs_equal_the_reference_stage(sub, dist, tmp_path):
    out = create_maximal_matching_contact_map(
        i_contact_map_dir=os.path.join(GOLD, "in"), families=FAMILIES, minimum_distance_for_nontrivial_contact=dist,
        num_processes=1, o_contact_map_dir=str(tmp_path / "out"))
    out_dir = out["o_contact_map_dir"] if isinstance(out, dict) else str(tmp_path / "out")
    for f in FAMILIES:
        got = open(os.path.join(out_dir, f + ".txt")).read()
        assert got == open(os.path.join(GOLD, sub, f + ".txt")).read()
        m = io.read_contact_map(os.path.join(out_dir, f + ".txt"))
        full = io.read_contact_map(os.path.join(GOLD, "in", f + ".txt"))
        assert (m.sum(axis=0) <= 1).all() and np.array_equal(m, m.T)  # a matching ...
        ii, jj = np.where(m == 1)
        assert all(full[i, j] == 1 and abs(i - j) >= dist for i, j in zip(ii, jj))  # ... of non-trivial contacts ...
        free = m.sum(axis=0) == 0  # ... that cannot be extended
        ii, jj = np.where(full == 1)
        assert not any(free[i] and free[j] and abs(i - j) >= dist for i, j in zip(ii, jj))
