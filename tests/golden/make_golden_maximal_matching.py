"""Regenerate tests/golden/maximal_matching: seeded random contact maps (``in/``) and the output of
the UNMODIFIED reference stage ``cherryml.evaluation.create_maximal_matching_contact_map`` (networkx
maximal matching) at minimum distance 3 (``ref/``) and 7 (``ref7/``).  Build container only.

    python tests/golden/make_golden_maximal_matching.py
"""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests", "golden"))
OUT = os.path.join(REPO, "tests/golden/maximal_matching")

if __name__ == "__main__":
    from make_golden import import_reference

    import_reference()
    import cherryml.caching as ref_caching
    import cherryml.evaluation as ref_eval
    import cherryml.io as ref_io

    rng = np.random.default_rng(4)
    families = []
    for k, (L, p) in enumerate(((12, 0.3), (40, 0.15), (7, 0.9), (5, 0.0), (60, 0.08))):
        cm = (rng.random((L, L)) < p).astype(int)
        cm = np.maximum(cm, cm.T)
        np.fill_diagonal(cm, 1)
        os.makedirs(os.path.join(OUT, "in"), exist_ok=True)
        ref_io.write_contact_map(cm, os.path.join(OUT, "in", f"f{k}.txt"))
        families.append(f"f{k}")
    ref_caching.set_cache_dir(None)
    for sub, dist in (("ref", 3), ("ref7", 7)):
        ref_eval.create_maximal_matching_contact_map(
            i_contact_map_dir=os.path.join(OUT, "in"), families=families,
            minimum_distance_for_nontrivial_contact=dist, num_processes=1, o_contact_map_dir=os.path.join(OUT, sub))
    for root, _, files in os.walk(OUT):  # the stage leaves its outputs read-only and adds tokens
        for f in files:
            os.chmod(os.path.join(root, f), 0o644)
            if not f.endswith(".txt") or f == "result.txt":
                os.remove(os.path.join(root, f))
