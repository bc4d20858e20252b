"""Regenerate tests/golden/io/*.txt by writing the objects of _io_cases.py with the UNMODIFIED
reference's cherryml.io (imported from /root/reference; build container only).

    python tests/golden/make_golden_io.py
"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests", "golden"))
OUT = os.path.join(REPO, "tests/golden/io")

if __name__ == "__main__":
    from _io_cases import objects, write
    from make_golden import import_reference

    import_reference()
    import cherryml.io as ref_io

    for name in objects():
        path = os.path.join(OUT, name)
        if os.path.exists(path):
            os.remove(path)
        write(ref_io, name, path)
        print(name, os.path.getsize(path), "bytes")
