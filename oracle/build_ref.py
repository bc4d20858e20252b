"""Build the oracle's native pieces (TEST INFRASTRUCTURE).

* ``oracle/_build/libcount_oracle.so``  -- the C restatement ``oracle/count_encoded.c``.
* ``oracle/_ref/count_transitions`` and ``oracle/_ref/count_co_transitions`` -- the UNMODIFIED
  reference counting programs, compiled from where they lie under /root/reference with a
  five-function single-rank ``mpi.h`` stand-in (``oracle/mpishim/mpi.h``).
* ``oracle/_ref/fast_cherries`` -- the UNMODIFIED reference FastCherries program
  (phylogeny_estimation/FastCherries/*.cpp + its matrix_exponential/ sources) with the flags
  of the reference's own Makefile (-std=c++11 -O3 -finline-functions -funroll-loops).  Built only when
  the reference checkout is present (the build container); the binaries are git-ignored and
  travel to the GPU box with the snapshot.  No reference source is copied into the repo.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
REF_COUNTING = "/root/reference/cherryml/counting"
REF_FC = "/root/reference/cherryml/phylogeny_estimation/FastCherries"
REF_DIR = os.path.join(HERE, "_ref")
C_LIB = os.path.join(HERE, "_build", "libcount_oracle.so")


def _newer(target: str, *sources: str) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def build_c_oracle() -> str:
    src = os.path.join(HERE, "count_encoded.c")
    os.makedirs(os.path.dirname(C_LIB), exist_ok=True)
    if not _newer(C_LIB, src):
        subprocess.run(["gcc", "-O2", "-shared", "-fPIC", "-o", C_LIB, src], check=True)
    return C_LIB


def build_reference_binaries() -> bool:
    """Returns True if the binaries exist afterwards."""
    names = {"count_transitions": "_count_transitions.cpp", "count_co_transitions": "_count_co_transitions.cpp"}
    if os.path.isdir(REF_COUNTING):
        os.makedirs(REF_DIR, exist_ok=True)
        for out, src in names.items():
            src_path, out_path = os.path.join(REF_COUNTING, src), os.path.join(REF_DIR, out)
            if not _newer(out_path, src_path):
                subprocess.run(
                    ["g++", "-std=c++11", "-O3", "-I", os.path.join(HERE, "mpishim"), "-o", out_path, src_path],
                    check=True,
                )
    fc_out = os.path.join(REF_DIR, "fast_cherries")
    if os.path.isdir(REF_FC):
        fc_src = [
            os.path.join(REF_FC, f)
            for f in (
                "fast_cherries.cpp", "io_helpers.cpp", "pairing_algorithms.cpp", "branch_length_estimation.cpp",
                "matrix_exponential/matrix_exponential.cpp", "matrix_exponential/r8lib.cpp",
                "matrix_exponential/c8lib.cpp",
            )
        ]
        if not _newer(fc_out, *fc_src):
            subprocess.run(
                ["g++", "-std=c++11", "-O3", "-finline-functions", "-funroll-loops", "-w", "-o", fc_out, *fc_src],
                check=True,
            )
    return all(os.path.exists(os.path.join(REF_DIR, out)) for out in list(names) + ["fast_cherries"])


if __name__ == "__main__":
    print(build_c_oracle())
    print("reference binaries:", build_reference_binaries())
