"""Build the CUDA library in-tree: ``python -m cherryml_b200.csrc.build``.

Produces ``cherryml_b200/csrc/libcherryml_b200.so`` for sm_100a with ``-lineinfo`` so that
ncu source pages map to the .cu files.  nvcc cross-compiles without a GPU.
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_NAME = "libcherryml_b200.so"
LIB_PATH = os.path.join(HERE, LIB_NAME)
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = [
    "-O3",
    "-std=c++17",
    "-lineinfo",
    "-Xcompiler",
    "-fPIC",
    "-Xcompiler",
    "-O3",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def sources():
    return sorted(f for f in os.listdir(HERE) if f.endswith(".cu"))


def _digest() -> str:
    h = hashlib.sha256()
    names = sources() + sorted(f for f in os.listdir(HERE) if f.endswith(".cuh"))
    names.append(os.path.join("..", "..", "include", "cherryml_b200.h"))
    for name in names:
        with open(os.path.join(HERE, name), "rb") as f:
            h.update(name.encode())
            h.update(f.read())
    h.update(" ".join(ARCH_FLAGS + NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    stamp = os.path.join(HERE, "build", "stamp.txt")
    digest = _digest()
    if (
        not force
        and os.path.exists(LIB_PATH)
        and os.path.exists(stamp)
        and open(stamp).read().strip() == digest
    ):
        return LIB_PATH
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    nvcc = _nvcc()
    extra = ["-Xptxas", "-v"] if verbose else []

    def compile_one(src: str) -> str:
        obj = os.path.join(HERE, "build", src[:-3] + ".o")
        cmd = [nvcc, *ARCH_FLAGS, *NVCC_FLAGS, *extra, "-c", os.path.join(HERE, src), "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{res.stdout}\n{res.stderr}")
        if verbose:
            sys.stderr.write(res.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(compile_one, sources()))
    cmd = [nvcc, *ARCH_FLAGS, "-shared", "-o", LIB_PATH, *objs, "-lcudart_static", "-lpthread", "-ldl", "-lrt"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    with open(stamp, "w") as f:
        f.write(digest)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
