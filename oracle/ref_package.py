"""The UNMODIFIED reference Python package, made importable for baselines (TEST / BENCH
INFRASTRUCTURE -- nothing under cherryml_b200/ imports this).

The GPU box has no /root/reference, and `bench.py --impl reference` / the `fit` section must
time the reference's own ``quantized_transitions_mle`` there (CPU arm with
all host threads and the reference's stock ``device="cuda"`` path on the same B200).  So, like
the reference C++ programs compiled into ``oracle/_ref``, the build step copies the
reference's ``cherryml/*.py`` tree into the git-ignored ``oracle/_ref/pkg`` (committed recipe,
no reference source enters the repository history) and the snapshot carries it to the box.

``import_reference()`` imports it with stand-ins for third-party modules that are absent in
this image (ete3, matplotlib, seaborn, biotite, wget, parameterized), a stub for its Cython
extension (only SiteRM uses it), a no-op pandas plotting backend (``ratelearner.py:167`` calls
``Series.plot``) and the pandas-3 spelling of ``delim_whitespace`` (``io/_rate_matrix.py``);
none of the hot-path arithmetic is touched.
"""
import os
import shutil
import sys
import types
from unittest import mock

HERE = os.path.dirname(os.path.abspath(__file__))
REF_CHECKOUT = "/root/reference"
PKG_DIR = os.path.join(HERE, "_ref", "pkg")


def build_reference_package() -> bool:
    """Copy cherryml/**/*.py from the reference checkout into oracle/_ref/pkg (build container
    only).  Returns True if the copy exists afterwards."""
    src = os.path.join(REF_CHECKOUT, "cherryml")
    dst = os.path.join(PKG_DIR, "cherryml")
    if os.path.isdir(src):
        stamp = os.path.join(PKG_DIR, ".stamp")
        newest = max(os.path.getmtime(os.path.join(r, f)) for r, _, fs in os.walk(src) for f in fs if f.endswith(".py"))
        if not (os.path.exists(stamp) and os.path.getmtime(stamp) >= newest):
            shutil.rmtree(dst, ignore_errors=True)
            for root, _, files in os.walk(src):
                for f in files:
                    if f.endswith(".py"):
                        rel = os.path.relpath(os.path.join(root, f), src)
                        out = os.path.join(dst, rel)
                        os.makedirs(os.path.dirname(out), exist_ok=True)
                        shutil.copyfile(os.path.join(root, f), out)
            with open(stamp, "w") as fh:
                fh.write("copied from /root/reference/cherryml by oracle/ref_package.py\n")
    return os.path.isdir(dst)


def reference_root():
    """Directory that holds the reference's ``cherryml`` package, or None."""
    if os.path.isdir(os.path.join(PKG_DIR, "cherryml")):
        return PKG_DIR
    if os.path.isdir(os.path.join(REF_CHECKOUT, "cherryml")):
        return REF_CHECKOUT
    return None


def import_reference():
    """Import the reference package (see the module docstring for the stand-ins)."""
    root = reference_root()
    if root is None:
        raise ImportError("the reference package is neither under oracle/_ref/pkg nor at /root/reference")
    for name in ("ete3", "matplotlib", "matplotlib.pyplot", "matplotlib.patches", "seaborn", "biotite",
                 "biotite.structure", "biotite.structure.io", "biotite.structure.io.pdb", "wget", "parameterized"):
        sys.modules.setdefault(name, mock.MagicMock())
    stub = types.ModuleType("cherryml._siterm.fast_site_rates")
    stub.compute_optimal_site_rates = None
    sys.modules.setdefault("cherryml._siterm.fast_site_rates", stub)
    import pandas as pd

    if not getattr(pd.read_csv, "_cherry_shim", False):
        _read_csv = pd.read_csv

        def read_csv(*args, **kwargs):
            if kwargs.pop("delim_whitespace", False):
                kwargs["sep"] = r"\s+"
            return _read_csv(*args, **kwargs)

        read_csv._cherry_shim = True
        pd.read_csv = read_csv
    backend = types.ModuleType("cherry_noop_backend")
    backend.plot = lambda *a, **k: None
    sys.modules["cherry_noop_backend"] = backend
    pd.options.plotting.backend = "cherry_noop_backend"
    if root not in sys.path:
        sys.path.insert(0, root)
    import cherryml  # noqa: F401
    import matplotlib.pyplot as plt

    plt.subplots = lambda *a, **k: (mock.MagicMock(), mock.MagicMock())
    return cherryml
