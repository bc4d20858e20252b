"""The UNMODIFIED reference Python package, made importable for baselines (TEST / BENCH
INFRASTRUCTURE -- nothing under cherryml_b200/ imports this).

The GPU box has no /root/reference, and `bench.py --impl reference` / the `fit` section must
time the reference's own ``quantized_transitions_mle`` there (CPU arm with
all host threads and the reference's stock ``device="cuda"`` path on the same B200).  So, like
the reference C++ programs compiled into ``oracle/_ref``, the build step packs the
reference's ``cherryml/*.py`` tree into ONE git-ignored archive ``oracle/_ref/reference_package.tar.gz``
(committed recipe; no reference source file enters the repository or its history) and the snapshot carries it
to the box, where it is unpacked into a temporary directory for the lifetime of the process that times it.

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
ARCHIVE = os.path.join(HERE, "_ref", "reference_package.tar.gz")
_EXTRACTED = {}


def build_reference_package() -> bool:
    """Pack cherryml/**/*.py of the reference checkout into ONE git-ignored archive under oracle/_ref (build
    container only; no reference source file is ever placed in the tree).  Returns True if the archive exists
    afterwards."""
    import tarfile

    src = os.path.join(REF_CHECKOUT, "cherryml")
    if os.path.isdir(src):
        newest = max(os.path.getmtime(os.path.join(r, f)) for r, _, fs in os.walk(src) for f in fs if f.endswith(".py"))
        if not (os.path.exists(ARCHIVE) and os.path.getmtime(ARCHIVE) >= newest):
            os.makedirs(os.path.dirname(ARCHIVE), exist_ok=True)
            tmp = ARCHIVE + ".tmp"
            with tarfile.open(tmp, "w:gz") as tf:
                for root, _, files in os.walk(src):
                    for f in sorted(files):
                        if f.endswith(".py"):
                            full = os.path.join(root, f)
                            tf.add(full, arcname=os.path.join("cherryml", os.path.relpath(full, src)))
            os.replace(tmp, ARCHIVE)
        shutil.rmtree(os.path.join(HERE, "_ref", "pkg"), ignore_errors=True)  # the unpacked tree of earlier builds
    return os.path.exists(ARCHIVE)


def reference_root():
    """Directory that holds the reference's ``cherryml`` package (the checkout in the build container, else the
    archive unpacked into a temporary directory for the lifetime of this process), or None."""
    if os.path.isdir(os.path.join(REF_CHECKOUT, "cherryml")):
        return REF_CHECKOUT
    if os.path.exists(ARCHIVE):
        if "dir" not in _EXTRACTED:
            import atexit
            import tarfile
            import tempfile

            d = tempfile.mkdtemp(prefix="cherry_ref_pkg_")
            atexit.register(shutil.rmtree, d, True)
            with tarfile.open(ARCHIVE) as tf:
                tf.extractall(d)
            _EXTRACTED["dir"] = d
        return _EXTRACTED["dir"]
    return None


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_CHECKOUT, "cherryml")) or os.path.exists(ARCHIVE)


def import_reference():
    """Import the reference package (see the module docstring for the stand-ins)."""
    root = reference_root()
    if root is None:
        raise ImportError("neither /root/reference nor oracle/_ref/reference_package.tar.gz is present")
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
