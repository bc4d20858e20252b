"""Run test files of the UNMODIFIED reference (read from /root/reference, build container only)
against THIS package: ``cherryml`` and its sub-modules are aliased to ``cherryml_b200`` before
the reference's unittest modules are imported, the working directory is the reference checkout
(its tests use relative data paths), nothing is written there.

    python tests/golden/run_reference_tests.py /root/reference/tests/io_tests/msa_test.py \\
        /root/reference/tests/evaluation_tests/metrics_test.py /root/reference/tests/estimation_tests/jtt_ipw_test.py

Those three files (9 tests) need no GPU and pass.  The reference's counting, fit, likelihood,
FastCherries and SiteRM tests need a GPU, which the build container does not have, while the GPU
box does not have the reference: their cases are restated in tests/test_gpu_*.py instead.  Run
here, counting_test.py and quantized_transitions_mle_test.py (42 tests) bind to this package's
API without a single TypeError / AttributeError / ImportError: 6 pass and 36 stop at "Found no
NVIDIA driver", i.e. where the CUDA path starts; likelihood_test.py (84 tests): 42 pass (input
validation, file handling), 42 stop at the same place, none on an API mismatch.  (`parameterized` is not installed: a minimal
stand-in for ``parameterized.expand`` is injected.)
(``assertEquals`` is aliased because Python 3.12 removed it.)
"""
import importlib, importlib.util, os, sys, types, unittest, pkgutil
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/repo")
import cherryml_b200
# alias the package and its sub-modules under the reference's name
sys.modules["cherryml"] = cherryml_b200
for sub in ("io", "caching", "utils", "types", "counting", "estimation", "estimation_end_to_end", "evaluation",
            "markov_chain", "phylogeny_estimation", "siterm"):
    m = importlib.import_module("cherryml_b200." + sub)
    sys.modules["cherryml." + sub] = m
unittest.TestCase.assertEquals = unittest.TestCase.assertEqual
try:
    import parameterized  # noqa
except Exception:
    # minimal stand-in for parameterized.expand: one method per parameter tuple, injected into
    # the class body that is being executed
    import inspect

    class _Parameterized:
        @staticmethod
        def expand(cases):
            def decorator(f):
                scope = inspect.currentframe().f_back.f_locals
                for i, case in enumerate(cases):
                    args = case if isinstance(case, (tuple, list)) else (case,)
                    label = str(args[0]).replace(" ", "_") if args else str(i)
                    scope[f"{f.__name__}_{i}_{label}"] = (lambda a: lambda self: f(self, *a))(tuple(args))
                return None
            return decorator

    shim = types.ModuleType("parameterized")
    shim.parameterized = _Parameterized
    sys.modules["parameterized"] = shim
os.chdir("/root/reference")
sys.path.insert(0, "/root/reference")  # the reference's test files import helpers from ITS `tests` package
for name in [m for m in sys.modules if m == "tests" or m.startswith("tests.")]:
    del sys.modules[name]
suite = unittest.TestSuite()
for path in sys.argv[1:]:
    spec = importlib.util.spec_from_file_location("reftest_" + os.path.basename(path)[:-3], path)
    mod = importlib.util.module_from_spec(spec)
    try:
        spec.loader.exec_module(mod)
    except Exception as e:
        print("IMPORT FAILED", path, type(e).__name__, e)
        continue
    suite.addTests(unittest.defaultTestLoader.loadTestsFromModule(mod))
res = unittest.TextTestRunner(verbosity=1).run(suite)
print("ran", res.testsRun, "failures", len(res.failures), "errors", len(res.errors), "skipped", len(res.skipped))
for t, tb in (res.failures + res.errors)[:8]:
    print("----", t); print(tb[-700:])
