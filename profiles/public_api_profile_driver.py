import sys, os, tarfile, tempfile, time, cProfile, pstats, io
sys.path.insert(0, os.getcwd())
from cherryml_b200 import cherryml_public_api, caching
root = tempfile.mkdtemp()
with tarfile.open("tests/golden/demo_data.tar.xz") as tf:
    tf.extractall(root)
def run(model, est=None, tag=""):
    cache = tempfile.mkdtemp(); out = os.path.join(cache, "Q.txt")
    kw = dict(output_path=out, model_name=model, msa_dir=f"{root}/msas", cache_dir=cache, num_epochs=500 if model=="LG" else 30)
    if est is None:
        kw.update(tree_dir=f"{root}/trees")
        if model == "LG": kw.update(site_rates_dir=f"{root}/site_rates")
    else:
        kw.update(tree_estimator_name=est)
    if model != "LG": kw.update(contact_map_dir=f"{root}/contact_maps")
    t = time.time(); cherryml_public_api(**kw); return time.time() - t
run("LG")  # warm-up (CUDA context, module load)
for model, est in (("LG", None), ("LG", "FastCherries"), ("co-evolution", None), ("co-evolution", "FastCherries")):
    pr = cProfile.Profile(); pr.enable(); dt = run(model, est); pr.disable()
    print(f"=== {model} trees={'given' if est is None else est}: {dt:.2f} s")
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(14); print("\n".join(l for l in s.getvalue().split("\n")[8:26]))
