"""Extract the known-answer tests of the reference's own C++ unit tests for FastCherries
(cherryml/phylogeny_estimation/FastCherries/tests/test_branch_length_estimation.cpp: the
test_branch_lengths* and test_get_site_rates* cases, with tests/lg.txt) into
tests/golden/fast_cherries/kats.json.  Data only (inputs and expected indices); build
container only.  test_pairing_algorithms.cpp is not usable: it still calls the two-argument,
rand()-seeded divide_and_pair that the program no longer has.
"""
import ast
import json
import os
import re

SRC = "/root/reference/cherryml/phylogeny_estimation/FastCherries/tests"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fast_cherries", "kats.json")


def braces(text):
    return ast.literal_eval(text.replace("{", "[").replace("}", "]"))


def main():
    src = open(os.path.join(SRC, "test_branch_length_estimation.cpp")).read()
    Q = [[float(v) for v in ln.split()] for ln in open(os.path.join(SRC, "lg.txt")).read().strip().split("\n")]

    def body(name):
        m = re.search(r"void %s\s*\([^)]*\)\s*\{(.*?)\n\}" % name, src, re.S)
        return m.group(1)

    def grid_and(name, *fields):
        b = body(name)
        out = {"grid": braces(re.search(r"quantization_points = (\{.*?\});", b, re.S).group(1))}
        for f in fields:
            out[f] = braces(re.search(r"%s = (\{.*?\});" % f, b, re.S).group(1))
        return out

    bl = grid_and("test_branch_lengths", "rate_categories", "site_to_rate")
    sr = grid_and("test_get_site_rates", "rate_categories", "lengths_index")
    kats = {"rate_matrix": Q, "branch_lengths": dict(bl, cases=[]), "site_rates": dict(sr, cases=[])}
    for kind, prefix in (("branch_lengths", "test_branch_lengths"), ("site_rates", "test_get_site_rates")):
        for i in range(1, 10):
            try:
                b = body(f"{prefix}{i}")
            except AttributeError:
                break
            cherries = braces(re.search(r"cherries = (\{.*?\});", b, re.S).group(1))
            expected = braces(re.search(r"expected = (\{.*?\});", b, re.S).group(1))
            kats[kind]["cases"].append({"cherries": cherries, "expected": expected})
    with open(OUT, "w") as f:
        json.dump(kats, f)
    print({k: len(v["cases"]) for k, v in kats.items() if isinstance(v, dict)}, os.path.getsize(OUT))


if __name__ == "__main__":
    main()
