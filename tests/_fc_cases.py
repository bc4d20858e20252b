"""Shared loader of the FastCherries goldens (tests/golden/fast_cherries)."""
import gzip
import json
import os
import tarfile

import numpy as np

from tests.conftest import GOLDEN

_DEMO_CACHE = {}


def load_cases():
    with gzip.open(os.path.join(GOLDEN, "fast_cherries", "cases.json.gz"), "rt") as f:
        return json.load(f)


def load_kats():
    with open(os.path.join(GOLDEN, "fast_cherries", "kats.json")) as f:
        return json.load(f)


def msa_text(case) -> str:
    if case["demo_family"] is None:
        return case["msa_text"]
    fam = case["demo_family"]
    if fam not in _DEMO_CACHE:
        with tarfile.open(os.path.join(GOLDEN, "demo_data.tar.xz")) as tf:
            _DEMO_CACHE[fam] = tf.extractfile(f"msas/{fam}.txt").read().decode()
    return _DEMO_CACHE[fam]


def parse_msa(text):
    lines = text.split("\n")
    if lines and lines[-1] == "":
        lines.pop()
    names, seqs = [], []
    i = 0
    while i < len(lines):
        if lines[i][:1] == ">" and i + 1 < len(lines):
            names.append(lines[i][1:])
            seqs.append(lines[i + 1])
            i += 2
        else:
            i += 1
    return names, seqs


def parse_rate_matrix(text):
    lines = text.strip().split("\n")
    alphabet = lines[0].split()
    Q = np.array([[float(v) for v in ln.split()[1:]] for ln in lines[1:]])
    return alphabet, Q


def expected_outputs(case):
    toks = case["cherries_file"].split("\n")
    if toks and toks[-1] == "":
        toks.pop()
    cherries = [(toks[j], toks[j + 1]) for j in range(0, len(toks), 3)]
    distances = [toks[j + 2] for j in range(0, len(toks), 3)]
    rl = case["site_rates_file"].split("\n")
    rates = rl[1].split() if len(rl) > 1 else []
    return cherries, distances, rates
