#!/usr/bin/env python
"""Golden vectors of the BAQ HMM from the compiled reference (oracle/_ref/libkparef.so = the unmodified kprobaln_ext.c):
   python tests/golden/make_golden_kpa.py   ->  tests/golden/kpa_glocal.npz"""
import os
import sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle

ref = pyoracle.KpaRef()
out = {}
cases = [(11, {}, 0.00001, 0.4, 10, True), (12, dict(lmin=1, lmax=12, flank=3), 0.001, 0.1, 10, True),
         (13, dict(sub=0.1, indel=0.03), 0.1, 0.4, 10, True), (14, dict(lmin=200, lmax=260, flank=25), 0.00001, 0.4, 10, False),
         (15, dict(lmin=60, lmax=101), 0.0001, 0.01, 3, True)]
# degenerate inputs: qualities 0 / 255, reads and windows of Ns only, a band of 1
cases += [(16, dict(lmin=1, lmax=60, flank=25, _qual="extreme"), 0.1, 0.4, 10, True), (17, dict(lmin=1, lmax=60, flank=5, _query=4), 0.00001, 0.4, 10, True),
          (18, dict(lmin=5, lmax=80, flank=12, _ref=4), 0.5, 0.9, 2, True), (19, dict(lmin=1, lmax=50, flank=20, _qual="any"), 0.001, 0.1, 1, True)]
for i, (seed, kw, d, e, bw, uq) in enumerate(cases):
    kw = dict(kw)
    special = {k: kw.pop(k) for k in list(kw) if k.startswith("_")}
    r = pyoracle.synth_reads(150, seed=seed, **kw)
    rng = np.random.default_rng(seed)
    if special.get("_qual") == "extreme":
        r["qual"][:] = rng.choice([0, 1, 2, 93, 200, 255], len(r["qual"]))
    if special.get("_qual") == "any":
        r["qual"][:] = rng.integers(0, 256, len(r["qual"]))
    if "_query" in special:
        r["query"][:] = special["_query"]
    if "_ref" in special:
        r["ref"][:] = special["_ref"]
    st, q, pr = ref.glocal(r, d, e, bw, uq)
    for k in ("ref", "ref_off", "query", "qry_off", "qual"):
        out["%s_%d" % (k, i)] = r[k]
    out["n_%d" % i] = r["n"]
    out["d_%d" % i], out["e_%d" % i], out["bw_%d" % i], out["use_qual_%d" % i] = d, e, bw, uq
    out["state_%d" % i], out["q_%d" % i] = st, q
out["n_cases"] = len(cases)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "kpa_glocal.npz"), **out)
print("wrote", len(cases), "cases")
