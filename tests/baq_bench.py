#!/usr/bin/env python
"""Throughput of the BAQ HMM (lfb200_kpa_glocal_batch) next to the compiled reference on one host core.
usage: baq_bench.py [reads]   — prints one JSON line; under ncu (-k regex:k_kpa_glocal) the kernel itself is captured"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))      # tests/ may use the oracle (reference timing, read generator)
sys.path.insert(0, ROOT)
import lofreq_b200
from oracle import pyoracle

n_target = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
base = pyoracle.synth_reads(10000, seed=5, lmin=100, lmax=152)
rep = max(1, n_target // base["n"])
reads = dict(n=base["n"] * rep, ref=np.tile(base["ref"], rep), query=np.tile(base["query"], rep), qual=np.tile(base["qual"], rep),
             ref_off=np.concatenate([[0], np.cumsum(np.tile(np.diff(base["ref_off"]), rep))]).astype(np.int64),
             qry_off=np.concatenate([[0], np.cumsum(np.tile(np.diff(base["qry_off"]), rep))]).astype(np.int64))
c = lofreq_b200.Caller()
c.kpa_glocal(reads)
t0 = time.perf_counter()
st, q = c.kpa_glocal(reads)
t_gpu = time.perf_counter() - t0
line = {"metric": "baq_hmm_reads_per_sec", "reads": reads["n"], "mean_read_length": float(reads["qry_off"][-1]) / reads["n"],
        "e2e_reads_per_sec": reads["n"] / t_gpu, "e2e_seconds": t_gpu,
        "e2e_region": "lfb200_kpa_glocal_batch with host buffers: H2D of windows / reads / qualities, k_kpa_glocal, D2H of state[] and q[]"}
if pyoracle.have_kpa_reference():
    ref = pyoracle.KpaRef()
    t0 = time.perf_counter()
    ws, wq, _ = ref.glocal(base)
    t_cpu = time.perf_counter() - t0
    line["cpu_baseline"] = {"value": base["n"] / t_cpu, "unit": "reads/s", "cores": 1, "kind": "reference",
                            "sample": "%d reads, kpa_ext_glocal of the unmodified kprobaln_ext.c" % base["n"]}
    line["identical_to_reference"] = bool(np.array_equal(st[:len(ws)], ws) and np.array_equal(q[:len(wq)], wq))
print(json.dumps(line))
