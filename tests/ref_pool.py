"""TEST INFRASTRUCTURE — the CPU checker over many columns on all host cores.

Chunks of consecutive synthetic columns go to a process pool; every worker regenerates its chunk with the numpy
generator (oracle/synth_np.py) and runs the compiled reference (oracle/_ref/libsnpref.so, when it travelled to this
box) or the C restatement over it, exactly like call_snvs walks the columns.  The running Bonferroni factor a chunk
starts from (lofreq_call.c:794-800) is derived from the `tested` flags of the chunks before it — the caller passes
the flags it wants checked (the GPU's), every worker returns its own, and the caller compares them chunk by chunk, so
a wrong start cannot go unnoticed: chunk 0 starts from the caller's conf, and chunk j's start is right as soon as the
flags of chunks 0..j-1 agree."""
import multiprocessing as mp
import os

import numpy as np


def oracle_kind():
    from oracle.pyoracle import have_reference
    return "reference" if have_reference() else "port"


def _work(args):
    kind, wl, c0, n, with_baq, conf = args
    from oracle import synth_np
    from oracle.pyoracle import Oracle
    b = synth_np.generate(wl, c0, n, with_baq=with_baq)
    out = Oracle(kind).call_columns(b, dict(conf))
    out["pvalue_bytes"] = out.pop("pvalues").tobytes()       # long double does not pickle portably
    return c0, out


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_chunks(wl, c0, n_cols, tested, conf, chunk=20000, with_baq=False, procs=None):
    """Oracle results for columns [c0, c0 + n_cols) as a list of (lo, hi, out) per chunk.  `tested`: uint8 flags of the
    same columns from the implementation under test (used only to derive each chunk's starting factor)."""
    kind = oracle_kind()
    jobs = []
    start, before = conf["bonf_subst"], 0
    for lo in range(0, n_cols, chunk):
        hi = min(n_cols, lo + chunk)
        cf = dict(conf)
        if conf.get("bonf_dynamic", 1) and before > 0:
            cf["bonf_subst"] = (0 if start == 1 else start) + 3 * before
        jobs.append((kind, wl, c0 + lo, hi - lo, with_baq, cf))
        before += int(np.asarray(tested[lo:hi]).sum())
    procs = procs or max(1, min(host_cores(), len(jobs)))
    if procs == 1:
        res = [_work(j) for j in jobs]
    else:
        with mp.get_context("spawn").Pool(procs) as pool:
            res = pool.map(_work, jobs, chunksize=1)
    out = []
    for (_, _, cc0, n, _, _), (_, o) in zip(jobs, res):
        o["pvalues"] = np.frombuffer(o.pop("pvalue_bytes"), dtype=np.longdouble).reshape(n, 3).copy()
        out.append((cc0 - c0, cc0 - c0 + n, o))
    return out, kind
