"""The BAQ HMM on the GPU (lfb200_kpa_glocal_batch = kpa_ext_glocal per read, kprobaln_ext.c:80-277):
state[] and q[] of every base identical to the compiled reference / its golden vectors, through the C ABI."""
import os
import sys
import time

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle  # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = os.path.join(ROOT, "tests", "golden", "kpa_glocal.npz")


@pytest.fixture(scope="module")
def caller():
    import lofreq_b200
    c = lofreq_b200.Caller()
    yield c
    c.close()


def test_golden_vectors(caller):
    z = np.load(GOLD)
    for i in range(int(z["n_cases"])):
        reads = dict(n=int(z["n_%d" % i]), ref=z["ref_%d" % i], ref_off=z["ref_off_%d" % i], query=z["query_%d" % i],
                     qry_off=z["qry_off_%d" % i], qual=z["qual_%d" % i])
        st, q = caller.kpa_glocal(reads, float(z["d_%d" % i]), float(z["e_%d" % i]), int(z["bw_%d" % i]), bool(z["use_qual_%d" % i]))
        assert np.array_equal(st, z["state_%d" % i]), (i, "state", np.argwhere(st != z["state_%d" % i])[:5].tolist())
        assert np.array_equal(q, z["q_%d" % i]), (i, "q", np.argwhere(q != z["q_%d" % i])[:5].tolist())


def test_empty_and_degenerate_reads(caller):
    # a batch with an empty query, an empty window, a single base, and a query longer than its window
    ref = np.array([0, 1, 2, 3, 0, 1, 2, 3, 1, 0, 1, 2, 3, 0], np.uint8)
    ref_off = np.array([0, 4, 4, 5, 8, 14], np.int64)
    query = np.array([0, 1, 2, 3, 2, 0, 1, 2, 3, 1, 2, 0, 1, 3, 3, 3, 0], np.uint8)
    qry_off = np.array([0, 0, 3, 4, 11, 17], np.int64)
    reads = dict(n=5, ref=ref, ref_off=ref_off, query=query, qry_off=qry_off, qual=np.full(17, 25, np.uint8))
    st, q = caller.kpa_glocal(reads)
    if pyoracle.have_kpa_reference():
        ws, wq, _ = pyoracle.KpaRef().glocal(reads)
        assert np.array_equal(st, ws) and np.array_equal(q, wq)
    assert np.all(st[0:3] == 0) and np.all(q[0:3] == 0)          # the read with an empty window is left untouched


@pytest.mark.skipif(not pyoracle.have_kpa_reference(), reason="compiled reference not available")
def test_against_the_compiled_reference_and_timing(caller):
    ref = pyoracle.KpaRef()
    reads = pyoracle.synth_reads(20000, seed=21, lmin=100, lmax=152)
    t0 = time.perf_counter()
    ws, wq, _ = ref.glocal(reads)
    t_cpu = time.perf_counter() - t0
    caller.kpa_glocal(reads)                                     # warm-up (scratch allocation)
    t0 = time.perf_counter()
    st, q = caller.kpa_glocal(reads)
    t_gpu = time.perf_counter() - t0
    assert np.array_equal(st, ws), np.argwhere(st != ws)[:5].tolist()
    assert np.array_equal(q, wq), np.argwhere(q != wq)[:5].tolist()
    print("BAQ HMM: %d reads (mean length %.0f): reference %.0f reads/s on one core, lfb200_kpa_glocal_batch %.0f reads/s "
          "(host buffers in and out)" % (reads["n"], reads["qry_off"][-1] / reads["n"], reads["n"] / t_cpu, reads["n"] / t_gpu))
    for d, e, bw, kw in ((0.001, 0.1, 10, dict(sub=0.08, indel=0.02)), (0.1, 0.4, 10, dict(lmin=1, lmax=40, flank=4)),
                         (0.00001, 0.4, 4, dict(lmin=180, lmax=260, flank=30))):
        r2 = pyoracle.synth_reads(1500, seed=int(bw * 7 + 1), **kw)
        for uq in (True, False):
            ws, wq, _ = ref.glocal(r2, d, e, bw, uq)
            st, q = caller.kpa_glocal(r2, d, e, bw, uq)
            assert np.array_equal(st, ws) and np.array_equal(q, wq), (d, e, bw, uq)
