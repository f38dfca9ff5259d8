"""The BAQ HMM (kpa_ext_glocal, kprobaln_ext.c:80-277) — CPU side.
The arithmetic of the device routine (lofreq_b200/csrc/baq_core.cuh) compiled for the host, against the compiled reference:
posterior state and quality of every base identical.  Pins the kernel's arithmetic without a GPU."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden", "kpa_glocal.npz")


def _host_lib(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("baq") / "libbaqhost.so")
    src = os.path.join(ROOT, "tests", "harness", "baq_host_harness.cpp")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", out, src, "-lm"])
    lib = C.CDLL(out)
    lib.lfb_kpa_host_batch.restype = C.c_int
    return lib


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    return _host_lib(tmp_path_factory)


def _run_host(lib, reads, d, e, bw, use_qual=True):
    tot = int(reads["qry_off"][-1])
    state = np.zeros(tot, np.int32)
    q = np.zeros(tot, np.uint8)
    nfix = C.c_longlong(0)
    p = pyoracle._ptr
    lib.lfb_kpa_host_batch(C.c_longlong(reads["n"]), p(reads["ref"]), p(reads["ref_off"]), p(reads["query"]), p(reads["qry_off"]),
                           p(reads["qual"]) if use_qual else None, C.c_float(d), C.c_float(e), C.c_int(bw), p(state), p(q), C.byref(nfix))
    return state, q, int(nfix.value)


PARS = [(0.00001, 0.4, 10), (0.001, 0.1, 10), (0.1, 0.4, 10), (0.0001, 0.01, 3)]      # kpa_ext_par_lofreq_illumina, _def, _pacbio, a narrow band


@pytest.mark.skipif(not pyoracle.have_kpa_reference(), reason="compiled reference not available")
@pytest.mark.parametrize("d,e,bw", PARS)
def test_host_instance_matches_the_compiled_reference(host_lib, d, e, bw):
    ref = pyoracle.KpaRef()
    for seed, kw in ((1, {}), (2, dict(lmin=1, lmax=12, flank=3)), (3, dict(sub=0.1, indel=0.03)), (4, dict(lmin=200, lmax=260, flank=25))):
        reads = pyoracle.synth_reads(400, seed=seed, **kw)
        for use_qual in (True, False):
            ws, wq, _ = ref.glocal(reads, d, e, bw, use_qual)
            gs, gq, _ = _run_host(host_lib, reads, d, e, bw, use_qual)
            assert np.array_equal(gs, ws), ("state", seed, np.argwhere(gs != ws)[:5].tolist())
            assert np.array_equal(gq, wq), ("q", seed, np.argwhere(gq != wq)[:5].tolist())


@pytest.mark.skipif(not pyoracle.have_kpa_reference(), reason="compiled reference not available")
def test_degenerate_inputs(host_lib):
    """qualities 0 (probability 1) and 255, reads / windows made of Ns, bands of 1 and 2, windows much longer than the read"""
    ref = pyoracle.KpaRef()
    rng = np.random.default_rng(99)
    for trial in range(15):
        r = pyoracle.synth_reads(120, seed=1000 + trial, lmin=1, lmax=60, flank=int(rng.integers(0, 30)))
        mode = trial % 5
        if mode == 0:
            r["qual"][:] = 0
        elif mode == 1:
            r["qual"][:] = rng.choice([0, 1, 2, 93, 200, 255], len(r["qual"]))
        elif mode == 2:
            r["query"][:] = 4
        elif mode == 3:
            r["ref"][:] = 4
        else:
            r["qual"][:] = rng.integers(0, 256, len(r["qual"]))
        for d, e, bw in ((0.00001, 0.4, 10), (0.5, 0.9, 2), (0.001, 0.1, 1)):
            ws, wq, _ = ref.glocal(r, d, e, bw)
            gs, gq, _ = _run_host(host_lib, r, d, e, bw)
            assert np.array_equal(gs, ws) and np.array_equal(gq, wq), (trial, mode, d, e, bw)


def test_host_instance_matches_the_golden_vectors(host_lib):
    z = np.load(GOLD)
    for i in range(int(z["n_cases"])):
        reads = dict(n=int(z["n_%d" % i]), ref=z["ref_%d" % i], ref_off=z["ref_off_%d" % i], query=z["query_%d" % i],
                     qry_off=z["qry_off_%d" % i], qual=z["qual_%d" % i])
        d, e, bw = float(z["d_%d" % i]), float(z["e_%d" % i]), int(z["bw_%d" % i])
        gs, gq, _ = _run_host(host_lib, reads, d, e, bw, bool(z["use_qual_%d" % i]))
        assert np.array_equal(gs, z["state_%d" % i]) and np.array_equal(gq, z["q_%d" % i]), i
