"""Pins the CPU oracle (oracle/snv_oracle.c, the restatement) against
 (a) the committed golden fixtures generated from the unmodified reference
     (tests/golden/make_golden.py) — bit-for-bit, including the long double images;
 (b) the compiled reference itself (oracle/_ref/libsnpref.so) when it is present.
No GPU involved."""
import glob
import os

import numpy as np
import pytest

from oracle import synth_np
from oracle.pyoracle import default_conf

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def ld_from_bytes(b):
    return np.ascontiguousarray(b).view(np.longdouble).reshape(b.shape[:-1])


def load_conf(z):
    conf = {k: v for k, v in zip(z["conf_keys"].tolist(), z["conf_vals"].tolist())}
    for k in conf:
        if k != "sig":
            conf[k] = int(conf[k])
    return conf


def test_known_answer_from_reference_comment(port_oracle):
    # snpcaller.c:1222-1232: ten reads with p = 0.001, at least one error; R ppoibin gives 0.00995512
    pv = port_oracle.snpcaller(np.full(10, 0.001), (1, 0, 0), 1, 1.0)
    assert abs(float(pv[0]) - 0.00995512) < 5e-9
    assert abs(float(pv[0]) - (1.0 - 0.999 ** 10)) < 1e-15


def test_snpcaller_grid_golden(port_oracle):
    z = np.load(os.path.join(GOLD, "snpcaller_grid.npz"))
    want = ld_from_bytes(z["pvalue_ld"])
    offs = z["offsets"]
    for i, name in enumerate(z["names"].tolist()):
        ep = z["err_probs"][offs[i]:offs[i + 1]]
        got = port_oracle.snpcaller(ep, z["counts"][i], int(z["bonf"][i]), float(z["sig"][i]))
        assert np.array_equal(got, want[i]), name


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "batch_*.npz"))))
def test_batch_golden(port_oracle, path):
    z = np.load(path)
    b = synth_np.generate(str(z["workload"]), int(z["c0"]), int(z["n_cols"]), with_baq=bool(z["with_baq"]))
    out = port_oracle.call_columns(b, load_conf(z))
    for k in ("alt_counts", "alt_raw_counts", "tested", "bonf_used", "called", "qual"):
        assert np.array_equal(out[k], z[k]), k
    assert np.array_equal(out["pvalues"], ld_from_bytes(z["pvalue_ld"]))
    assert out["bonf_subst"] == int(z["bonf_subst"]) and out["num_snv_tests"] == int(z["num_snv_tests"])


def test_scalar_helpers_match_reference(port_oracle, ref_oracle):
    for q in range(0, 256):
        assert port_oracle.phred_to_prob(q) == ref_oracle.phred_to_prob(q)
    rng = np.random.default_rng(7)
    for _ in range(2000):
        sq, mq, baq, bq = (int(rng.integers(-1, 94)), int(rng.integers(-1, 256)),
                           int(rng.integers(-1, 94)), int(rng.integers(0, 94)))
        a, b = port_oracle.merge(sq, mq, baq, bq), ref_oracle.merge(sq, mq, baq, bq)
        assert a == b
        assert port_oracle.prob_to_phred_safe(a) == ref_oracle.prob_to_phred_safe(b)
        x, y = float(-rng.random() * 800), float(-rng.random() * 800)
        assert port_oracle.log_sum(x, y) == ref_oracle.log_sum(x, y)


def test_port_vs_reference_random_columns(port_oracle, ref_oracle):
    rng = np.random.default_rng(11)
    for wl, c0, n, baq in (("C2", 12345, 400, True), ("C3", 99, 60, False), ("C5", 777, 60, True)):
        b = synth_np.generate(wl, c0, n, with_baq=baq)
        conf = default_conf(bonf_subst=int(rng.integers(1, 10**6)))
        a = port_oracle.call_columns(b, dict(conf))
        r = ref_oracle.call_columns(b, dict(conf))
        for k in ("alt_counts", "alt_raw_counts", "tested", "bonf_used", "pvalues", "called", "qual"):
            assert np.array_equal(a[k], r[k]), (wl, k)
        assert a["bonf_subst"] == r["bonf_subst"] and a["num_snv_tests"] == r["num_snv_tests"]


def test_poissbin_row_matches_reference(port_oracle, ref_oracle):
    ep = np.sort(10.0 ** (-np.random.default_rng(3).integers(20, 41, 300) / 10.0))
    for k, bonf in ((1, 1), (5, 1), (5, 30000), (40, 300), (300, 1)):
        pa, ra = port_oracle.poissbin(ep, k, bonf, 0.05)
        pr, rr = ref_oracle.poissbin(ep, k, bonf, 0.05)
        assert pa == pr and np.array_equal(ra[:k + 1], rr[:k + 1])


def test_poissbin_rows_golden(port_oracle):
    """the port's poissbin() rows — partial rows after the early exit included — are the reference's, bit for bit"""
    z = np.load(os.path.join(GOLD, "poissbin_rows.npz"))
    for i, name in enumerate(z["names"]):
        ep = z["err_probs"][z["offsets"][i]:z["offsets"][i + 1]]
        pv, row = port_oracle.poissbin(ep, int(z["k"][i]), int(z["bonf"][i]), float(z["sig"][i]))
        want = z["rows"][z["row_offsets"][i]:z["row_offsets"][i + 1]]
        assert np.array_equal(row, want), name
        assert pv == ld_from_bytes(z["pvalue_ld"][i]), name


def test_binom_reference_vs_scipy():
    from oracle.pyoracle import BinomRef, BINOM_SO
    if not os.path.exists(BINOM_SO):
        pytest.skip("oracle/_ref/libbinomref.so not built")
    from scipy.stats import binom as sb
    br = BinomRef()
    for n, k, p in ((500, 3, 0.05), (100, 50, 0.5), (10000, 120, 0.01), (30, 0, 0.2)):
        cdf, sf = br.cdf_sf(n, k, p)
        assert abs(cdf - sb.cdf(k, n, p)) <= 1e-12 * max(cdf, 1e-300) + 1e-15
        assert abs(sf - sb.sf(k, n, p)) <= 1e-12 * max(sf, 1e-300) + 1e-15


def test_binom_golden_is_the_reference():
    """tests/golden/binom_grid.npz are outputs of the compiled reference binom()"""
    from oracle.pyoracle import BinomRef, BINOM_SO
    if not os.path.exists(BINOM_SO):
        pytest.skip("oracle/_ref/libbinomref.so not built")
    import ctypes as C
    br = BinomRef()
    z = np.load(os.path.join(GOLD, "binom_grid.npz"))
    for i in range(0, len(z["status"]), 3):
        a, b = C.c_double(float("nan")), C.c_double(float("nan"))
        rc = br.lib.binom(C.byref(a), C.byref(b), int(z["num_trials"][i]), int(z["num_success"][i]), float(z["prob_success"][i]))
        assert rc == int(z["status"][i])
        if rc == 0:
            assert a.value == z["cdf"][i] and b.value == z["sf"][i]
