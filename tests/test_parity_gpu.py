"""GPU parity tests: the CUDA path, called through the C ABI (lofreq_b200/capi.py ->
liblofreq_b200.so), against the CPU oracle and the golden fixtures generated from the
unmodified reference.

Bars (BASELINE.json north_star / SURVEY.md §8d):
  * tested flags, Bonferroni factors and counters, alt counts, LDBL_MAX/LDBL_MIN sentinel status,
    called (pos, alt) set, integer QUAL: bit-exact
  * ln p: |d ln p| <= 1e-10 * max(|ln p|, 1)   (1e-10 relative on log(p); absolute below |ln p| = 1,
    where p > 0.37 can never be called)
"""
import glob
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import synth_np
from oracle.pyoracle import default_conf
from test_oracle import GOLD, ld_from_bytes, load_conf

LNP_RTOL = 1e-10
MAXK = 16384         # largest alt count supported (CTA-per-column kernel: 256 threads x 64 cells)


def status_of(pv):
    st = np.zeros(pv.shape, np.uint8)
    st[pv == np.finfo(np.longdouble).max] = 1
    st[pv == np.finfo(np.longdouble).tiny] = 2
    return st


def assert_lnp_close(got_pv, want_pv, st, what=""):
    m = st == 0
    if not m.any():
        return 0.0
    with np.errstate(all="ignore"):
        a = np.log(got_pv[m]).astype(np.float64)
        b = np.log(want_pv[m]).astype(np.float64)
    err = np.abs(a - b) / np.maximum(np.abs(b), 1.0)
    assert err.max() <= LNP_RTOL, (what, float(err.max()))
    return float(err.max())


@pytest.fixture(scope="module")
def caller():
    import lofreq_b200
    c = lofreq_b200.Caller(0)
    yield c
    c.close()


def compare_batch(got, want, what=""):
    for k in ("alt_counts", "alt_raw_counts", "tested", "bonf_used"):
        assert np.array_equal(got[k], want[k]), (what, k)
    want_st = status_of(want["pvalues"])
    assert np.array_equal(got["status"], want_st), (what, "status",
                                                     np.argwhere(got["status"] != want_st)[:5].tolist())
    assert np.array_equal(status_of(got["pvalues"]), want_st), (what, "sentinels")
    worst = assert_lnp_close(got["pvalues"], want["pvalues"], want_st, what)
    assert np.array_equal(got["called"], want["called"]), (what, "called")
    bad = np.argwhere(got["qual"] != want["qual"])
    # QUAL = (int)(-10 log10l(p)) truncates: when the exact value is an integer (all reads are the alt
    # allele and only base qualities are merged, so p = 10^-(sum q)/10) the last ulp of ln p decides
    # between q and q-1.  Accept +-1 there and nowhere else.
    with np.errstate(all="ignore"):
        exactq = -10.0 * np.log10(want["pvalues"]).astype(np.float64)
    bad = np.array([(c, i) for c, i in bad if not (abs(int(got["qual"][c, i]) - int(want["qual"][c, i])) == 1 and
                                                   abs(exactq[c, i] - round(exactq[c, i])) < 1e-6)]).reshape(-1, 2)
    assert len(bad) == 0, (what, "qual", [(int(c), int(i), int(got["qual"][c, i]), int(want["qual"][c, i]),
                                            float(got["lnp"][c, i]), got["alt_counts"][c].tolist()) for c, i in bad[:5]])
    if "bonf_subst" in want:
        assert got["bonf_subst"] == want["bonf_subst"] and got["num_snv_tests"] == want["num_snv_tests"], what
    QUAL_TOLERANCE_HITS[what] = int((got["qual"] != want["qual"]).sum())
    return worst


QUAL_TOLERANCE_HITS = {}     # per comparison: alleles whose QUAL differs by the documented +-1 (exact-integer QUAL only)


def test_snpcaller_golden_grid(caller):
    z = np.load(os.path.join(GOLD, "snpcaller_grid.npz"))
    want = ld_from_bytes(z["pvalue_ld"])
    offs = z["offsets"]
    eps = [z["err_probs"][offs[i]:offs[i + 1]] for i in range(len(offs) - 1)]
    sig = float(np.float32(0.01))
    kmax = z["counts"].max(axis=1)
    sel = np.nonzero((z["sig"] == sig) & (kmax <= MAXK))[0]
    pv, lnp, st = caller.snpcaller_batch([eps[i] for i in sel], z["counts"][sel], z["bonf"][sel], sig)
    names = z["names"][sel]
    bad = np.argwhere(st != z["status"][sel])
    assert len(bad) == 0, [(names[i], st[i].tolist(), z["status"][sel][i].tolist()) for i, _ in bad[:5]]
    assert np.array_equal(status_of(pv), z["status"][sel])
    assert_lnp_close(pv, want[sel], z["status"][sel], "grid")
    # the remaining cases one by one through the snpcaller() mirror
    for i in np.nonzero((z["sig"] != sig) & (kmax <= MAXK))[0]:
        got = caller.snpcaller(eps[i], z["counts"][i], int(z["bonf"][i]), float(z["sig"][i]))
        assert np.array_equal(status_of(got), z["status"][i]), z["names"][i]
        assert_lnp_close(got, want[i], z["status"][i], str(z["names"][i]))
    # alt counts beyond the largest register tile must fail loudly, not silently compute something else
    from lofreq_b200.capi import Lfb200Error
    for i in np.nonzero(kmax > MAXK)[0]:
        with pytest.raises(Lfb200Error):
            caller.snpcaller(eps[i], z["counts"][i], int(z["bonf"][i]), float(z["sig"][i]))
    kat = caller.snpcaller(np.full(10, 0.001), (1, 0, 0), 1, 1.0)
    assert abs(float(kat[0]) - 0.00995512) < 5e-9


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "batch_*.npz"))))
def test_batch_golden(caller, path):
    z = np.load(path)
    b = synth_np.generate(str(z["workload"]), int(z["c0"]), int(z["n_cols"]), with_baq=bool(z["with_baq"]))
    got = caller.call_columns(b, load_conf(z))
    want = {k: z[k] for k in ("alt_counts", "alt_raw_counts", "tested", "bonf_used", "called", "qual")}
    want["pvalues"] = ld_from_bytes(z["pvalue_ld"])
    want["bonf_subst"] = int(z["bonf_subst"])
    want["num_snv_tests"] = int(z["num_snv_tests"])
    compare_batch(got, want, os.path.basename(path))


@pytest.mark.parametrize("wl,c0,n,baq", [("C2", 31337, 4000, True), ("C3", 20000, 600, False),
                                        ("C4", 999000, 5000, True), ("C5", 4100, 400, False)])
def test_batch_vs_oracle(caller, port_oracle, wl, c0, n, baq):
    b = synth_np.generate(wl, c0, n, with_baq=baq)
    conf = default_conf(bonf_subst=1)
    want = port_oracle.call_columns(b, dict(conf))
    got = caller.call_columns(b, dict(conf))
    compare_batch(got, want, wl)
    assert got["n_sites"] >= int(want["called"].any(axis=1).sum())


def test_running_bonferroni_carries_over_batches(caller, port_oracle):
    """two consecutive batches with the conf carried over == one batch (lofreq_call.c:794-801)"""
    b_all = synth_np.generate("C2", 0, 3000)
    want = port_oracle.call_columns(b_all, default_conf())
    import lofreq_b200
    cf = lofreq_b200.varcall_conf()
    first = synth_np.generate("C2", 0, 1200)
    second = synth_np.generate("C2", 1200, 1800)
    g1 = caller.call_columns(first, cf)
    g2 = caller.call_columns(second, cf)
    assert np.array_equal(np.r_[g1["bonf_used"], g2["bonf_used"]], want["bonf_used"])
    assert np.array_equal(np.r_[g1["called"], g2["called"]], want["called"])
    assert np.array_equal(np.r_[g1["qual"], g2["qual"]], want["qual"])
    assert cf.bonf_subst == want["bonf_subst"] and cf.num_snv_tests == want["num_snv_tests"]


def _custom_batch(cols, pad=1, with_baq=True):
    """cols: list of dict(ref=, groups=[(bq, mq, baq) arrays per A,C,G,T], coverage=optional)"""
    col_off = [0]
    bq, mq, baq, nt, ref, cov = [], [], [], [], [], []
    for c in cols:
        n = 0
        for g in range(4):
            q = c["groups"][g]
            nt.append(len(q[0]))
            bq.extend(q[0]); mq.extend(q[1]); baq.extend(q[2])
            n += len(q[0])
        padn = (-n) % pad + c.get("gap", 0)
        bq.extend([0] * padn); mq.extend([0] * padn); baq.extend([0] * padn)
        col_off.append(col_off[-1] + n + padn)
        ref.append(ord(c["ref"]))
        cov.append(c.get("coverage", n))
    tail = 64
    nb = [c.get("num_bases", sum(len(c["groups"][g][0]) for g in range(4))) for c in cols]
    return dict(num_bases=np.array(nb, np.int32), col_off=np.array(col_off, np.int64), nt_cnt=np.array(nt, np.int32).reshape(-1, 4),
                ref_base=np.array(ref, np.uint8), coverage=np.array(cov, np.int32),
                bq=np.array(bq + [0] * tail, np.uint8), mq=np.array(mq + [0] * tail, np.uint8),
                baq=np.array(baq + [0] * tail, np.uint8) if with_baq else None, sq=None)


def test_edge_columns(caller, port_oracle):
    rng = np.random.default_rng(5)

    def grp(n, qlo=20, qhi=41, mqv=None, baqv=None):
        return (rng.integers(qlo, qhi, n), rng.integers(0, 256, n) if mqv is None else np.full(n, mqv),
                rng.integers(0, 94, n) if baqv is None else np.full(n, baqv))
    e = (np.zeros(0, int),) * 3
    cols = [
        dict(ref="A", groups=[e, e, e, e]),                                   # empty column
        dict(ref="N", groups=[grp(30), grp(3), e, e]),                        # ref N: skipped (lofreq_call.c:892)
        dict(ref="C", groups=[grp(2), grp(40), grp(1), grp(5)], gap=7),       # ragged, unaligned next column
        dict(ref="G", groups=[grp(9), e, grp(100), e], coverage=500),         # num_bases*2 < coverage: skipped
        dict(ref="T", groups=[grp(13), grp(1), grp(1), grp(333)], gap=3),
        dict(ref="G", groups=[grp(9), e, grp(100), e], coverage=300, num_bases=160),   # reads showing N count as bases
        dict(ref="G", groups=[grp(9), e, grp(100), e], coverage=300, num_bases=140),
        dict(ref="A", groups=[e, grp(50), e, e]),                             # no ref reads at all, K == N
        dict(ref="A", groups=[grp(60, mqv=255), grp(8, mqv=255), e, e]),      # mq unknown
        dict(ref="A", groups=[grp(60, mqv=0), grp(8, mqv=0), e, e]),          # mq 0 -> 0.5
        dict(ref="C", groups=[grp(4, baqv=255), grp(200, baqv=255), grp(3, baqv=255), e]),   # baq not available
        dict(ref="G", groups=[grp(5, 0, 6), grp(5, 0, 6), grp(90, 0, 12), grp(7, 0, 12)]),   # around min_bq = 6
        dict(ref="T", groups=[grp(1), e, e, e]),                              # N == 1, alt only
        dict(ref="T", groups=[e, e, e, grp(1)]),                              # N == 1, ref only
        dict(ref="A", groups=[grp(700), grp(300), grp(120), grp(15)]),        # tri-allelic, large counts
        dict(ref="C", groups=[grp(9, 60, 94), grp(2500, 60, 94, 60, 93), grp(1200, 60, 94, 60, 93), grp(1, 60, 94)]),
    ]
    for pad in (1, 16):
        b = _custom_batch(cols, pad=pad)
        for conf in (default_conf(), default_conf(min_cov=10), default_conf(flag=0), default_conf(def_alt_bq=-1),
                     default_conf(min_bq=3, min_alt_bq=20, min_jq=15, min_alt_jq=25, def_alt_jq=35)):
            want = port_oracle.call_columns(b, dict(conf))
            got = caller.call_columns(b, dict(conf))
            compare_batch(got, want, "edge pad=%d %s" % (pad, conf))
    empty = _custom_batch([])
    got = caller.call_columns(empty, default_conf())
    assert got["n_sites"] == 0 and got["num_snv_tests"] == 0 and got["bonf_subst"] == 1


def test_random_snpcaller_problems(caller, port_oracle):
    rng = np.random.default_rng(77)
    sig = float(np.float32(0.01))
    eps, cnts, bonfs = [], [], []
    for it in range(300):
        n = int(rng.integers(1, 2600))
        mode = rng.integers(0, 4)
        if mode == 0:
            q = rng.integers(20, 41, n)
        elif mode == 1:
            q = rng.integers(6, 60, n)
        elif mode == 2:
            q = np.full(n, int(rng.integers(10, 45)))
        else:
            q = np.where(rng.random(n) < 0.3, 3, rng.integers(20, 41, n))
        ep = np.sort(10.0 ** (-q / 10.0))
        k1 = int(rng.integers(1, n + 1)) if rng.random() < 0.7 else int(rng.integers(1, min(n, 20) + 1))
        k2 = int(rng.integers(0, min(k1, n - k1) + 1)) if rng.random() < 0.7 else 0
        k3 = int(rng.integers(0, min(5, n - k1 - k2) + 1))
        c = [k1, k2, k3]
        rng.shuffle(c)
        eps.append(ep); cnts.append(c); bonfs.append(int(rng.integers(1, 10 ** 7)))
    pv, lnp, st = caller.snpcaller_batch(eps, cnts, bonfs, sig)
    want = np.array([port_oracle.snpcaller(e, c, b, sig) for e, c, b in zip(eps, cnts, bonfs)])
    wst = status_of(want)
    bad = np.argwhere(st != wst)
    assert len(bad) == 0, [(len(eps[i]), cnts[i], bonfs[i], st[i].tolist(), wst[i].tolist()) for i, _ in bad[:5]]
    assert_lnp_close(pv, want, wst, "random problems")


@pytest.mark.parametrize("wl,c0,n,baq", [("C2", 0, 3000, True), ("C3", 123456, 300, True), ("C4", 5, 3000, False),
                                        ("C5", 1000, 500, True)])
def test_synth_matches_numpy(wl, c0, n, baq):
    import torch
    from lofreq_b200 import synth
    t = synth.generate_device(wl, c0, n, with_baq=baq)
    torch.cuda.synchronize()
    b = synth_np.generate(wl, c0, n, with_baq=baq)
    assert np.array_equal(t["col_off"].cpu().numpy(), b["col_off"])
    assert np.array_equal(t["nt_cnt"].cpu().numpy(), b["nt_cnt"])
    assert np.array_equal(t["ref_base"].cpu().numpy(), b["ref_base"])
    tot = int(b["col_off"][-1])
    assert np.array_equal(t["bq"].cpu().numpy()[:tot], b["bq"])
    assert np.array_equal(t["mq"].cpu().numpy()[:tot], b["mq"])
    if baq:
        assert np.array_equal(t["baq"].cpu().numpy()[:tot], b["baq"])


def test_full_size_c2_properties(caller, port_oracle):
    """BASELINE.json configs[1]: 1M columns at depth 500, Q30, generated on the device.
    Size-independent properties + a sampled comparison against the oracle."""
    import lofreq_b200
    from lofreq_b200 import synth
    n = 1_000_000
    t = synth.generate_device("C2", 0, n)
    db = caller.device_batch(t)
    cf = lofreq_b200.varcall_conf()
    caller.screen(db, cf)
    n_tested = caller.ntested()
    caller.test(cf)
    sites, sm = caller.sites(cf, n)
    assert sm.n_tested == n_tested and sm.num_snv_tests == 3 * n_tested and cf.bonf_subst == 3 * n_tested
    # the same columns through the host entry point on a slice, against the oracle
    sel0, sel1 = 250_000, 262_000
    b = synth_np.generate("C2", sel0, sel1 - sel0)
    conf = default_conf(bonf_subst=1)
    want = port_oracle.call_columns(b, dict(conf))
    got = caller.call_columns(b, dict(conf))
    compare_batch(got, want, "C2 slice")
    # sites of the full run restricted to the slice must be the called columns of the slice, with the
    # full-run Bonferroni factors being larger (running factor) -> subset relation
    s = sites
    assert np.all(np.diff(s["col"]) > 0)                      # sorted, unique
    in_slice = s[(s["col"] >= sel0) & (s["col"] < sel1)]
    full_called = set(int(c) - sel0 for c, cl in zip(in_slice["col"], in_slice["called"]) if cl.any())
    slice_called = set(np.nonzero(want["called"].any(axis=1))[0].tolist())
    assert full_called <= slice_called
    # idempotence: a second pass over the same resident batch gives identical sites
    cf2 = lofreq_b200.varcall_conf()
    caller.screen(db, cf2)
    caller.test(cf2)
    sites2, sm2 = caller.sites(cf2, n)
    s2 = sites2
    assert sm2.n_sites == sm.n_sites and np.array_equal(s["col"], s2["col"]) and np.array_equal(s["lnp"], s2["lnp"])
    assert np.array_equal(s["qual"], s2["qual"])
    # 1 % of the columns are variant sites, nearly all of them significant
    assert 0.005 * n < sm.n_sites < 0.02 * n


def test_graph_replay_gives_the_same_sites():
    """The launches of a phase are replayed as a CUDA graph once the same sequence has been queued twice in a row
    (host_api.cpp: run_graphed).  The same resident batch through plain launches (legacy stream: never captured), through
    the capture and through replays; a different configuration in between must not be served from the cache."""
    import torch
    import lofreq_b200
    from lofreq_b200 import synth, capi
    c = lofreq_b200.Caller()
    n = 60_000
    t = synth.generate_device("C2", 0, n)
    db = c.device_batch(t)

    def run(stream, **over):
        cf = lofreq_b200.varcall_conf(**over)
        c.screen(db, cf, stream)
        c.test(cf, stream)
        sites, sm = c.sites(cf, n, stream)
        return sites.copy(), int(sm.n_sites), int(sm.n_tested), int(cf.bonf_subst)

    want = run(None)
    want_strict = run(None, sig=1e-4)
    assert want_strict[1] <= want[1]
    st = torch.cuda.Stream()
    sp = st.cuda_stream
    before = c.lib.lfb200_graph_replays(c._ctx)
    for i in range(5):
        got = run(sp)
        assert got[1:] == want[1:], (i, got[1:], want[1:])
        for f in ("col", "bonf", "lnp", "qual", "called", "status", "alt_count"):
            assert np.array_equal(got[0][f], want[0][f]), (i, f)
    assert c.lib.lfb200_graph_replays(c._ctx) - before >= 6          # both phases, from the second repeat on
    got = run(sp, sig=1e-4)                                          # another configuration: plain launches again
    assert got[1:] == want_strict[1:]
    for f in ("col", "bonf", "lnp", "qual", "called"):
        assert np.array_equal(got[0][f], want_strict[0][f]), f
    got = run(sp)
    assert got[1:] == want[1:] and np.array_equal(got[0]["lnp"], want[0]["lnp"])
    torch.cuda.synchronize()
    c.close()


def test_low_mapping_qualities_under_a_large_factor(caller, port_oracle):
    """The first stage of the early exit looks at the base qualities alone (a lower bound of every read's merged error
    probability); with poor mapping qualities the merged probabilities are far above it, so columns that this stage lets
    through are still ruled out (or called) exactly by the later stages.  A large fixed factor makes the prune bite.
    Also the division-free step parameters of k_dp at mq = 0 (probability 0.5), mq unknown (255) and BAQ 0..60."""
    rng = np.random.default_rng(2024)
    e = (np.zeros(0, int),) * 3
    cols = []
    for i in range(6000):
        n = int(rng.integers(30, 400))
        kind = rng.random()
        k = int(rng.integers(1, 4)) if kind < 0.7 else int(rng.integers(4, 9)) if kind < 0.9 else int(rng.integers(9, min(n // 2, 120)))
        mq_pool = rng.choice([0, 1, 3, 10, 20, 30, 60, 255], size=n, p=[.1, .05, .05, .1, .1, .1, .45, .05])

        def grp(m, lo, hi, off):
            return (rng.integers(lo, hi, m), mq_pool[off:off + m], rng.integers(0, 61, m))
        alt_q = (25, 42) if rng.random() < 0.5 else (8, 25)
        groups = [grp(n - k, 20, 42, 0), grp(k, alt_q[0], alt_q[1], n - k), e, e]
        rng.shuffle(groups)
        ref = "ACGT"[max(range(4), key=lambda g: len(groups[g][0]))]
        cols.append(dict(ref=ref, groups=groups, gap=int(rng.integers(0, 9))))
    b = _custom_batch(cols, pad=1)
    for conf in (default_conf(bonf_subst=3_000_000, bonf_dynamic=0), default_conf(bonf_subst=1),
                 default_conf(bonf_subst=3_000_000, bonf_dynamic=0, flag=2)):       # 2 = USE_MQ only
        want = port_oracle.call_columns(b, dict(conf))
        got = caller.call_columns(b, dict(conf))
        compare_batch(got, want, "low mq %s" % conf)
        assert 0 < want["called"].any(axis=1).sum() < len(cols)


def _slice_out(d, lo, hi):
    return {k: (v[lo:hi] if isinstance(v, np.ndarray) else v) for k, v in d.items()}


def _against_pool(caller, wl, c0, n, baq, chunk):
    """`n` consecutive columns of a BASELINE.json workload through lfb200_call_columns (one call, host buffers made from
    the device generator), site for site against the CPU checker on all host cores (tests/ref_pool.py): the compiled,
    unmodified reference when oracle/_ref travelled to this box.  Returns (kind, worst d ln p, QUAL +-1 hits)."""
    import ref_pool
    from lofreq_b200 import synth
    t = synth.generate_device(wl, c0, n, with_baq=baq)
    tot = t["total_bytes"]
    b = dict(col_off=t["col_off"].cpu().numpy(), nt_cnt=t["nt_cnt"].cpu().numpy(), ref_base=t["ref_base"].cpu().numpy(),
             bq=t["bq"].cpu().numpy(), mq=t["mq"].cpu().numpy(), baq=t["baq"].cpu().numpy() if baq else None, sq=None)
    assert int(b["col_off"][-1]) == tot
    del t
    conf = default_conf()
    got = caller.call_columns(b, dict(conf))
    res, kind = ref_pool.run_chunks(wl, c0, n, got["tested"], conf, chunk=chunk, with_baq=baq)
    worst, hits, n_tested = 0.0, 0, 0
    for lo, hi, want in res:
        want = dict(want)
        want.pop("bonf_subst"); want.pop("num_snv_tests")
        what = "%s[%d:%d]" % (wl, c0 + lo, c0 + hi)
        worst = max(worst, compare_batch(_slice_out(got, lo, hi), want, what))
        hits += QUAL_TOLERANCE_HITS[what]
        n_tested += int(want["tested"].sum())
    assert got["num_snv_tests"] == 3 * n_tested and got["bonf_subst"] == 3 * n_tested
    print("%s: %d columns vs %s on %d cores: %d tested, %d sites, %d called alleles, worst d(ln p) %.2e, QUAL +-1 hits %d, kernels %s"
          % (wl, n, kind, ref_pool.host_cores(), n_tested, got["n_sites"], int(got["called"].sum()), worst, hits, got["job_counts"]))
    return kind, worst, hits


def test_full_size_c2_vs_reference(caller):
    """BASELINE.json configs[1] at its full size — 1 M columns, depth 500, Q30 — column for column and site for site
    against the compiled reference (the C restatement where oracle/_ref did not travel).  QUAL bit-exact: the +-1
    allowance of compare_batch must not fire once."""
    kind, worst, hits = _against_pool(caller, "C2", 0, 1_000_000, False, 25_000)
    assert hits == 0


@pytest.mark.parametrize("wl,c0,n,baq,chunk", [("C3", 3_000_000, 50_000, False, 2_500), ("C5", 7_000_000, 50_000, False, 1_250),
                                              ("C4", 50_000_000, 200_000, True, 10_000)])
def test_config_samples_vs_reference(caller, wl, c0, n, baq, chunk):
    """C3 (depth 2000, Q20-40), C5 (depth 50-10000) and C4 (depth 300, with BAQ: the general merge) samples large enough
    that the K > 256 kernels see hundreds of real columns; same bar as the C2 test."""
    kind, worst, hits = _against_pool(caller, wl, c0, n, baq, chunk)
    assert hits == 0


def test_two_shards_equal_one(caller):
    """region shards with the tested-count exchange (lofreq_b200/shard.py) == one batch; both shards on
    this GPU, one context each, driven exactly like bench.py drives one rank per GPU"""
    import lofreq_b200
    from lofreq_b200 import shard, synth
    n, world = 200_000, 2
    whole = synth.generate_device("C2", 0, n)
    cf = lofreq_b200.varcall_conf()
    caller.screen(caller.device_batch(whole), cf)
    caller.test(cf)
    want, want_sm = caller.sites(cf, n)
    ctxs = [lofreq_b200.Caller(0) for _ in range(world)]
    parts, counts = [], []
    for r in range(world):
        lo, hi = shard.shard_range(n, r, world)
        t = synth.generate_device("C2", lo, hi - lo)
        parts.append((lo, hi, t))
        ctxs[r].screen(ctxs[r].device_batch(t), lofreq_b200.varcall_conf())
        counts.append(ctxs[r].ntested())
    got = []
    import ctypes as C
    import torch
    from lofreq_b200 import capi
    for r in range(world):
        start = shard.bonf_start_for_rank(counts, r)
        if r == 0:
            cfr = lofreq_b200.varcall_conf(bonf_subst=start)
            ctxs[r].test(cfr)
        else:
            # the running factor handed over in device memory (what the NCCL exchange of bench.py produces)
            cfr = lofreq_b200.varcall_conf()
            dst = torch.zeros(1, dtype=torch.int64, device="cuda:0")
            capi.check(ctxs[0].lib.lfb200_ntested_copy_device(ctxs[0]._ctx, None, C.c_void_p(dst.data_ptr())))
            assert int(dst.item()) == counts[0]
            start_dev = torch.tensor([start], dtype=torch.int64, device="cuda:0")
            capi.check(ctxs[r].lib.lfb200_test_device_from(ctxs[r]._ctx, C.byref(cfr), None, C.c_void_p(start_dev.data_ptr())))
        s, sm = ctxs[r].sites(cfr, n)
        assert cfr.bonf_subst == shard.final_counters(counts[: r + 1])[0]
        s = s.copy()
        s["col"] += parts[r][0]
        got.append(s)
    got = np.concatenate(got)
    assert shard.final_counters(counts) == (want_sm.bonf_subst_final, want_sm.num_snv_tests)
    assert len(got) == len(want)
    for k in ("col", "bonf", "alt_count", "alt_raw_count", "qual", "status", "called"):
        assert np.array_equal(got[k], want[k]), k
    assert np.array_equal(got["lnp"], want["lnp"])
    for c in ctxs:
        c.close()


def test_column_builder_callback_surface(caller, port_oracle):
    """one column at a time through lfb200_builder_add_column (the call_vars() replacement), several
    flushes, sites reported in input order; equals the oracle's per-column loop"""
    import lofreq_b200
    from lofreq_b200.snpcaller import ColumnBuilder
    b = synth_np.generate("C4", 4242, 1300, with_baq=True)
    want = port_oracle.call_columns(b, default_conf())
    got = []
    cf = lofreq_b200.varcall_conf()
    bld = ColumnBuilder(caller, cf, 256, got.append)
    for c in range(1300):
        lo = int(b["col_off"][c])
        groups_bq, groups_mq, groups_baq = [], [], []
        for g in range(4):
            k = int(b["nt_cnt"][c, g])
            groups_bq.append(b["bq"][lo:lo + k].astype(np.int32))
            groups_mq.append(b["mq"][lo:lo + k].astype(np.int32))
            groups_baq.append(b["baq"][lo:lo + k].astype(np.int32))
            lo += k
        bld.add_column(1000 + c, chr(b["ref_base"][c]), int(b["nt_cnt"][c].sum()), int(b["nt_cnt"][c].sum()), groups_bq,
                       groups_mq, groups_baq)
    bld.flush()
    assert bld.conf.bonf_subst == want["bonf_subst"] and bld.conf.num_snv_tests == want["num_snv_tests"]
    tags = [s["tag"] for s in got]
    assert tags == sorted(tags)
    called_cols = np.nonzero(want["called"].any(axis=1))[0]
    got_called = [s for s in got if any(s["called"])]
    assert [s["tag"] - 1000 for s in got_called] == called_cols.tolist()
    for s in got_called:
        c = s["tag"] - 1000
        assert s["qual"] == want["qual"][c].tolist() and s["called"] == want["called"][c].tolist()
        assert s["bonf"] == want["bonf_used"][c] and s["alt_count"] == want["alt_counts"][c].tolist()
    bld.close()


def test_strict_fenv_mode_gives_the_same_sentinels():
    """host finishing with the literal feclearexcept/fetestexcept sequence of the reference
    (LFB200_STRICT_FENV=1) == the default result-based underflow test, on the golden grid"""
    import subprocess
    import sys
    code = (
        "import os, sys, numpy as np\n"
        "sys.path.insert(0, %r)\n"
        "import lofreq_b200\n"
        "z = np.load(%r)\n"
        "offs = z['offsets']; sig = float(np.float32(0.01))\n"
        "sel = np.nonzero((z['sig'] == sig) & (z['counts'].max(axis=1) <= 2048))[0]\n"
        "c = lofreq_b200.Caller(0)\n"
        "pv, lnp, st = c.snpcaller_batch([z['err_probs'][offs[i]:offs[i+1]] for i in sel], z['counts'][sel], z['bonf'][sel], sig)\n"
        "assert np.array_equal(st, z['status'][sel])\n"
        "sys.stdout.write(pv.tobytes().hex())\n" % (os.path.dirname(GOLD + '/..').rsplit('/tests', 1)[0], os.path.join(GOLD, 'snpcaller_grid.npz')))
    outs = []
    for strict in ("", "1"):
        env = dict(os.environ)
        env.pop("LFB200_STRICT_FENV", None)
        if strict:
            env["LFB200_STRICT_FENV"] = "1"
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(r.stdout)
    assert outs[0] == outs[1] and len(outs[0]) > 1000


def test_deep_columns_cta_per_column_kernel(caller, port_oracle):
    """alt counts above 2048 (deep amplicon columns) run in the CTA-per-column kernel k_heavy_xl"""
    rng = np.random.default_rng(9)

    def grp(n, qlo=20, qhi=41):
        return (rng.integers(qlo, qhi, n), np.full(n, 60), rng.integers(30, 61, n))
    e = (np.zeros(0, int),) * 3
    cols = [
        dict(ref="A", groups=[grp(5000), grp(4000), grp(7), e]),                  # K = 4000
        dict(ref="C", groups=[grp(2500), grp(3000), grp(2100), grp(12)]),         # tri-allelic, two counts above 2048
        dict(ref="G", groups=[grp(3), grp(2), grp(30), grp(6000)]),               # nearly all reads alt
        dict(ref="T", groups=[grp(40), grp(3000, 6, 12), e, grp(9000, 6, 12)]),   # low qualities: large lambda, no tilt
        dict(ref="A", groups=[grp(600), grp(350), e, e]),                         # an ordinary heavy column in between
    ]
    b = _custom_batch(cols, pad=16)
    for conf in (default_conf(), default_conf(flag=0)):
        want = port_oracle.call_columns(b, dict(conf))
        got = caller.call_columns(b, dict(conf))
        compare_batch(got, want, "deep %s" % conf)
    # and through the snpcaller() mirror
    ep = np.sort(10.0 ** (-rng.integers(20, 41, 7000) / 10.0))
    for counts in ((3000, 0, 0), (2500, 2200, 4)):
        want = port_oracle.snpcaller(ep, counts, 30000, float(np.float32(0.01)))
        got = caller.snpcaller(ep, counts, 30000, float(np.float32(0.01)))
        assert np.array_equal(status_of(got), status_of(want))
        assert_lnp_close(got, want, status_of(want), "deep snpcaller")


def test_alt_count_beyond_every_kernel_is_reported_not_dropped(caller, port_oracle):
    """a column whose largest alt count exceeds what the kernels take (16384) comes back as a site with status
    LFB200_ST_UNSUPPORTED; the other columns of the batch are computed and the Bonferroni state advances as if nothing
    had happened (it used to fail the whole batch)"""
    rng = np.random.default_rng(12)

    def grp(n, qlo=20, qhi=41):
        return (rng.integers(qlo, qhi, n), np.full(n, 60), rng.integers(30, 61, n))
    e = (np.zeros(0, int),) * 3
    cols = [dict(ref="A", groups=[grp(400), grp(60), e, e]),
            dict(ref="C", groups=[grp(17000), grp(300), grp(4), e]),      # K = 17000 > 16384
            dict(ref="G", groups=[grp(3), e, grp(500), grp(45)]),
            dict(ref="T", groups=[e, e, e, grp(80)])]
    b = _custom_batch(cols, pad=16)
    want = port_oracle.call_columns(b, default_conf())
    got = caller.call_columns(b, default_conf())
    assert got["n_unsupported"] == 1
    for k in ("alt_counts", "alt_raw_counts", "tested", "bonf_used"):
        assert np.array_equal(got[k], want[k]), k
    assert got["bonf_subst"] == want["bonf_subst"] and got["num_snv_tests"] == want["num_snv_tests"]
    ok = np.array([0, 2, 3])
    assert np.array_equal(got["called"][ok], want["called"][ok]) and np.array_equal(got["qual"][ok], want["qual"][ok])
    assert_lnp_close(got["pvalues"][ok], want["pvalues"][ok], status_of(want["pvalues"][ok]), "beside the unsupported column")
    assert got["status"][1].tolist() == [3, 3, 1] and not got["called"][1].any()      # the third allele has no reads: LDBL_MAX as ever
    s = got["sites"]
    assert (s["flags"][s["col"] == 1] & 2).all()


def test_packed_kernel_classes(caller, port_oracle):
    """8 < K <= 2048: k_dp, 4/8/16/32 lanes per column and 8..64 cells per lane, columns of one warp in lock step.
    Every class, tilted and untilted rows, secondary alleles read off a tilted row, ragged depths inside
    one warp, neutral steps for filtered reads, and the hand-over to the fallback list."""
    rng = np.random.default_rng(77)

    def grp(n, qlo=20, qhi=41, mqv=60, baqv=None):
        return (rng.integers(qlo, qhi, n), np.full(n, mqv), rng.integers(30, 61, n) if baqv is None else np.full(n, baqv))
    e = (np.zeros(0, int),) * 3
    cols = []
    # every group width at several depths (each depth bin has its own job list), ragged inside a bin
    for K in (9, 17, 32, 33, 50, 64, 65, 100, 128, 129, 200, 256, 257, 400, 512, 513, 1000, 1024, 1025, 1500, 2048):
        for depth in (K + 1, 2 * K + 37, 600, 1100, 3000):
            if depth <= K:
                continue
            cols.append(dict(ref="A", groups=[grp(depth - K), grp(K), e, e]))
    # high qualities: far tails -> tilted rows; second and third allele read off the same row
    for K, c2, c3 in ((40, 3, 0), (90, 12, 1), (130, 60, 9), (250, 120, 119), (256, 1, 1), (70, 70, 2)):
        cols.append(dict(ref="C", groups=[grp(c2, 35, 41), grp(700, 35, 41), grp(K, 35, 41), grp(c3, 35, 41)]))
    # nearly all reads alt (a low cell of the tilted row underflows -> fallback list), K == N
    cols.append(dict(ref="G", groups=[grp(250, 38, 41), e, grp(3, 38, 41), grp(2, 38, 41)]))
    cols.append(dict(ref="G", groups=[grp(200, 38, 41), e, e, e]))
    cols.append(dict(ref="T", groups=[grp(120, 30, 41), grp(2, 30, 41), grp(1, 30, 41), e]))
    # filtered reads in between (bq < min_bq): neutral steps; mq 0 -> p >= 0.5, odds >= 1
    cols.append(dict(ref="A", groups=[grp(400, 0, 41), grp(60, 0, 41), grp(5, 0, 12), e]))
    cols.append(dict(ref="A", groups=[grp(300, 20, 41, mqv=0), grp(40, 20, 41, mqv=0), e, e]))
    # baq 0 -> merged probability exactly 1: 1/q = 2^52 -> parameters above 2^20 -> fallback list
    cols.append(dict(ref="C", groups=[grp(20, baqv=0), grp(300), grp(30), e]))
    # deeper than the packed form holds (n > 16384): per-column kernel; and a deep one it does hold
    cols.append(dict(ref="T", groups=[grp(100), e, e, grp(17000)]))
    cols.append(dict(ref="T", groups=[grp(100), e, grp(3), grp(9000)]))
    order = rng.permutation(len(cols))
    cols = [cols[i] for i in order]
    for pad in (1, 16):
        b = _custom_batch(cols, pad=pad)
        for conf in (default_conf(), default_conf(flag=0), default_conf(bonf_dynamic=0, bonf_subst=3000000),
                     default_conf(min_bq=3, min_alt_bq=20, min_jq=15, min_alt_jq=25, def_alt_jq=35), default_conf(def_alt_bq=25)):
            want = port_oracle.call_columns(b, dict(conf))
            got = caller.call_columns(b, dict(conf))
            compare_batch(got, want, "packed pad=%d %s" % (pad, conf))
            jc = got["job_counts"]
            assert jc["packed"] >= 60, jc                      # the packed kernel really took them
            assert jc["fallback"] < jc["packed"] // 4, jc
    # median override (def_alt_bq = -1): the histogram of the reference-base qualities is built inside k_dp
    got = caller.call_columns(b, default_conf(def_alt_bq=-1))
    compare_batch(got, port_oracle.call_columns(b, default_conf(def_alt_bq=-1)), "packed median")
    assert got["job_counts"]["packed"] >= 60


def test_dp_job_list_overflow(port_oracle):
    """more columns of one (class, depth bin) than its job list holds (an eighth of the batch, at least 4096): the excess
    goes to the class's unbinned list and is computed all the same"""
    import lofreq_b200
    caller = lofreq_b200.Caller(0)
    rng = np.random.default_rng(3)
    e = (np.zeros(0, int),) * 3
    cols = []
    for i in range(5000):
        n, K = 100 + int(rng.integers(0, 28)), int(rng.integers(9, 33))
        cols.append(dict(ref="A", groups=[(rng.integers(25, 41, n - K), np.full(n - K, 60), np.full(n - K, 40)),
                                          (rng.integers(25, 41, K), np.full(K, 60), np.full(K, 40)), e, e]))
    b = _custom_batch(cols, pad=16)
    want = port_oracle.call_columns(b, default_conf())
    got = caller.call_columns(b, default_conf())
    compare_batch(got, want, "job list overflow")
    jc = got["job_counts"]
    assert jc["packed"] == 5000 and int(want["called"].any(axis=1).sum()) > 4500, jc
    caller.close()


def test_poissbin_rows_golden(caller):
    """poissbin() (snpcaller.h:93-96): the whole row, the LDBL clamp of *pvalue and — through the row — the read at
    which the reference's early exit fired, against rows produced by the compiled reference"""
    z = np.load(os.path.join(GOLD, "poissbin_rows.npz"))
    n = len(z["names"])
    eps = [z["err_probs"][z["offsets"][i]:z["offsets"][i + 1]] for i in range(n)]
    for i, name in enumerate(z["names"]):
        want = z["rows"][z["row_offsets"][i]:z["row_offsets"][i + 1]]
        pv, row, n_end = caller.poissbin_batch([eps[i]], [z["k"][i]], [z["bonf"][i]], float(z["sig"][i]))[0]
        assert row.shape == want.shape, name
        err = np.abs(row - want) / np.maximum(np.abs(want), 1.0)
        assert err.max() <= LNP_RTOL, (name, float(err.max()), int(np.argmax(err)), n_end)
        want_pv = ld_from_bytes(z["pvalue_ld"][i])
        assert status_of(np.array([pv])) == status_of(np.array([want_pv])), name
        if status_of(np.array([want_pv]))[0] == 0:
            assert abs(float(np.log(pv)) - float(np.log(want_pv))) <= LNP_RTOL * max(abs(float(np.log(want_pv))), 1.0), name
        if i % 6 == 0:                                  # the link-compatible symbol (malloc'ed row)
            pv1, row1 = caller.poissbin(eps[i], int(z["k"][i]), int(z["bonf"][i]), float(z["sig"][i]))
            assert np.array_equal(row1, row) and pv1 == pv, name
    # several problems in one call
    sel = [i for i in range(n) if float(z["sig"][i]) == float(np.float32(0.01))]
    many = caller.poissbin_batch([eps[i] for i in sel], z["k"][sel], z["bonf"][sel], float(np.float32(0.01)))
    for (pv, row, _), i in zip(many, sel):
        want = z["rows"][z["row_offsets"][i]:z["row_offsets"][i + 1]]
        assert (np.abs(row - want) / np.maximum(np.abs(want), 1.0)).max() <= LNP_RTOL, z["names"][i]


def test_plp_to_errprobs(caller, port_oracle):
    """plp_to_errprobs() (snpcaller.h:72-75): error probabilities in pileup order bit-exact, alt bases and both counts"""
    for wl, c0, ncol, baq in (("C2", 77, 300, True), ("C3", 5000, 60, False), ("C5", 900, 80, True)):
        b = synth_np.generate(wl, c0, ncol, with_baq=baq)
        for conf in (default_conf(), default_conf(flag=0), default_conf(def_alt_bq=-1), default_conf(def_alt_bq=30),
                     default_conf(min_bq=3, min_alt_bq=20, min_jq=15, min_alt_jq=25, def_alt_jq=35)):
            got = caller.batch_errprobs(b, dict(conf))
            for c in range(ncol):
                if chr(int(b["ref_base"][c])) not in "ACGT":
                    continue
                ep, ab, ac, ar = port_oracle.column_errprobs(b, c, dict(conf))
                gep, gab, gac, gar = got[c]
                assert np.array_equal(gep, ep), (wl, c, conf)
                assert np.array_equal(gab, ab) and np.array_equal(gac, ac) and np.array_equal(gar, ar), (wl, c, conf)
    # the link-compatible symbol on int arrays per nt4 (plp_col_t varrays), -1 = quality not available
    rng = np.random.default_rng(11)
    bqs = [rng.integers(0, 42, n) for n in (40, 3, 0, 7)]
    mqs = [rng.integers(0, 61, len(x)) for x in bqs]
    mqs[0][:3] = 255
    baqs = [rng.integers(-1, 60, len(x)) for x in bqs]
    cols = [dict(ref="A", groups=[(bqs[g], np.where(mqs[g] == 255, 255, mqs[g]), np.where(baqs[g] < 0, 255, baqs[g])) for g in range(4)])]
    bb = _custom_batch(cols)
    want = port_oracle.column_errprobs(bb, 0, default_conf())
    got = caller.plp_to_errprobs("A", 50, bqs, mqs, baqs, None, default_conf())
    assert np.array_equal(got[0], want[0]) and all(np.array_equal(g, w) for g, w in zip(got[1:], want[1:]))


def test_indel_tests_golden(caller):
    """the indel path: quality bytes -> merged probabilities -> snpcaller with (event count, 0, 0),
    against p-values the reference's plp_to_ins_errprobs / plp_to_del_errprobs + snpcaller produced"""
    z = np.load(os.path.join(GOLD, "indel_tests.npz"))
    sig = float(z["sig"])
    for flag in sorted(set(z["flag"].tolist())):
        sel = np.flatnonzero(z["flag"] == flag)
        off = np.zeros(len(sel) + 1, np.int64)
        parts = {k: [] for k in ("iq", "mq", "aq", "sq")}
        for j, i in enumerate(sel):
            lo, hi = int(z["read_off"][i]), int(z["read_off"][i + 1])
            off[j + 1] = off[j] + hi - lo
            for k in parts:
                parts[k].append(z[k][lo:hi])
        planes = {k: np.concatenate(v) for k, v in parts.items()}
        got = caller.indel_tests(off, planes["iq"], planes["mq"], planes["aq"], planes["sq"], z["event_count"][sel], z["bonf"][sel],
                                 default_conf(flag=int(flag), sig=sig))
        want = ld_from_bytes(z["pvalue_ld"][sel])
        st = status_of(want)
        assert np.array_equal(got["status"], st), (flag, np.argwhere(got["status"] != st)[:5].tolist())
        assert_lnp_close(got["pvalues"], want, st, "indel flag=%d" % flag)
        with np.errstate(all="ignore"):
            want_called = (want * z["bonf"][sel].astype(np.longdouble) < np.longdouble(sig)).astype(np.uint8)
        assert np.array_equal(got["called"], want_called), flag
        with np.errstate(all="ignore"):
            want_q = np.where(want_called == 1, (-10.0 * np.log10(want)).astype(np.int64), -1)
        assert np.all(np.abs(got["qual"] - want_q) <= (np.abs(-10.0 * np.log10(want).astype(np.float64) % 1.0) < 1e-6)), flag


def test_binom_golden(caller):
    """binom() (binom.c:52-93 -> cdflib cdfbin): status codes exact, cdf and sf within 1e-10 relative of the compiled
    reference — through the batched entry point and through the link-compatible single call"""
    import ctypes as C
    z = np.load(os.path.join(GOLD, "binom_grid.npz"))
    st, p, q = caller.binom_batch(z["num_trials"], z["num_success"], z["prob_success"])
    assert np.array_equal(st, z["status"])
    ok = st == 0
    assert np.all(np.isnan(p[~ok])) and np.all(np.isnan(q[~ok]))       # untouched where cdfbin refuses
    for got, want in ((p[ok], z["cdf"][ok]), (q[ok], z["sf"][ok])):
        big = want >= 1e-290
        assert np.all(np.abs(got[big] - want[big]) <= 1e-10 * want[big])
        assert np.all(got[~big] < 1e-280)
    assert caller.binom_batch([], [], [])[0].shape == (0,)
    lib = caller.lib
    for i in (0, 1, 2, 8, 9, 10, 11, 12, 14, 40, 100, 900):
        a, b = C.c_double(-7.0), C.c_double(-7.0)
        rc = lib.lfb200_binom(C.byref(a), C.byref(b), int(z["num_trials"][i]), int(z["num_success"][i]), float(z["prob_success"][i]))
        assert rc == int(z["status"][i])
        if rc == 0:
            assert a.value == p[i] and b.value == q[i]
        else:
            assert a.value == -7.0 and b.value == -7.0
    a = C.c_double(-7.0)
    assert lib.lfb200_binom(C.byref(a), None, 500, 3, 0.05) == 0 and abs(a.value - 2.46755e-08) < 1e-12   # SURVEY §8c probe value


def test_pinned_planes_read_in_place(caller, port_oracle):
    """host plane mode 1 (lfb200_set_host_planes): the kernels read pinned quality planes in place over PCIe.
    Same results as the oracle, with unaligned/ragged columns, with and without baq, and for pageable planes
    (which fall back to the copy)."""
    import torch
    from lofreq_b200 import capi
    for wl, c0, n, baq in (("C2", 777, 3000, False), ("C5", 99, 150, True)):
        b = synth_np.generate(wl, c0, n, with_baq=baq)
        want = port_oracle.call_columns(b, default_conf())
        pinned = dict(b)
        hold = []
        for k in ("bq", "mq", "baq"):
            if b.get(k) is not None:
                t = torch.from_numpy(np.ascontiguousarray(b[k], np.uint8)).pin_memory()
                hold.append(t)
                pinned[k] = t.numpy()
        capi.check(caller.lib.lfb200_set_host_planes(caller._ctx, 1))
        try:
            got = caller.call_columns(pinned, default_conf())
            got_pageable = caller.call_columns(b, default_conf())
        finally:
            capi.check(caller.lib.lfb200_set_host_planes(caller._ctx, 0))
        compare_batch(got, want, wl + " pinned in place")
        compare_batch(got_pageable, want, wl + " pageable, mode 1")
