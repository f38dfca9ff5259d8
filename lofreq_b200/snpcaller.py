"""Host-side mirror of the reference's per-column interface (snpcaller.h:72-102,
lofreq_call.c:734-935), over the C ABI.  Names and argument meaning follow the reference:

  varcall_conf(**over)                       init_varcall_conf (snpcaller.c:626-651)
  Caller.snpcaller(err_probs, counts, bonf, sig)          snpcaller() (snpcaller.h:97-102)
  Caller.call_columns(batch, conf)           one call_vars() per column (lofreq_call.c:886-935), batched

Everything is computed on the GPU; a missing library or device raises."""
import ctypes as C

import numpy as np

from . import capi
from .capi import ST_VALUE, ST_LDBLMAX, ST_LDBLMIN  # noqa: F401

LDBL_MAX = np.finfo(np.longdouble).max
LDBL_MIN = np.finfo(np.longdouble).tiny


def varcall_conf(**over):
    lib = capi.load()
    c = capi.Conf()
    lib.lfb200_init_conf(C.byref(c))
    for k, v in over.items():
        if not hasattr(c, k):
            raise KeyError(k)
        setattr(c, k, v)
    return c


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def conf_from(conf):
    if conf is None:
        return varcall_conf()
    if isinstance(conf, capi.Conf):
        return conf
    return varcall_conf(**conf)


class Caller:
    """Owns one lfb200_ctx on one GPU."""

    def __init__(self, device=0):
        self.lib = capi.load()
        self._ctx = C.c_void_p()
        capi.check(self.lib.lfb200_create(C.byref(self._ctx), int(device)))
        self.device = int(device)

    def close(self):
        if self._ctx:
            self.lib.lfb200_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- snpcaller(): one problem or many ------------------------------------------------
    def snpcaller(self, err_probs, noncons_counts, bonf_factor, sig_level):
        pv, _, _ = self.snpcaller_batch([np.asarray(err_probs, np.float64)], [noncons_counts], [bonf_factor], sig_level)
        return pv[0]

    def snpcaller_batch(self, err_probs_list, counts, bonf, sig_level):
        n = len(err_probs_list)
        off = np.zeros(n + 1, np.int64)
        off[1:] = np.cumsum([len(e) for e in err_probs_list])
        ep = np.ascontiguousarray(np.concatenate([np.asarray(e, np.float64) for e in err_probs_list])
                                  if n else np.zeros(0), np.float64)
        cn = np.ascontiguousarray(counts, np.int32).reshape(n, 3)
        bf = np.ascontiguousarray(bonf, np.int64).reshape(n)
        pv = np.zeros((n, 3), np.longdouble)
        lnp = np.zeros((n, 3), np.float64)
        st = np.zeros((n, 3), np.uint8)
        capi.check(self.lib.lfb200_snpcaller_batch(self._ctx, n, _ptr(ep), _ptr(off), _ptr(cn), _ptr(bf),
                                                    float(sig_level), _ptr(pv), _ptr(lnp), _ptr(st)))
        return pv, lnp, st

    # ---- poissbin(): the row of ln probabilities (snpcaller.h:93-96) -------------------------------
    def poissbin_batch(self, err_probs_list, num_failures, bonf, sig):
        """[(pvalue longdouble, row float64[K+1], n_end)] — row may be partial after the early exit, like the reference's"""
        n = len(err_probs_list)
        off = np.zeros(n + 1, np.int64)
        off[1:] = np.cumsum([len(e) for e in err_probs_list])
        ep = np.ascontiguousarray(np.concatenate([np.asarray(e, np.float64) for e in err_probs_list]) if n else np.zeros(0))
        ks = np.ascontiguousarray(num_failures, np.int32).reshape(n)
        bf = np.ascontiguousarray(bonf, np.int64).reshape(n)
        roff = np.zeros(n + 1, np.int64)
        roff[1:] = np.cumsum(ks.astype(np.int64) + 1)
        rows = np.zeros(int(roff[-1]), np.float64)
        pv = np.zeros(n, np.longdouble)
        nend = np.zeros(n, np.int32)
        capi.check(self.lib.lfb200_poissbin_batch(self._ctx, n, _ptr(ep), _ptr(off), _ptr(ks), _ptr(bf), float(sig), _ptr(roff),
                                                   _ptr(rows), _ptr(pv), _ptr(nend)))
        return [(pv[i], rows[roff[i]:roff[i + 1]].copy(), int(nend[i])) for i in range(n)]

    def poissbin(self, err_probs, num_failures, bonf, sig):
        """the link-compatible symbol: malloc()ed row handed back, freed here"""
        ep = np.ascontiguousarray(err_probs, np.float64)
        pv = np.zeros(1, np.longdouble)
        p = self.lib.lfb200_poissbin(_ptr(pv), _ptr(ep), len(ep), int(num_failures), int(bonf), float(sig))
        if not p:
            raise capi.Lfb200Error("lfb200_poissbin returned NULL")
        row = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_double)), shape=(int(num_failures) + 1,)).copy()
        C.CDLL(None).free(C.c_void_p(p))
        return pv[0], row

    # ---- plp_to_errprobs() (snpcaller.h:72-75) ----------------------------------------------------
    def batch_errprobs(self, batch, conf=None):
        """per column: (err_probs in pileup order, alt_bases, alt_counts, alt_raw_counts)"""
        cf = conf_from(conf)
        keep = []

        def arr(x, dt):
            if x is None:
                return None
            a = np.ascontiguousarray(x, dtype=dt)
            keep.append(a)
            return _ptr(a)
        n = len(batch["ref_base"])
        col_off = np.ascontiguousarray(batch["col_off"], np.int64)
        hb = capi.Batch(n, arr(col_off, np.int64), arr(batch["nt_cnt"], np.int32), arr(batch["ref_base"], np.uint8), None,
                        arr(batch["bq"], np.uint8), arr(batch.get("mq"), np.uint8), arr(batch.get("baq"), np.uint8),
                        arr(batch.get("sq"), np.uint8), None)
        ep = np.zeros(max(int(col_off[-1]), 1), np.float64)
        ne = np.zeros(n, np.int32)
        ab = np.zeros((n, 3), np.int32); ac = np.zeros((n, 3), np.int32); ar = np.zeros((n, 3), np.int32)
        capi.check(self.lib.lfb200_batch_errprobs(self._ctx, C.byref(cf), C.byref(hb), _ptr(ep), _ptr(ne), _ptr(ab), _ptr(ac), _ptr(ar)))
        return [(ep[col_off[c]:col_off[c] + ne[c]].copy(), ab[c], ac[c], ar[c]) for c in range(n)]

    def plp_to_errprobs(self, ref_base, coverage_plp, base_quals, map_quals=None, baq_quals=None, source_quals=None, conf=None):
        """the link-compatible symbol on one column given as int arrays per A,C,G,T (plp_col_t varrays)"""
        cf = conf_from(conf)
        keep = []
        pc = capi.PlpCol()
        pc.ref_base = ref_base.encode() if isinstance(ref_base, str) else ref_base
        pc.coverage_plp = int(coverage_plp)
        for g in range(4):
            pc.n[g] = len(base_quals[g])
            for name, src in (("base_quals", base_quals), ("map_quals", map_quals), ("baq_quals", baq_quals), ("source_quals", source_quals)):
                if src is None or (name != "base_quals" and len(src[g]) == 0):
                    getattr(pc, name)[g] = None
                    continue
                a = np.ascontiguousarray(src[g], np.int32)
                keep.append(a)
                getattr(pc, name)[g] = a.ctypes.data if len(a) else None
        p = C.c_void_p()
        ne = C.c_int(0)
        ab = np.zeros(3, np.int32); ac = np.zeros(3, np.int32); ar = np.zeros(3, np.int32)
        self.lib.lfb200_plp_to_errprobs(C.byref(p), C.byref(ne), _ptr(ab), _ptr(ac), _ptr(ar), C.byref(pc), C.byref(cf))
        if not p.value:
            raise capi.Lfb200Error("lfb200_plp_to_errprobs returned NULL")
        ep = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_double)), shape=(max(ne.value, 1),))[:ne.value].copy()
        C.CDLL(None).free(p)
        return ep, ab, ac, ar

    # ---- indel tests (call_indels, lofreq_call.c:618-726) -------------------------------------------
    def indel_tests(self, read_off, iq, mq, aq, sq, event_count, bonf_indel, conf=None):
        """one test per indel event; reads of the event last in its range; 255 = quality not available.
        Returns dict(pvalues longdouble[n], lnp, status, called, qual)"""
        cf = conf_from(conf)
        n = len(event_count)
        ro = np.ascontiguousarray(read_off, np.int64)
        total = int(ro[-1]) if n else 0

        def plane(x):
            if x is None:
                return None, None
            a = np.zeros(total + 32, np.uint8)
            a[:total] = np.asarray(x, np.uint8)[:total]
            return a, _ptr(a)
        k_iq, p_iq = plane(iq); k_mq, p_mq = plane(mq); k_aq, p_aq = plane(aq); k_sq, p_sq = plane(sq)
        ec = np.ascontiguousarray(event_count, np.int32)
        bf = np.ascontiguousarray(bonf_indel, np.int64)
        out = dict(pvalues=np.zeros(n, np.longdouble), lnp=np.zeros(n), status=np.zeros(n, np.uint8),
                   called=np.zeros(n, np.uint8), qual=np.zeros(n, np.int32))
        capi.check(self.lib.lfb200_indel_tests(self._ctx, C.byref(cf), n, _ptr(ro), p_iq, p_mq, p_aq, p_sq, _ptr(ec), _ptr(bf),
                                                _ptr(out["pvalues"]), _ptr(out["lnp"]), _ptr(out["status"]), _ptr(out["called"]),
                                                _ptr(out["qual"])))
        return out

    # ---- binom(): binomial CDF / survival function (binom.c:52-93) ------------------------------
    def binom_batch(self, num_trials, num_success, prob_success):
        """(status, p, q) arrays; p = P(X <= num_success), q = 1 - p, NaN where status != 0"""
        nt = np.ascontiguousarray(num_trials, np.int32).reshape(-1)
        ns = np.ascontiguousarray(num_success, np.int32).reshape(-1)
        pr = np.ascontiguousarray(prob_success, np.float64).reshape(-1)
        n = len(nt)
        p = np.full(n, np.nan)
        q = np.full(n, np.nan)
        st = np.zeros(n, np.int32)
        capi.check(self.lib.lfb200_binom_batch(self._ctx, n, _ptr(nt), _ptr(ns), _ptr(pr), _ptr(p), _ptr(q), _ptr(st)))
        return st, p, q

    def binom(self, num_trials, num_success, prob_success):
        st, p, q = self.binom_batch([num_trials], [num_success], [prob_success])
        return int(st[0]), float(p[0]), float(q[0])

    # ---- host batch -----------------------------------------------------------------------
    def kpa_glocal(self, reads, d=0.00001, e=0.4, bw=10, use_qual=True):
        """kpa_ext_glocal (kprobaln_ext.h:38-40) for a batch of reads: reads = dict(n, ref, ref_off, query, qry_off, qual) as
        CSR arrays (oracle.pyoracle.synth_reads has the layout); returns (state, q) per query base."""
        tot = int(reads["qry_off"][-1])
        state = np.zeros(max(tot, 1), np.int32)
        q = np.zeros(max(tot, 1), np.uint8)

        def p(a):
            return a.ctypes.data_as(C.c_void_p)
        capi.check(self.lib.lfb200_kpa_glocal_batch(self._ctx, int(reads["n"]), p(np.ascontiguousarray(reads["ref"], np.uint8)),
                                                    p(np.ascontiguousarray(reads["ref_off"], np.int64)),
                                                    p(np.ascontiguousarray(reads["query"], np.uint8)),
                                                    p(np.ascontiguousarray(reads["qry_off"], np.int64)),
                                                    p(np.ascontiguousarray(reads["qual"], np.uint8)) if use_qual else None,
                                                    d, e, bw, p(state), p(q)))
        return state[:tot], q[:tot]

    def call_columns(self, batch, conf=None, dense=True, max_sites=None):
        """batch: dict of numpy arrays (col_off, nt_cnt, ref_base, bq, mq, baq, sq, coverage), the packed
        form of plp_col_t described in include/lofreq_b200.h.  Returns a dict with the dense per-column
        arrays (when dense=True), 'sites' (structured array) and the advanced counters."""
        cf = conf_from(conf)
        keep = []

        def arr(x, dt):
            if x is None:
                return None
            a = np.ascontiguousarray(x, dtype=dt)
            keep.append(a)
            return _ptr(a)
        n = len(batch["ref_base"])
        hb = capi.Batch(n, arr(batch["col_off"], np.int64), arr(batch["nt_cnt"], np.int32),
                        arr(batch["ref_base"], np.uint8), arr(batch.get("coverage"), np.int32),
                        arr(batch["bq"], np.uint8), arr(batch.get("mq"), np.uint8),
                        arr(batch.get("baq"), np.uint8), arr(batch.get("sq"), np.uint8), arr(batch.get("num_bases"), np.int32))
        max_sites = n if max_sites is None else max_sites
        sites = (capi.Site * max(max_sites, 1))()
        sm = capi.Summary()
        out = {}
        dptr = None
        if dense:
            out = dict(alt_counts=np.zeros((n, 3), np.int32), alt_raw_counts=np.zeros((n, 3), np.int32),
                       tested=np.zeros(n, np.uint8), bonf_used=np.zeros(n, np.int64),
                       lnp=np.zeros((n, 3), np.float64), status=np.zeros((n, 3), np.uint8),
                       pvalues=np.zeros((n, 3), np.longdouble), called=np.zeros((n, 3), np.uint8),
                       qual=np.zeros((n, 3), np.int32))
            d = capi.DenseOut(*[_ptr(out[k]) for k in ("alt_counts", "alt_raw_counts", "tested", "bonf_used",
                                                       "lnp", "status", "pvalues", "called", "qual")])
            dptr = C.byref(d)
        capi.check(self.lib.lfb200_call_columns(self._ctx, C.byref(cf), C.byref(hb), dptr, sites, max_sites,
                                                 C.byref(sm)))
        out["sites"] = capi.sites_to_numpy(sites, sm.n_sites)
        out["n_sites"] = sm.n_sites
        out["n_tested"] = sm.n_tested
        out["n_heavy"] = sm.n_heavy
        out["n_unsupported"] = sm.n_unsupported
        jc = (C.c_longlong * 4)()
        capi.check(self.lib.lfb200_last_job_counts(self._ctx, jc))
        out["job_counts"] = dict(packed=jc[0], fallback=jc[1], per_column=jc[2], mid=jc[3])
        out["bonf_subst"] = cf.bonf_subst
        out["num_snv_tests"] = cf.num_snv_tests
        return out

    # ---- device-resident batch (torch tensors or raw pointers) -----------------------------
    @staticmethod
    def device_batch(t):
        """t: dict of CUDA torch tensors (see lofreq_b200.synth.generate_device)."""
        def p(x):
            return None if x is None else C.c_void_p(x.data_ptr())
        return capi.Batch(int(t["ref_base"].numel()), p(t["col_off"]), p(t["nt_cnt"]), p(t["ref_base"]),
                          p(t.get("coverage")), p(t["bq"]), p(t.get("mq")), p(t.get("baq")), p(t.get("sq")), p(t.get("num_bases")))

    def screen(self, dev_batch, conf, stream=None):
        capi.check(self.lib.lfb200_screen_device(self._ctx, C.byref(conf), C.byref(dev_batch), stream))

    def ntested(self, stream=None):
        n = C.c_longlong()
        capi.check(self.lib.lfb200_ntested_device(self._ctx, stream, C.byref(n)))
        return n.value

    def test(self, conf, stream=None):
        capi.check(self.lib.lfb200_test_device(self._ctx, C.byref(conf), stream))

    def sites(self, conf, max_sites, stream=None):
        sites = (capi.Site * max(max_sites, 1))()
        sm = capi.Summary()
        capi.check(self.lib.lfb200_sites_device(self._ctx, C.byref(conf), stream, sites, max_sites, C.byref(sm)))
        return capi.sites_to_numpy(sites, sm.n_sites), sm


class ColumnBuilder:
    """Mirror of the per-column callback surface (plp.h:159-163): feed one pileup column at a time the way
    mpileup() feeds call_vars(); sites come back through on_site(site_dict) in input order."""

    def __init__(self, caller, conf, batch_cols, on_site):
        self.caller = caller
        self.conf = conf_from(conf)
        self.lib = caller.lib
        self._on_site = on_site

        def tramp(site_p, tag, ref_base, coverage, user):
            s = site_p.contents
            on_site(dict(tag=tag, ref_base=ref_base.decode(), coverage=coverage, col=s.col, bonf=s.bonf,
                         lnp=list(s.lnp), pvalue=np.array(list(s.pvalue), np.longdouble), alt_count=list(s.alt_count),
                         alt_raw_count=list(s.alt_raw_count), qual=list(s.qual), status=list(s.status),
                         called=list(s.called)))
        self._tramp = capi.SITE_FN(tramp)
        self._b = C.c_void_p()
        capi.check(self.lib.lfb200_builder_create(C.byref(self._b), caller._ctx, C.byref(self.conf), int(batch_cols),
                                                   self._tramp, None))

    def add_column(self, tag, ref_base, coverage_plp, num_bases, base_quals, map_quals=None, baq_quals=None,
                   source_quals=None):
        """*_quals: four int arrays (A, C, G, T) like plp_col_t.base_quals[i].data"""
        keep = []

        def four(groups):
            if groups is None:
                return None
            arr = (C.c_void_p * 4)()
            for i, g in enumerate(groups):
                a = np.ascontiguousarray(g, np.int32)
                keep.append(a)
                arr[i] = a.ctypes.data if len(a) else None
            keep.append(arr)
            return C.cast(arr, C.c_void_p)
        n = np.array([len(g) for g in base_quals], np.int32)
        capi.check(self.lib.lfb200_builder_add_column(self._b, int(tag), ref_base.encode(), int(coverage_plp), int(num_bases),
                                                       four(base_quals), four(map_quals), four(baq_quals),
                                                       four(source_quals), n.ctypes.data_as(C.c_void_p)))

    def flush(self):
        capi.check(self.lib.lfb200_builder_flush(self._b))

    def close(self):
        if self._b:
            self.lib.lfb200_builder_destroy(self._b)
            self._b = C.c_void_p()
