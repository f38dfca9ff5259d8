"""TEST INFRASTRUCTURE — ctypes front end for the two CPU checkers under oracle/.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline leg and the
``--impl reference`` arm) may import this module.  The product package
(lofreq_b200) never does.

  kind="reference": oracle/_ref/libsnpref.so — the unmodified reference sources
                    (/root/reference/src/lofreq/{snpcaller,utils,log}.c) plus
                    oracle/ref_harness.c.  Built only where /root/reference is
                    mounted; the built .so travels with the repo snapshot.
  kind="port":      oracle/libsnvoracle.so — oracle/snv_oracle.c, the from-scratch
                    C restatement (buildable anywhere with gcc).
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libsnpref.so")
BINOM_SO = os.path.join(HERE, "_ref", "libbinomref.so")
PORT_SO = os.path.join(HERE, "libsnvoracle.so")

VARCALL_USE_BAQ, VARCALL_USE_MQ, VARCALL_USE_SQ, VARCALL_USE_IDAQ = 1, 2, 4, 8   # defaults.h:76-80
LDBL_MAX = np.finfo(np.longdouble).max
LDBL_MIN = np.finfo(np.longdouble).tiny


def build(quiet=True):
    """Compile the checkers (gcc). The reference half is skipped when
    /root/reference is not mounted (the GPU box): the prebuilt .so is kept."""
    out = subprocess.run(["make", "-C", HERE], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
    if not quiet:
        print(out.stdout)


class _Conf(C.Structure):
    _fields_ = [("min_bq", C.c_int), ("min_alt_bq", C.c_int), ("def_alt_bq", C.c_int),
                ("min_jq", C.c_int), ("min_alt_jq", C.c_int), ("def_alt_jq", C.c_int),
                ("min_cov", C.c_int), ("bonf_dynamic", C.c_int), ("flag", C.c_int),
                ("sig", C.c_float), ("bonf_subst", C.c_longlong), ("num_snv_tests", C.c_longlong)]


class _Batch(C.Structure):
    _fields_ = [("n_cols", C.c_longlong), ("col_off", C.c_void_p), ("nt_cnt", C.c_void_p),
                ("ref_base", C.c_void_p), ("coverage", C.c_void_p),
                ("bq", C.c_void_p), ("mq", C.c_void_p), ("baq", C.c_void_p), ("sq", C.c_void_p),
                ("num_bases", C.c_void_p)]


class _Out(C.Structure):
    _fields_ = [("alt_counts", C.c_void_p), ("alt_raw_counts", C.c_void_p), ("tested", C.c_void_p),
                ("bonf_used", C.c_void_p), ("pvalues", C.c_void_p), ("called", C.c_void_p),
                ("qual", C.c_void_p)]


def default_conf(**over):
    """init_varcall_conf defaults (snpcaller.c:626-651)."""
    d = dict(min_bq=6, min_alt_bq=6, def_alt_bq=0, min_jq=0, min_alt_jq=0, def_alt_jq=0,
             min_cov=1, bonf_dynamic=1, flag=VARCALL_USE_MQ | VARCALL_USE_BAQ | VARCALL_USE_IDAQ,
             sig=0.01, bonf_subst=1, num_snv_tests=0)
    d.update(over)
    return d


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Oracle:
    def __init__(self, kind="port"):
        self.kind = kind
        path = REF_SO if kind == "reference" else PORT_SO
        if not os.path.exists(path):
            build()
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = lib = C.CDLL(path)
        self.libc = C.CDLL(None)
        self.libc.free.argtypes = [C.c_void_p]
        pre = "lfref_" if kind == "reference" else "lfo_"
        self._call_columns = getattr(lib, pre + "call_columns")
        self._call_columns.restype = C.c_int
        self._call_columns.argtypes = [C.POINTER(_Conf), C.POINTER(_Batch), C.POINTER(_Out)]
        self._errprobs = getattr(lib, pre + "column_errprobs")
        self._errprobs.restype = C.c_int
        self._errprobs.argtypes = [C.POINTER(_Conf), C.POINTER(_Batch), C.c_longlong,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        if kind == "reference":
            names = dict(snpcaller="snpcaller", poissbin="poissbin", merge="merge_srcq_mapq_baq_and_bq",
                         log_sum="log_sum", tailsum="probvec_tailsum", dp="pruned_calc_prob_dist",
                         p2q="lfref_prob_to_phredqual", p2q_safe="lfref_prob_to_phredqual_safe",
                         q2p="lfref_phredqual_to_prob")
        else:
            names = dict(snpcaller="lfo_snpcaller", poissbin="lfo_poissbin", merge="lfo_merge_quals",
                         log_sum="lfo_log_sum", tailsum="lfo_tailsum", dp="lfo_pruned_dp",
                         p2q="lfo_prob_to_phred", p2q_safe="lfo_prob_to_phred_safe", q2p="lfo_phred_to_prob")
        f = getattr(lib, names["snpcaller"])
        f.restype = C.c_int
        if kind == "reference":   # snpcaller.h:97-102 has the trailing approx_threshold_n
            f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_longlong, C.c_double, C.c_int]
        else:
            f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_longlong, C.c_double]
        self._snpcaller = f
        f = getattr(lib, names["poissbin"])
        f.restype = C.c_void_p
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_longlong, C.c_double]
        self._poissbin = f
        f = getattr(lib, names["merge"]); f.restype = C.c_double; f.argtypes = [C.c_int] * 4
        self._merge = f
        f = getattr(lib, names["log_sum"]); f.restype = C.c_double; f.argtypes = [C.c_double] * 2
        self.log_sum = f
        f = getattr(lib, names["tailsum"]); f.restype = C.c_double; f.argtypes = [C.c_void_p, C.c_int, C.c_int]
        self._tailsum = f
        f = getattr(lib, names["p2q"]); f.restype = C.c_int; f.argtypes = [C.c_longdouble]
        self.prob_to_phred = f
        f = getattr(lib, names["p2q_safe"]); f.restype = C.c_int; f.argtypes = [C.c_double]
        self.prob_to_phred_safe = f
        f = getattr(lib, names["q2p"]); f.restype = C.c_double; f.argtypes = [C.c_int]
        self.phred_to_prob = f

    # -- scalar/vector entry points -------------------------------------
    def merge(self, sq, mq, baq, bq):
        return self._merge(int(sq), int(mq), int(baq), int(bq))

    def snpcaller(self, err_probs, counts, bonf, sig):
        ep = np.ascontiguousarray(err_probs, dtype=np.float64)
        cn = np.ascontiguousarray(counts, dtype=np.int32)
        pv = np.empty(3, dtype=np.longdouble)
        args = [_ptr(pv), _ptr(ep), len(ep), _ptr(cn), int(bonf), float(sig)]
        if self.kind == "reference":
            args.append(-1)
        rc = self._snpcaller(*args)
        if rc:
            raise RuntimeError("snpcaller returned %d" % rc)
        return pv

    def poissbin(self, err_probs, k, bonf, sig):
        """returns (pvalue longdouble, row float64[k+1]) — row is the malloc'd
        vector the reference hands back (may be partial after an early exit)."""
        ep = np.ascontiguousarray(err_probs, dtype=np.float64)
        pv = np.empty(1, dtype=np.longdouble)
        p = self._poissbin(_ptr(pv), _ptr(ep), len(ep), int(k), int(bonf), float(sig))
        row = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_double)), shape=(k + 1,)).copy()
        self.libc.free(p)
        return pv[0], row

    def indel_test(self, is_del, flag, sig, bonf, other_q, other_mq, events, test_event):
        """reference only: events = list of (q, aq, mq, sq) int arrays (-1 = not available) -> (pvalue, count, n_err_probs)"""
        if self.kind != "reference":
            raise RuntimeError("the indel harness drives the compiled reference (oracle/_ref/libsnpref.so)")
        f = self.lib.lfref_indel_test
        f.restype = C.c_int
        f.argtypes = [C.c_int, C.c_int, C.c_double, C.c_longlong, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        oq = np.ascontiguousarray(other_q, np.int32); om = np.ascontiguousarray(other_mq, np.int32)
        off = np.zeros(len(events) + 1, np.int32)
        off[1:] = np.cumsum([len(e[0]) for e in events])
        cat = [np.ascontiguousarray(np.concatenate([np.asarray(e[k], np.int32) for e in events]), np.int32) for k in range(4)]
        pv = np.zeros(1, np.longdouble)
        cnt = C.c_int(0)
        n = f(int(is_del), int(flag), float(sig), int(bonf), len(oq), _ptr(oq), _ptr(om), len(events), _ptr(off),
              _ptr(cat[0]), _ptr(cat[1]), _ptr(cat[2]), _ptr(cat[3]), int(test_event), _ptr(pv), C.byref(cnt))
        return pv[0], cnt.value, n

    def tailsum(self, row, start):
        r = np.ascontiguousarray(row, dtype=np.float64)
        return self._tailsum(_ptr(r), int(start), len(r))

    # -- batch entry points ---------------------------------------------
    @staticmethod
    def _mk_batch(b):
        keep = []

        def arr(x, dt):
            if x is None:
                return None
            a = np.ascontiguousarray(x, dtype=dt)
            keep.append(a)
            return _ptr(a)
        n = len(b["ref_base"])
        s = _Batch(n, arr(b["col_off"], np.int64), arr(b["nt_cnt"], np.int32), arr(b["ref_base"], np.uint8),
                   arr(b.get("coverage"), np.int32), arr(b["bq"], np.uint8), arr(b.get("mq"), np.uint8),
                   arr(b.get("baq"), np.uint8), arr(b.get("sq"), np.uint8), arr(b.get("num_bases"), np.int32))
        return s, keep, n

    def call_columns(self, batch, conf=None):
        """Run the whole per-column path over a packed batch (dict of numpy
        arrays, see oracle/column_batch.h). Returns dict of outputs plus the
        final bonf_subst / num_snv_tests."""
        conf = default_conf() if conf is None else conf
        cf = _Conf(**conf)
        sb, keep, n = self._mk_batch(batch)
        out = dict(alt_counts=np.zeros((n, 3), np.int32), alt_raw_counts=np.zeros((n, 3), np.int32),
                   tested=np.zeros(n, np.uint8), bonf_used=np.zeros(n, np.int64),
                   pvalues=np.zeros((n, 3), np.longdouble), called=np.zeros((n, 3), np.uint8),
                   qual=np.zeros((n, 3), np.int32))
        so = _Out(*[_ptr(out[k]) for k in ("alt_counts", "alt_raw_counts", "tested", "bonf_used",
                                            "pvalues", "called", "qual")])
        rc = self._call_columns(C.byref(cf), C.byref(sb), C.byref(so))
        if rc:
            raise RuntimeError("call_columns returned %d" % rc)
        out["bonf_subst"] = cf.bonf_subst
        out["num_snv_tests"] = cf.num_snv_tests
        return out

    def column_errprobs(self, batch, c, conf=None):
        conf = default_conf() if conf is None else conf
        cf = _Conf(**conf)
        sb, keep, n = self._mk_batch(batch)
        depth = int(np.asarray(batch["nt_cnt"]).reshape(-1, 4)[c].sum())
        ep = np.zeros(max(depth, 1), np.float64)
        ab = np.zeros(3, np.int32); ac = np.zeros(3, np.int32); ar = np.zeros(3, np.int32)
        m = self._errprobs(C.byref(cf), C.byref(sb), int(c), _ptr(ep), _ptr(ab), _ptr(ac), _ptr(ar))
        return ep[:m].copy(), ab, ac, ar


class BinomRef:
    """binom() of the reference (src/lofreq/binom.c:52-93 over cdflib90)."""

    def __init__(self):
        if not os.path.exists(BINOM_SO):
            build()
        self.lib = C.CDLL(BINOM_SO)
        self.lib.binom.restype = C.c_int
        self.lib.binom.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int, C.c_int, C.c_double]

    def cdf_sf(self, num_trials, num_success, prob):
        p, q = C.c_double(), C.c_double()
        rc = self.lib.binom(C.byref(p), C.byref(q), int(num_trials), int(num_success), float(prob))
        if rc:
            raise RuntimeError("binom status %d" % rc)
        return p.value, q.value


class _Indels(C.Structure):
    """oracle_indels_t of oracle/call_harness.c"""
    _fields_ = [("n_events", C.c_longlong), ("ev_col", C.c_void_p), ("ev_is_del", C.c_void_p), ("ev_key", C.c_void_p),
                ("ev_read_off", C.c_void_p), ("ev_q", C.c_void_p), ("ev_aq", C.c_void_p), ("ev_mq", C.c_void_p), ("ev_sq", C.c_void_p),
                ("ev_rv", C.c_void_p), ("non_off", C.c_void_p), ("non_iq", C.c_void_p), ("non_dq", C.c_void_p), ("non_mq", C.c_void_p),
                ("hrun", C.c_void_p), ("num_tails", C.c_void_p), ("non_ins_fw_rv", C.c_void_p), ("non_del_fw_rv", C.c_void_p)]


MAX_INDELSIZE = 256          # utils.h


def synth_indels(n_cols, depths, seed=7, frac=0.06):
    """Synthetic indel events for the columns of a batch (test input for the call_indels path): per column the reads
    without an indel (insertion / deletion / mapping quality each), and for a fraction of the columns 1-3 events (key,
    reads with indel quality, alignment quality, mapping quality, strand) — low-AF single-base A/T pairs included, which
    the reference's poly-AT filter ignores (lofreq_call.c:650-680)."""
    rng = np.random.default_rng(seed)
    depths = np.asarray(depths, np.int64)
    ev_col, ev_is_del, keys, ev_reads = [], [], [], []
    for c in np.flatnonzero(rng.random(n_cols) < frac):
        kind = rng.integers(0, 4)
        evs = []
        if kind == 3:                                    # poly-AT pair: ins X and del X, both rare
            b = "AT"[int(rng.integers(0, 2))]
            evs = [(0, b, int(rng.integers(1, 4))), (1, b, int(rng.integers(1, 4)))]
        else:
            for _ in range(int(rng.integers(1, 4))):
                key = "".join("ACGT"[int(x)] for x in rng.integers(0, 4, int(rng.integers(1, 6))))
                evs.append((int(rng.integers(0, 2)), key, int(rng.integers(1, max(2, min(40, depths[c] // 8))))))
        seen = set()
        for is_del, key, cnt in sorted(evs):               # insertions first (the harness wants them grouped per column)
            if (is_del, key) in seen:
                continue
            seen.add((is_del, key))
            ev_col.append(int(c)); ev_is_del.append(is_del); keys.append(key); ev_reads.append(cnt)
    order = sorted(range(len(ev_col)), key=lambda i: (ev_col[i], ev_is_del[i], i))
    ev_col = [ev_col[i] for i in order]; ev_is_del = [ev_is_del[i] for i in order]
    keys = [keys[i] for i in order]; ev_reads = [ev_reads[i] for i in order]
    # reads without an indel: what is left of the column's depth (plp_to_ins/del_errprobs size their vector by coverage_plp)
    used = np.zeros(n_cols, np.int64)
    for c, r in zip(ev_col, ev_reads):
        used[c] += r
    non_off = np.zeros(n_cols + 1, np.int64)
    non_off[1:] = np.cumsum(np.maximum(depths - used, 0))
    tot = int(non_off[-1])
    ind = dict(non_off=non_off, non_iq=rng.integers(25, 46, tot).astype(np.int32), non_dq=rng.integers(25, 46, tot).astype(np.int32),
               non_mq=np.where(rng.random(tot) < 0.01, 255, rng.integers(20, 61, tot)).astype(np.int32),
               hrun=rng.integers(1, 6, n_cols).astype(np.int32), num_tails=rng.integers(0, 5, n_cols).astype(np.int32),
               non_ins_fw_rv=rng.integers(0, 200, (n_cols, 2)).astype(np.int32), non_del_fw_rv=rng.integers(0, 200, (n_cols, 2)).astype(np.int32))
    ne = len(ev_col)
    ro = np.zeros(ne + 1, np.int64)
    ro[1:] = np.cumsum(ev_reads)
    nr = int(ro[-1])
    kb = np.zeros((ne, MAX_INDELSIZE), np.uint8)
    for i, k in enumerate(keys):
        kb[i, :len(k)] = np.frombuffer(k.encode(), np.uint8)
    ind.update(n_events=ne, ev_col=np.array(ev_col, np.int64), ev_is_del=np.array(ev_is_del, np.uint8), ev_key=kb, ev_read_off=ro,
               ev_q=rng.integers(15, 46, nr).astype(np.int32), ev_aq=np.where(rng.random(nr) < 0.05, -1, rng.integers(5, 61, nr)).astype(np.int32),
               ev_mq=np.where(rng.random(nr) < 0.02, 255, rng.integers(20, 61, nr)).astype(np.int32), ev_sq=np.full(nr, -1, np.int32),
               ev_rv=rng.integers(0, 2, nr).astype(np.uint8))
    return ind


CALLREF_SO = os.path.join(HERE, "_ref", "libcallref.so")
CALLB200_SO = os.path.join(HERE, "_ref", "libcallb200.so")
CALLSWAP_SO = os.path.join(HERE, "_ref", "libcallswap.so")


class CallOracle:
    """Callback-boundary oracle (oracle/call_harness.c): identical plp_col_t objects go through the reference's REAL
    call_vars() -> call_snvs() -> report_var() -> vcf_write_var() (lofreq_call.c, vcf.c, fet.c compiled unmodified), or —
    adapter=True, library _ref/libcallb200.so — through the product's drop-in callback lfb200_call_vars() + lfb200_flush()
    (lofreq_b200/adapter/lofreq_adapter.c compiled against the reference's own headers).  Returns the raw VCF text."""

    def __init__(self, adapter=False, swap=False):
        """swap=True: _ref/libcallswap.so — the reference's own call_vars() on top of the GPU-backed snpcaller() /
        plp_to_errprobs() / poissbin() symbols (lofreq_b200/adapter/snpcaller_shim.c)"""
        path = CALLSWAP_SO if swap else CALLB200_SO if adapter else CALLREF_SO
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.adapter = adapter
        self.lib = C.CDLL(path)
        f = self.lib.lfref_call_vars_vcf
        f.restype = C.c_int
        f.argtypes = [C.POINTER(_Conf), C.POINTER(_Batch), C.c_void_p, C.c_void_p, C.c_void_p, C.c_char_p, C.c_char_p,
                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        g = self.lib.lfref_sb_qual
        g.restype = C.c_int
        g.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]

    def sb_qual(self, ref_fw, ref_rv, alt_fw, alt_rv):
        two = C.c_double()
        return self.lib.lfref_sb_qual(int(ref_fw), int(ref_rv), int(alt_fw), int(alt_rv), C.byref(two)), two.value

    def call_vars_vcf(self, batch, strand8, conf=None, pos=None, cons0=None, target="synthetic", adapter=None, indels=None):
        """-> (vcf text, bonf_subst, num_snv_tests[, bonf_indel, num_indel_tests] with indels).  strand8: int32 [n][8] = fw
        A,C,G,T then rv A,C,G,T.  indels: dict from synth_indels() — switches --call-indels on (conf.no_indels = 0)."""
        import tempfile
        use_adapter = self.adapter if adapter is None else adapter
        conf = default_conf() if conf is None else conf
        cf = _Conf(**conf)
        sb, keep, n = Oracle._mk_batch(batch)
        s8 = np.ascontiguousarray(strand8, np.int32)
        ps = None if pos is None else np.ascontiguousarray(pos, np.int32)
        c0 = None if cons0 is None else np.ascontiguousarray(cons0, np.uint8)
        with tempfile.NamedTemporaryFile(suffix=".vcf", delete=False) as tf:
            path = tf.name
        try:
            ind_p, bi, nit, hold = None, C.c_longlong(1), C.c_longlong(0), []
            if indels is not None:
                st = _Indels()
                st.n_events = int(indels["n_events"])
                for k, dt in (("ev_col", np.int64), ("ev_is_del", np.uint8), ("ev_key", np.uint8), ("ev_read_off", np.int64), ("ev_q", np.int32),
                              ("ev_aq", np.int32), ("ev_mq", np.int32), ("ev_sq", np.int32), ("ev_rv", np.uint8), ("non_off", np.int64),
                              ("non_iq", np.int32), ("non_dq", np.int32), ("non_mq", np.int32), ("hrun", np.int32), ("num_tails", np.int32),
                              ("non_ins_fw_rv", np.int32), ("non_del_fw_rv", np.int32)):
                    a = np.ascontiguousarray(indels[k], dt)
                    hold.append(a)
                    setattr(st, k, a.ctypes.data)
                ind_p = C.cast(C.pointer(st), C.c_void_p)
            rc = self.lib.lfref_call_vars_vcf(C.byref(cf), C.byref(sb), _ptr(s8), _ptr(ps), _ptr(c0), target.encode(), path.encode(),
                                              ind_p, C.cast(C.byref(bi), C.c_void_p), C.cast(C.byref(nit), C.c_void_p),
                                              1 if use_adapter else 0)
            if rc:
                raise RuntimeError("lfref_call_vars_vcf returned %d" % rc)
            with open(path) as f:
                text = f.read()
        finally:
            os.unlink(path)
        if indels is not None:
            return text, cf.bonf_subst, cf.num_snv_tests, bi.value, nit.value
        return text, cf.bonf_subst, cf.num_snv_tests


def have_call_oracle(adapter=False, swap=False):
    return os.path.exists(CALLSWAP_SO if swap else CALLB200_SO if adapter else CALLREF_SO)


def have_reference():
    return os.path.exists(REF_SO)


# ------------------------------------------------------------------------------------------------
# BAQ HMM (kpa_ext_glocal, kprobaln_ext.c:80-277) — SURVEY.md 8f #2
# ------------------------------------------------------------------------------------------------
KPAREF_SO = os.path.join(HERE, "_ref", "libkparef.so")


def have_kpa_reference():
    return os.path.exists(KPAREF_SO)


def synth_reads(n, seed=1, lmin=30, lmax=151, flank=10, sub=0.01, indel=0.003, n_frac=0.002):
    """Reads the way bam_prob_realn_core_ext hands them to kpa_ext_glocal (bam_md_ext.c:380-407): a reference window of
    the read's span plus `flank` bases on either side (0..3, 4 = ambiguous), the read with substitutions, short indels and
    a few Ns, base qualities.  Returns CSR arrays."""
    rng = np.random.default_rng(seed)
    refs, qrys, quals = [], [], []
    for _ in range(n):
        lq = int(rng.integers(lmin, lmax))
        core = rng.integers(0, 4, lq + 8).astype(np.uint8)
        q = []
        i = 0
        while len(q) < lq and i < len(core):
            u = rng.random()
            if u < indel:                       # deletion from the read
                i += int(rng.integers(1, 4))
                continue
            if u < 2 * indel:                   # insertion into the read
                q.extend(rng.integers(0, 4, int(rng.integers(1, 4))).tolist())
                continue
            b = int(core[i])
            if rng.random() < sub:
                b = (b + int(rng.integers(1, 4))) % 4
            if rng.random() < n_frac:
                b = 4
            q.append(b)
            i += 1
        q = np.array(q[:lq] if len(q) >= lq else q + rng.integers(0, 4, lq - len(q)).tolist(), np.uint8)
        used = core[:max(i, 1)]
        left = rng.integers(0, 4, int(rng.integers(0, flank + 1))).astype(np.uint8)
        right = rng.integers(0, 4, int(rng.integers(0, flank + 1))).astype(np.uint8)
        r = np.concatenate([left, used, right])
        if rng.random() < 0.02:
            r[int(rng.integers(0, len(r)))] = 4
        refs.append(r); qrys.append(q)
        quals.append(rng.integers(2, 42, len(q)).astype(np.uint8))
    ref_off = np.concatenate([[0], np.cumsum([len(x) for x in refs])]).astype(np.int64)
    qry_off = np.concatenate([[0], np.cumsum([len(x) for x in qrys])]).astype(np.int64)
    return dict(n=n, ref=np.concatenate(refs), ref_off=ref_off, query=np.concatenate(qrys), qry_off=qry_off,
                qual=np.concatenate(quals))


class KpaRef:
    """the reference's kpa_ext_glocal, one call per read (oracle/kpa_harness.c)"""

    def __init__(self):
        if not os.path.exists(KPAREF_SO):
            raise RuntimeError("oracle/_ref/libkparef.so is not built (needs /root/reference: make -C oracle ref)")
        self.lib = C.CDLL(KPAREF_SO)
        self.lib.lfref_kpa_glocal_batch.restype = C.c_int

    def glocal(self, reads, d=0.00001, e=0.4, bw=10, use_qual=True):
        n = reads["n"]
        tot = int(reads["qry_off"][-1])
        state = np.zeros(tot, np.int32)
        q = np.zeros(tot, np.uint8)
        pr = np.zeros(n, np.int32)
        self.lib.lfref_kpa_glocal_batch(C.c_longlong(n), _ptr(reads["ref"]), _ptr(reads["ref_off"]), _ptr(reads["query"]),
                                        _ptr(reads["qry_off"]), _ptr(reads["qual"]) if use_qual else None, C.c_float(d),
                                        C.c_float(e), C.c_int(bw), _ptr(state), _ptr(q), _ptr(pr))
        return state, q, pr
