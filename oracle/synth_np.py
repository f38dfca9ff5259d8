"""TEST INFRASTRUCTURE — numpy restatement of the synthetic pileup-column model
(SURVEY.md §8(d)), bit-identical to the device generator in
lofreq_b200/csrc/synth.cu (tests/test_synth.py checks that on the GPU).

Everything is integer arithmetic on a stateless 64-bit hash, so host and device
produce the same columns:

  h(c, r, salt) = splitmix64(seed ^ (c << 20) ^ r ^ (salt << 56))

Per read r of column c:  bq from the workload's quality model, mq = 60,
(optional) baq uniform 30..60; the read shows the reference base ACGT[c & 3]
unless it errs — with probability 10^(-bq/10), compared as a 32-bit integer
threshold — to one of the three other bases, or it is one of the first
round(AF*depth) reads of a variant column (1 % of columns; AF log-uniform in
[0.5 %, 50 %]) which carry ACGT[(c + 1) & 3].  Reads are then grouped by base in
A,C,G,T order (stable), which is the layout plp_col_t has (plp.h:89-92).
"""
import numpy as np

SEED = 20261017
MASK = np.uint64(0xFFFFFFFFFFFFFFFF)

WORKLOADS = {
    # name: (depth, quality model)
    "C2": dict(depth=500, qmodel="q30"),
    "C3": dict(depth=2000, qmodel="u20_40"),
    "C4": dict(depth=300, qmodel="mix30"),
    "C5": dict(depth=None, qmodel="u20_40"),     # depth log-uniform 50..10000 per column
}


def splitmix64(x):
    x = np.asarray(x, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = x + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def err_threshold_table():
    """thr[q] = floor(10^(-q/10) * 2^32) as uint32 (q = 0 saturates)."""
    q = np.arange(256, dtype=np.float64)
    t = np.floor(np.power(10.0, -q / 10.0) * 4294967296.0)
    return np.minimum(t, 4294967295.0).astype(np.uint32)


# AF table: 1024 log-uniform steps in [0.005, 0.5], stored as round(AF * 2^20)
def af_table():
    i = np.arange(1024, dtype=np.float64)
    af = 0.005 * np.power(100.0, (i + 0.5) / 1024.0)
    return np.round(af * (1 << 20)).astype(np.uint32)


# depth table for C5: 1024 log-uniform steps in [50, 10000]
def depth_table():
    i = np.arange(1024, dtype=np.float64)
    return np.round(50.0 * np.power(200.0, (i + 0.5) / 1024.0)).astype(np.uint32)


def col_hash(cols, salt, seed=SEED):
    cols = np.asarray(cols, dtype=np.uint64)
    return splitmix64(np.uint64(seed) ^ (cols << np.uint64(20)) ^ (np.uint64(salt) << np.uint64(56)))


def read_hash(cols, reads, salt, seed=SEED):
    cols = np.asarray(cols, dtype=np.uint64)
    reads = np.asarray(reads, dtype=np.uint64)
    return splitmix64(np.uint64(seed) ^ (cols << np.uint64(20)) ^ reads ^ (np.uint64(salt) << np.uint64(56)))


def column_depths(workload, c0, n_cols):
    w = WORKLOADS[workload]
    cols = np.arange(c0, c0 + n_cols, dtype=np.uint64)
    if w["depth"] is not None:
        return np.full(n_cols, w["depth"], dtype=np.int64)
    idx = (col_hash(cols, 3) >> np.uint64(54)).astype(np.int64)       # top 10 bits
    return depth_table()[idx].astype(np.int64)


def generate(workload, c0, n_cols, with_baq=False, seed=SEED, pad=16, with_strand=False):
    """Packed column batch (dict, see oracle/column_batch.h) for columns
    [c0, c0+n_cols) of the named workload."""
    w = WORKLOADS[workload]
    thr = err_threshold_table()
    aft = af_table()
    depths = column_depths(workload, c0, n_cols)
    pitch = (depths + pad - 1) // pad * pad
    col_off = np.zeros(n_cols + 1, dtype=np.int64)
    np.cumsum(pitch, out=col_off[1:])
    total = int(col_off[-1])
    bq_pl = np.zeros(total, np.uint8)
    mq_pl = np.zeros(total, np.uint8)
    baq_pl = np.zeros(total, np.uint8) if with_baq else None
    nt_cnt = np.zeros((n_cols, 4), np.int32)
    strand8 = np.zeros((n_cols, 8), np.int32) if with_strand else None     # fw A,C,G,T then rv A,C,G,T (plp_col_t.fw_counts / rv_counts)
    ref = np.frombuffer(b"ACGT", dtype=np.uint8)[(np.arange(c0, c0 + n_cols) & 3)]

    # group columns by depth so the per-read work vectorises
    for d in np.unique(depths):
        sel = np.nonzero(depths == d)[0]
        cols = (c0 + sel).astype(np.uint64)[:, None]
        reads = np.arange(d, dtype=np.uint64)[None, :]
        hq = read_hash(cols, reads, 0, seed)
        he = read_hash(cols, reads, 1, seed)
        if w["qmodel"] == "q30":
            bq = np.full(hq.shape, 30, np.int64)
        elif w["qmodel"] == "u20_40":
            bq = 20 + ((hq & np.uint64(0xFFFF)) % np.uint64(21)).astype(np.int64)
        elif w["qmodel"] == "mix30":
            u = 20 + ((hq & np.uint64(0xFFFF)) % np.uint64(21)).astype(np.int64)
            keep30 = ((hq >> np.uint64(16)) & np.uint64(0xFFFF)) < np.uint64(45875)   # 0.7 * 65536
            bq = np.where(keep30, 30, u)
        else:
            raise ValueError(w["qmodel"])
        baq = 30 + (((hq >> np.uint64(32)) & np.uint64(0xFFFF)) % np.uint64(31)).astype(np.int64)
        ref_nt = ((c0 + sel) & 3)[:, None]
        errs = (he >> np.uint64(32)).astype(np.uint32) < thr[bq]
        err_nt = (ref_nt + 1 + ((he & np.uint64(0xFFFF)) % np.uint64(3)).astype(np.int64)) & 3
        nt = np.where(errs, err_nt, ref_nt)
        hv = col_hash((c0 + sel).astype(np.uint64), 2, seed)
        is_var = (hv % np.uint64(100)) == 0
        af_q20 = aft[((hv >> np.uint64(40)) & np.uint64(1023)).astype(np.int64)].astype(np.int64)
        n_alt = (af_q20 * int(d) + (1 << 19)) >> 20
        n_alt = np.where(is_var, n_alt, 0)[:, None]
        var_nt = ((c0 + sel + 1) & 3)[:, None]
        nt = np.where(reads.astype(np.int64) < n_alt, var_nt, nt)

        order = np.argsort(nt, axis=1, kind="stable")
        bq_s = np.take_along_axis(bq, order, axis=1).astype(np.uint8)
        baq_s = np.take_along_axis(baq, order, axis=1).astype(np.uint8)
        for g in range(4):
            nt_cnt[sel, g] = (nt == g).sum(axis=1)
        if with_strand:
            rv = ((he >> np.uint64(16)) & np.uint64(1)).astype(bool)        # strand = one bit of the read's hash
            for g in range(4):
                strand8[sel, g] = ((nt == g) & ~rv).sum(axis=1)
                strand8[sel, 4 + g] = ((nt == g) & rv).sum(axis=1)
        idx = col_off[sel][:, None] + np.arange(d)[None, :]
        bq_pl[idx] = bq_s
        mq_pl[idx] = 60
        if with_baq:
            baq_pl[idx] = baq_s
    return dict(col_off=col_off, nt_cnt=nt_cnt, ref_base=ref.copy(), bq=bq_pl, mq=mq_pl,
                baq=baq_pl, sq=None, coverage=None, depths=depths, strand8=strand8)
