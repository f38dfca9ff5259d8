"""GPU micro-timing of the K > 2048 path: batches of identical deep columns through lfb200_call_columns (device time via
LFB200 profiling events: ms4[3] = the heavy phase)."""
import ctypes as C
import sys
import time
import numpy as np
sys.path.insert(0, ".")
import lofreq_b200
from lofreq_b200 import capi


def batch(ncol, n, k, q=(20, 41), seed=1):
    rng = np.random.default_rng(seed)
    pitch = (n + 15) // 16 * 16
    bq = np.zeros(ncol * pitch + 64, np.uint8)
    mq = np.full(ncol * pitch + 64, 60, np.uint8)
    for c in range(ncol):
        bq[c * pitch:c * pitch + n] = rng.integers(q[0], q[1], n)
    nt = np.zeros((ncol, 4), np.int32)
    nt[:, 0] = n - k
    nt[:, 1] = k
    return dict(col_off=np.arange(ncol + 1, dtype=np.int64) * pitch, nt_cnt=nt, ref_base=np.full(ncol, ord("A"), np.uint8),
                bq=bq, mq=mq, baq=None, sq=None, coverage=None)


c = lofreq_b200.Caller(0)
capi.check(c.lib.lfb200_set_profiling(c._ctx, 1))
for (ncol, n, k) in [(1, 7000, 3000), (148, 7000, 3000), (296, 7000, 3000), (148, 10000, 5000), (148, 4200, 2100), (148, 16000, 8000), (148, 7000, 1500), (148, 7000, 600)]:
    b = batch(ncol, n, k)
    for rep in range(2):
        t0 = time.perf_counter()
        out = c.call_columns(b, None, dense=False)
        dt = time.perf_counter() - t0
    ms = (C.c_float * 4)()
    capi.check(c.lib.lfb200_get_profile(c._ctx, ms))
    print("cols %4d n %6d K %5d: heavy phase %.3f ms (front %.3f), call %.1f ms, sites %d, kernels %s" % (ncol, n, k, ms[3], ms[0], dt * 1e3, out["n_sites"], out["job_counts"]))
