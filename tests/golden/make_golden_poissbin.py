"""poissbin() rows from the UNMODIFIED reference (oracle/_ref/libsnpref.so): tests/golden/poissbin_rows.npz.
Run where /root/reference is mounted:  python tests/golden/make_golden_poissbin.py
Cases: unpruned rows (bonf = sig = 1 can never prune... it does when p > 1, so sig is large), rows cut short by the
Bonferroni-aware early exit (snpcaller.c:916-958), unsorted input (lofreq_uniq.c:299-312 does not sort), the
source_qual call shape (plp.c:554: bonf 1, sig 0.05), the DBL_EPSILON guards, K == N, K == 1."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.pyoracle import Oracle  # noqa: E402


def cases():
    rng = np.random.default_rng(4242)
    sig = float(np.float32(0.01))
    out = []
    for n, k in ((10, 1), (10, 10), (50, 3), (200, 17), (500, 25), (500, 250), (1000, 40), (2000, 300), (300, 299)):
        q = rng.integers(20, 41, n)
        ep = np.sort(10.0 ** (-q / 10.0))
        out.append(("full_n%d_k%d" % (n, k), ep, k, 1, 1e300))              # never pruned
        out.append(("sig_n%d_k%d" % (n, k), ep, k, 3_000_000, sig))         # pruned iff insignificant
        out.append(("unsorted_n%d_k%d" % (n, k), rng.permutation(ep), k, 30000, sig))
    for n, k in ((100, 1), (100, 2), (150, 4), (250, 7), (76, 3)):          # source_qual shape: a read's mismatches
        ep = 10.0 ** (-rng.integers(2, 41, n) / 10.0)
        out.append(("srcqual_n%d_k%d" % (n, k), np.sort(ep), k, 1, 0.05))
    out.append(("guards", np.array([0.0, 1e-20, 1.0, 1.0 - 1e-17, 0.5, 1e-3, 0.25, 0.125]), 3, 1, 1e300))
    out.append(("prune_first_step", np.full(400, 0.03), 1, 3, sig))
    out.append(("q30_k1_b3e6", np.full(500, 1.000999e-3), 1, 3_000_000, sig))
    out.append(("q30_k4_b3e6", np.full(500, 1.000999e-3), 4, 3_000_000, sig))
    out.append(("deep_n6000_k700", np.sort(10.0 ** (-rng.integers(20, 41, 6000) / 10.0)), 700, 3_000_000, sig))
    return out


def main():
    ref = Oracle("reference")
    cs = cases()
    names, eps, offs, ks, bonfs, sigs, rows, roffs, pvs = [], [], [0], [], [], [], [], [0], []
    for name, ep, k, bonf, sg in cs:
        pv, row = ref.poissbin(ep, k, bonf, sg)
        names.append(name); eps.append(ep); offs.append(offs[-1] + len(ep)); ks.append(k); bonfs.append(bonf); sigs.append(sg)
        rows.append(row); roffs.append(roffs[-1] + len(row))
        pvs.append(np.frombuffer(np.array([pv], np.longdouble).tobytes(), np.uint8))
    np.savez_compressed(os.path.join(HERE, "poissbin_rows.npz"), names=np.array(names), err_probs=np.concatenate(eps),
                        offsets=np.array(offs, np.int64), k=np.array(ks, np.int32), bonf=np.array(bonfs, np.int64),
                        sig=np.array(sigs, np.float64), rows=np.concatenate(rows), row_offsets=np.array(roffs, np.int64),
                        pvalue_ld=np.stack(pvs))
    print("wrote %d poissbin cases, %d row cells" % (len(cs), roffs[-1]))


if __name__ == "__main__":
    main()
