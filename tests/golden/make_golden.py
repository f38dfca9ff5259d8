"""Generates tests/golden/*.npz from the UNMODIFIED reference compiled into
oracle/_ref/libsnpref.so (see oracle/Makefile).  Run in the build container,
where /root/reference is mounted:

    python tests/golden/make_golden.py

The fixtures pin (1) snpcaller() on the Q x N x K grid the reference authors
sketched at lofreq_call.c:1032-1050 plus tri-allelic / FE-clamp / K==N / tiny-p
edge cases, and (2) the whole per-column path
(call_vars/call_snvs minus VCF output) on small synthetic column batches.
long double p-values are stored as their raw 16-byte images (exact) plus
ln p as float64 for readability.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.pyoracle import Oracle, default_conf, VARCALL_USE_MQ, VARCALL_USE_BAQ  # noqa: E402
from oracle import synth_np  # noqa: E402


def snpcaller_cases():
    """list of (name, err_probs, counts, bonf, sig)"""
    rng = np.random.default_rng(1234)
    sig = float(np.float32(0.01))
    cases = []
    # the one known-answer value in the reference (snpcaller.c:1222-1232)
    cases.append(("kat_10x0.001_k1", np.full(10, 0.001), (1, 0, 0), 1, 1.0))
    for q in (20, 25, 30, 35, 40):
        p = 10.0 ** (-q / 10.0)
        for n in (10, 50, 100, 500, 1000, 5000):
            for frac in (0.0, 0.002, 0.01, 0.05, 0.25, 0.5, 1.0):
                k = max(1, min(n, int(round(frac * n)))) if frac > 0 else 1
                if n * k > 1_500_000:
                    continue
                bonf = 3 * n * 10
                cases.append((f"grid_q{q}_n{n}_k{k}", np.full(n, p), (k, 0, 0), bonf, sig))
    # mixed qualities, sorted ascending as call_snvs does (lofreq_call.c:784)
    for n, ks in ((300, (3, 1, 0)), (500, (25, 2, 1)), (2000, (40, 4, 2)), (2000, (217, 3, 1)),
                  (2000, (1000, 300, 5)), (2000, (1000, 900, 0)), (5000, (2500, 7, 0))):
        q = rng.integers(20, 41, size=n)
        ep = np.sort(10.0 ** (-q / 10.0) + 1e-6 - 1e-6 * 10.0 ** (-q / 10.0))
        cases.append((f"mixed_n{n}_k{'_'.join(map(str, ks))}", ep, ks, 3_000_000, sig))
    # FE-clamp quirk: tri-allelic with a wide spread
    p30 = np.full(500, 10.0 ** -3 + 1e-6 - 1e-9)
    for ks in ((250, 20, 0), (250, 5, 0), (250, 249, 248), (250, 100, 1), (400, 1, 1), (500, 499, 1)):
        cases.append((f"clamp_n500_k{'_'.join(map(str, ks))}", p30, ks, 1_500_000, sig))
    # ln p below ln(LDBL_MIN): LDBL_MIN sentinel from poissbin itself
    cases.append(("tiny_n3000_q20_k3000", np.full(3000, 0.01), (3000, 0, 0), 3_000_000, sig))
    cases.append(("tiny_n4000_q30_k2500", np.full(4000, 0.001), (2500, 10, 0), 3_000_000, sig))
    # K == N, N == 1, K == 1
    cases.append(("k_eq_n_5", np.full(5, 0.001), (5, 0, 0), 300, sig))
    cases.append(("n1_k1", np.array([1e-4]), (1, 0, 0), 3, sig))
    cases.append(("k1_n2000", np.full(2000, 1e-9), (1, 1, 1), 3000, sig))
    # significance straddling the Bonferroni threshold
    for bonf in (1, 3, 300, 30000, 3_000_000, 300_000_000):
        cases.append((f"straddle_b{bonf}", np.full(500, 0.001), (4, 2, 0), bonf, sig))
    # p with guards: p < DBL_EPSILON and p close to 1
    cases.append(("guard_small_p", np.sort(np.r_[np.full(50, 1e-20), np.full(50, 1e-3)]), (3, 0, 0), 300, sig))
    cases.append(("guard_p_one", np.sort(np.r_[np.full(20, 1e-3), np.full(3, 1.0)]), (6, 1, 0), 30, sig))
    # random small cases
    for i in range(60):
        n = int(rng.integers(1, 400))
        q = rng.integers(6, 60, size=n)
        ep = np.sort(10.0 ** (-q / 10.0))
        c = np.sort(rng.integers(0, max(1, n // 3) + 1, size=3))[::-1]
        if c.sum() > n:
            c = np.array([min(n, c[0]), 0, 0])
        if c.max() == 0:
            c[0] = 1
        cases.append((f"rand{i}", ep, tuple(int(x) for x in c[rng.permutation(3)]), int(rng.integers(1, 10**7)), sig))
    return cases


def ld_bytes(a):
    a = np.ascontiguousarray(a, dtype=np.longdouble)
    return a.view(np.uint8).reshape(a.shape + (16,)).copy()


def status_of(pv):
    """0 = value, 1 = LDBL_MAX (not computed / insignificant), 2 = LDBL_MIN (clamped)"""
    mx, mn = np.finfo(np.longdouble).max, np.finfo(np.longdouble).tiny
    st = np.zeros(pv.shape, np.uint8)
    st[pv == mx] = 1
    st[pv == mn] = 2
    return st


def main():
    ref = Oracle("reference")
    # ---- snpcaller grid ----
    names, eps, offs, counts, bonfs, sigs, pvs = [], [], [0], [], [], [], []
    for name, ep, ks, bonf, sig in snpcaller_cases():
        pv = ref.snpcaller(ep, ks, bonf, sig)
        names.append(name); eps.append(np.asarray(ep, np.float64)); offs.append(offs[-1] + len(ep))
        counts.append(ks); bonfs.append(bonf); sigs.append(sig); pvs.append(pv)
    pvs = np.array(pvs, dtype=np.longdouble)
    with np.errstate(all="ignore"):
        lnp = np.log(pvs).astype(np.float64)
    np.savez_compressed(os.path.join(HERE, "snpcaller_grid.npz"), names=np.array(names),
                        err_probs=np.concatenate(eps), offsets=np.array(offs, np.int64),
                        counts=np.array(counts, np.int32), bonf=np.array(bonfs, np.int64),
                        sig=np.array(sigs, np.float64), pvalue_ld=ld_bytes(pvs), lnp=lnp,
                        status=status_of(pvs))
    print("snpcaller_grid:", len(names), "cases")

    # ---- whole-path batches ----
    for tag, wl, c0, n, baq, conf in (
            ("c2", "C2", 0, 1500, False, default_conf()),
            ("c3", "C3", 5000, 300, False, default_conf()),
            ("c4baq", "C4", 100, 1500, True, default_conf()),
            ("c5", "C5", 40, 200, True, default_conf(flag=VARCALL_USE_MQ | VARCALL_USE_BAQ)),
            ("c2_fixedbonf", "C2", 0, 1500, False, default_conf(bonf_dynamic=0, bonf_subst=4500)),
            ("c4_filters", "C4", 7, 800, True, default_conf(min_bq=25, min_alt_bq=28, min_jq=22, min_alt_jq=24,
                                                            def_alt_jq=30, min_cov=10)),
            ("c4_defaltbq", "C4", 7, 800, True, default_conf(def_alt_bq=20)),
            ("c4_medianbq", "C4", 7, 800, True, default_conf(def_alt_bq=-1)),
            ("c4_nomq", "C4", 7, 800, True, default_conf(flag=VARCALL_USE_BAQ)),
    ):
        b = synth_np.generate(wl, c0, n, with_baq=baq)
        out = ref.call_columns(b, dict(conf))
        with np.errstate(all="ignore"):
            lnp = np.log(out["pvalues"]).astype(np.float64)
        np.savez_compressed(
            os.path.join(HERE, f"batch_{tag}.npz"), workload=wl, c0=c0, n_cols=n, with_baq=baq,
            conf_keys=np.array(list(conf.keys())), conf_vals=np.array([float(v) for v in conf.values()]),
            alt_counts=out["alt_counts"], alt_raw_counts=out["alt_raw_counts"], tested=out["tested"],
            bonf_used=out["bonf_used"], pvalue_ld=ld_bytes(out["pvalues"]), lnp=lnp,
            status=status_of(out["pvalues"]), called=out["called"], qual=out["qual"],
            bonf_subst=out["bonf_subst"], num_snv_tests=out["num_snv_tests"],
            bq_sha=np.frombuffer(__import__("hashlib").sha256(b["bq"].tobytes()).digest(), np.uint8))
        print(f"batch_{tag}: tested={int(out['tested'].sum())} called={int(out['called'].sum())}"
              f" tests={out['num_snv_tests']}")


if __name__ == "__main__":
    main()
