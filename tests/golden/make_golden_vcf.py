"""Generates tests/golden/vcf_boundary.npz from the callback-boundary oracle (oracle/_ref/libcallref.so: the
reference's real call_vars / call_snvs / report_var / vcf_write_var / kt_fisher_exact compiled unmodified, driven by
the fake pileup of oracle/call_harness.c).  Run here, where /root/reference is mounted:

    python tests/golden/make_golden_vcf.py

Contents: raw VCF text + final counters for a few synthetic column batches (with strand counts), and a grid of DP4
tables with the SB the reference's report_var() computes for them."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import synth_np                      # noqa: E402
from oracle.pyoracle import CallOracle, default_conf   # noqa: E402

CASES = [  # name, workload, c0, n, with_baq, conf overrides
    ("c2", "C2", 0, 6000, False, {}),
    ("c4_baq", "C4", 123456, 6000, True, {}),
    ("c3", "C3", 77000, 400, False, {}),
    ("c5", "C5", 5000, 300, True, {}),
    ("c2_fixedbonf", "C2", 9000, 3000, False, dict(bonf_dynamic=0, bonf_subst=3000000)),
    ("c4_filters", "C4", 40000, 3000, True, dict(min_bq=3, min_alt_bq=20, min_jq=15, min_alt_jq=25, def_alt_jq=35)),
]


def sb_tables():
    rng = np.random.default_rng(2026)
    t = []
    for _ in range(3000):
        depth = int(rng.choice([10, 50, 300, 500, 2000, 10000]))
        alt = int(rng.integers(1, max(2, depth // 2)))
        ref = depth - alt
        bias = rng.choice([0.5, 0.5, 0.3, 0.1, 0.02])
        rf = int(rng.binomial(ref, 0.5))
        af = int(rng.binomial(alt, bias))
        t.append((rf, ref - rf, af, alt - af))
    t += [(0, 0, 5, 0), (0, 0, 0, 7), (0, 0, 3, 4), (10, 0, 0, 10), (500, 500, 0, 300), (1, 0, 0, 1), (0, 1, 1, 0), (3, 3, 3, 3),
          (5000, 5000, 5000, 0), (100000, 100000, 300, 0), (7, 0, 0, 0), (0, 0, 0, 0), (250, 250, 250, 250)]
    return np.array(t, np.int32)


def main():
    o = CallOracle()
    out = {}
    for name, wl, c0, n, baq, over in CASES:
        b = synth_np.generate(wl, c0, n, with_baq=baq, with_strand=True)
        txt, bonf, tests = o.call_vars_vcf(b, b["strand8"], default_conf(**over), pos=np.arange(c0, c0 + n) % 100000)
        out["vcf_" + name] = np.frombuffer(txt.encode(), np.uint8)
        out["counters_" + name] = np.array([bonf, tests], np.int64)
        print(name, len(txt.splitlines()), "records", bonf, tests)
    t = sb_tables()
    out["sb_tables"] = t
    out["sb_qual"] = np.array([o.sb_qual(*row)[0] for row in t], np.int64)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "vcf_boundary.npz"), **out)


if __name__ == "__main__":
    main()
