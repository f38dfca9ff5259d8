"""Generates tests/golden/binom_grid.npz from the UNMODIFIED reference binom() (binom.c:52-93 over the vendored
cdflib90) compiled into oracle/_ref/libbinomref.so (oracle/Makefile).  Run where /root/reference is mounted:

    python tests/golden/make_golden_binom.py

The reference ships no runnable test for binom() (tests/uniq.sh needs absent data, tests/binom_vs_poisson.FIXME is
a stub), so the fixture holds outputs of the reference itself: the call `lofreq uniq` makes
(coverage, alt_count, af; lofreq_uniq.c:381) over a grid of depths and frequencies, random problems, both tails, the
degenerate corners (pr = 0, pr = 1, s = 0, s = n) and every argument check cdfbin(which=1) performs.
"""
import ctypes as C
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.pyoracle import BinomRef  # noqa: E402


def cases():
    out = [(500, 3, 0.05), (100, 50, 0.5), (10000, 120, 0.01), (30, 0, 0.2), (1, 0, 0.3), (5, 5, 0.3), (10, 3, 0.0),
           (10, 3, 1.0), (0, 0, 0.5), (-3, 0, 0.5), (10, 11, 0.5), (10, -1, 0.5), (10, 3, 1.5), (10, 3, -0.1),
           (1000000, 5000, 0.01), (1000000, 4000, 0.005), (1000, 999, 0.999), (1000, 0, 1e-9), (20, 19, 0.5),
           (2, 1, 0.5), (15, 7, 0.5), (16, 8, 0.5), (35, 1, 0.9), (36, 35, 0.1), (80, 40, 0.3), (81, 2, 0.3),
           (500, 250, 0.5), (501, 499, 0.5)]
    # what lofreq uniq asks: coverage x alt_count at the other sample's allele frequency
    for cov in (10, 37, 100, 500, 2000, 10000, 100000):
        for af in (0.005, 0.01, 0.05, 0.2, 0.5, 0.9):
            for frac in (0.0, 0.25, 0.5, 1.0, 1.5, 3.0):
                out.append((cov, min(cov, int(round(frac * af * cov))), af))
    rng = np.random.default_rng(20261017)
    for _ in range(1500):
        n = int(10 ** rng.uniform(0, 5.5))
        p = float(rng.uniform(0, 1)) if rng.random() < 0.5 else float(10 ** rng.uniform(-6, 0))
        if rng.random() < 0.5:
            s = int(rng.integers(0, n + 1))
        else:
            s = min(n, max(0, int(n * p + rng.normal() * 3 * math.sqrt(n * p * (1 - p) + 1))))
        out.append((n, s, p))
    return out


def main():
    br = BinomRef()
    cs = cases()
    n = np.array([c[0] for c in cs], np.int32)
    s = np.array([c[1] for c in cs], np.int32)
    p = np.array([c[2] for c in cs], np.float64)
    status = np.zeros(len(cs), np.int32)
    cdf = np.full(len(cs), np.nan)
    sf = np.full(len(cs), np.nan)
    for i in range(len(cs)):
        a, b = C.c_double(float("nan")), C.c_double(float("nan"))
        status[i] = br.lib.binom(C.byref(a), C.byref(b), int(n[i]), int(s[i]), float(p[i]))
        cdf[i], sf[i] = a.value, b.value
    np.savez_compressed(os.path.join(HERE, "binom_grid.npz"), num_trials=n, num_success=s, prob_success=p,
                        status=status, cdf=cdf, sf=sf)
    print("binom_grid:", len(cs), "cases,", int((status != 0).sum()), "refused by cdfbin")


if __name__ == "__main__":
    main()
