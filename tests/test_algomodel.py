"""CPU check of the arithmetic the kernels use (tests/algomodel.py: linear-space recurrence, odds form,
power-of-two rescaling, saddlepoint tilt, the clamp rule) against the golden vectors of the reference."""
import os

import numpy as np

import algomodel as M
from test_oracle import GOLD


def test_model_reproduces_the_reference_grid():
    z = np.load(os.path.join(GOLD, "snpcaller_grid.npz"))
    offs = z["offsets"]
    worst = 0.0
    for i, name in enumerate(z["names"].tolist()):
        if z["counts"][i].max() * (offs[i + 1] - offs[i]) > 400_000:
            continue                      # keep the pure-Python model quick
        ep = z["err_probs"][offs[i]:offs[i + 1]]
        pv, st, _ = M.model_snpcaller(ep, z["counts"][i], int(z["bonf"][i]), float(z["sig"][i]))
        assert np.array_equal(st, z["status"][i]), name
        for j in range(3):
            if st[j] == 0:
                a, b = float(np.log(pv[j])), float(z["lnp"][i][j])
                worst = max(worst, abs(a - b) / max(abs(b), 1.0))
    assert worst < 1e-10


def test_model_random_triallelic_against_port(port_oracle):
    rng = np.random.default_rng(5)
    sig = float(np.float32(0.01))
    for _ in range(60):
        n = int(rng.integers(20, 1200))
        q = rng.integers(6, 60, n) if rng.random() < 0.5 else np.where(rng.random(n) < 0.3, 3, rng.integers(20, 41, n))
        ep = np.sort(10.0 ** (-q / 10.0))
        k1 = int(rng.integers(1, n + 1))
        k2 = int(rng.integers(0, min(k1, n - k1) + 1))
        k3 = int(rng.integers(0, min(5, n - k1 - k2) + 1))
        counts = [k1, k2, k3]
        rng.shuffle(counts)
        bonf = int(rng.integers(1, 10 ** 7))
        want = port_oracle.snpcaller(ep, counts, bonf, sig)
        wst = np.zeros(3, np.uint8)
        wst[want == M.LDBL_MAX] = 1
        wst[want == M.LDBL_MIN] = 2
        pv, st, _ = M.model_snpcaller(ep, counts, bonf, sig)
        assert np.array_equal(st, wst), (n, counts, bonf)
        for j in range(3):
            if st[j] == 0:
                a, b = float(np.log(pv[j])), float(np.log(want[j]))
                assert abs(a - b) <= 1e-10 * max(abs(b), 1.0)


def binom_close(got, want, rtol=1e-10):
    """relative agreement on a probability; below 1e-290 both must simply be that small"""
    if want < 1e-290:
        return got < 1e-280
    return abs(got - want) <= rtol * want


def test_binom_model_reproduces_the_reference():
    """the arithmetic of csrc/binom.cu (mass function summed on the short side of the mode) against the outputs of
    the reference's binom() over cdflib's incomplete beta function"""
    z = np.load(os.path.join(GOLD, "binom_grid.npz"))
    for i in range(len(z["status"])):
        st, cum, ccum = M.model_binom(int(z["num_trials"][i]), int(z["num_success"][i]), float(z["prob_success"][i]))
        assert st == int(z["status"][i]), i
        if st == 0:
            assert binom_close(cum, float(z["cdf"][i])) and binom_close(ccum, float(z["sf"][i])), \
                (i, int(z["num_trials"][i]), int(z["num_success"][i]), float(z["prob_success"][i]))


def test_packed_model_against_port(port_oracle):
    """the packed kernels' arithmetic (tests/algomodel.py: packed_column): fp32 Newton sums, secondary alleles read off a
    tilted row, and the conservative early exit — which may only ever fire on columns the reference leaves at LDBL_MAX"""
    rng = np.random.default_rng(17)
    sig = float(np.float32(0.01))
    fired = kept = 0
    for it in range(70):
        n = int(rng.integers(40, 900))
        K = int(rng.integers(9, min(256, n) + 1))
        hi_q = it % 2 == 0                         # high qualities: far tails, tilted rows
        qv = rng.integers(33, 41, n) if hi_q else rng.integers(8, 30, n)
        ep = 10.0 ** (-qv / 10.0)                  # pileup order, not sorted (the device does not sort)
        c2 = int(rng.integers(0, min(K, n - K) + 1)) if it % 3 else 0
        c3 = int(rng.integers(0, min(4, n - K - c2) + 1))
        counts = [K, c2, c3]
        rng.shuffle(counts)
        bonf = int(rng.choice([3, 3000, 3_000_000]))
        want = port_oracle.snpcaller(np.sort(ep), counts, bonf, sig)
        dead, lnp, fl, blocks = M.packed_column(ep, counts, bonf, sig)
        if dead:
            fired += 1
            assert np.all(want == M.LDBL_MAX), (n, counts, bonf)      # pruned columns are insignificant in the reference
            assert blocks <= (n + 31) // 32
            continue
        kept += 1
        pv, st = M.host_finish(counts, bonf, sig, False, lnp, fl)
        wst = np.zeros(3, np.uint8)
        wst[want == M.LDBL_MAX] = 1
        wst[want == M.LDBL_MIN] = 2
        assert np.array_equal(st, wst), (n, counts, bonf, lnp)
        for j in range(3):
            if st[j] == 0:
                a, b = float(np.log(pv[j])), float(np.log(want[j]))
                assert abs(a - b) <= 1e-10 * max(abs(b), 1.0), (n, counts, j, a, b)
    assert fired >= 5 and kept >= 20, (fired, kept)
