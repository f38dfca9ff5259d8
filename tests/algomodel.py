"""TEST INFRASTRUCTURE — numpy model of the arithmetic the CUDA kernels use
(lofreq_b200/csrc/snv_kernels.cu), so the maths (linear-space recurrence, odds
form with power-of-two rescaling, exponential tilting, the tail/clamp rule) can
be checked against the oracle on a CPU box before any GPU time is spent.  The
product never imports this.

The reference evaluates the Poisson-binomial DP in log space
(snpcaller.c:830-971).  The device evaluates the same recurrence in linear
space:

  small K (<= KS):   P[k] <- P[k] q + P[k-1] p ,  T <- T + P[K-1] p
                     on 32 disjoint read subsets, merged by truncated
                     convolution (all terms positive, no cancellation)
  large K:           odds form  E[k] <- E[k] + E[k-1] o ,  o = p s / q
                     T~ <- T~ / q + E[K-1] o ,  rescaled by 2^-e every 32 reads;
                     ln P[k] = ln E[k] + e ln 2 + sum ln q - k ln s
                     s = 1 unless the tail is too far out, then s solves
                     sum o/(1+o) = K (exponential tilting / saddlepoint)
"""
import math

import numpy as np

DBL_EPS = np.finfo(np.float64).eps
LN_DBL_MIN = -708.3964185322641       # glibc exp() raises FE_UNDERFLOW below this
LN_LDBL_MIN = -11355.137111933024     # glibc expl() raises FE_UNDERFLOW below this
KS = 8
LD = np.longdouble
LDBL_MAX = np.finfo(LD).max
LDBL_MIN = np.finfo(LD).tiny


def guards(ep):
    """p and q = 1-p with the reference's DBL_EPSILON guards (snpcaller.c:872-881)."""
    ep = np.asarray(ep, np.float64)
    p = np.where(np.abs(ep) < DBL_EPS, DBL_EPS, ep)
    q = np.where(np.abs(ep - 1.0) < DBL_EPS, 1.0 - ep + DBL_EPS, 1.0 - ep)
    return p, q


def small_k_dist(p, q, K, lanes=32):
    """distribution truncated at K: (P[0..K-1], T) via per-lane recurrences + tree merge."""
    n = len(p)
    parts = []
    chunk = 16
    # lane l owns reads [16 i, 16 i + 16) for i = l, l+32, ... (the device's 128-bit chunks)
    for l in range(lanes):
        P = np.zeros(K); P[0] = 1.0; T = 0.0
        for start in range(l * chunk, n, lanes * chunk):
            for j in range(start, min(start + chunk, n)):
                T = T + P[K - 1] * p[j]
                for k in range(K - 1, 0, -1):
                    P[k] = P[k] * q[j] + P[k - 1] * p[j]
                P[0] = P[0] * q[j]
        parts.append((P, T))
    step = 1
    while step < lanes:
        nxt = []
        for i in range(0, len(parts), 2):
            nxt.append(merge(parts[i], parts[i + 1], K))
        parts = nxt
        step *= 2
    return parts[0]


def merge(A, B, K):
    a, tA = A
    b, tB = B
    c = np.zeros(K)
    for k in range(K):
        for i in range(k + 1):
            c[k] += a[i] * b[k - i]
    totB = b.sum() + tB
    t = tA * totB + tB * a.sum()
    for j in range(1, K):
        t += b[j] * a[K - j:].sum()
    return c, t


def newton_tilt(p, q, K):
    """ln s with sum o/(1+o) = Kt, o = p s / q; Kt = min(K, N - 0.5)."""
    n = len(p)
    kt = min(float(K), n - 0.5)
    r = p / q
    lam = float(np.sum(p))
    s0 = kt * max(n - lam, 1e-300) / (max(lam, 1e-300) * (n - kt))
    ls = math.log(max(s0, 1.0))
    lo, hi = 0.0, 700.0
    for _ in range(40):
        o = r * math.exp(ls)
        w = o / (1.0 + o)
        g = float(np.sum(w)) - kt
        if g > 0:
            hi = min(hi, ls)
        else:
            lo = max(lo, ls)
        d = float(np.sum(w / (1.0 + o)))
        nl = ls - g / d if d > 0 else 0.5 * (lo + hi)
        if not (lo < nl < hi):
            nl = 0.5 * (lo + hi)
        if abs(nl - ls) < 1e-3:
            ls = nl
            break
        ls = nl
    return ls


def heavy_row(p, q, K, ln_s, chunk=32):
    """odds-form DP. Returns (lnrow[0..K-1], lnT)."""
    n = len(p)
    s = math.exp(ln_s)
    o = p * s / q
    rq = 1.0 / q
    E = np.zeros(K); E[0] = 1.0
    T = 0.0
    e2 = 0
    for c0 in range(0, n, chunk):
        for j in range(c0, min(c0 + chunk, n)):
            top = E[K - 1]
            E[1:] = E[1:] + E[:-1] * o[j]
            T = T * rq[j] + top * o[j]
        m = max(E.max(), T)
        _, ex = math.frexp(m)
        E = np.ldexp(E, -ex); T = math.ldexp(T, -ex); e2 += ex
    sum_lq = float(np.sum(np.log1p(-p) if False else np.log(q)))
    with np.errstate(divide="ignore"):
        lnrow = np.log(E) + e2 * math.log(2.0) + sum_lq - np.arange(K) * ln_s
        lnT = (math.log(T) if T > 0 else -np.inf) + e2 * math.log(2.0) + sum_lq - K * ln_s
    ratio = (math.log(T) if T > 0 else -np.inf) - math.log(max(E.max(), T))
    return lnrow, lnT, ratio


def chernoff_exponent(K, lam):
    return K * math.log(K / lam) - K + lam if K > lam else 0.0


def tail_K(p, q, K):
    """ln P(X >= K) and ln P(X = K-1) the way the device gets them."""
    if K <= KS:
        P, T = small_k_dist(p, q, K)
        with np.errstate(divide="ignore"):
            return math.log(T), (math.log(P[K - 1]) if P[K - 1] > 0 else -np.inf), (P, T)
    lam = float(np.sum(p))
    ln_s = 0.0
    if chernoff_exponent(K, lam) > 300.0:
        ln_s = newton_tilt(p, q, K)
    lnrow, lnT, ratio = heavy_row(p, q, K, ln_s)
    if ln_s == 0.0 and ratio < -400.0:
        ln_s = newton_tilt(p, q, K)
        lnrow, lnT, ratio = heavy_row(p, q, K, ln_s)
    assert ratio > -600.0, ratio
    return lnT, lnrow[K - 1], (lnrow, lnT, ln_s)


def device_column(ep, counts, bonf, sig):
    """What the device hands to the host for one tested column:
    decided_insig, lnp[3] (= ln P(X >= c_i)), ln_floor = min(ln P[K-1], ln P(>=K))."""
    counts = [int(c) for c in counts]
    K = max(counts)
    p, q = guards(ep)
    lnT, lnKm1, aux = tail_K(p, q, K)
    if lnT > -700 and math.exp(lnT) * float(bonf) > sig * (1.0 + 1e-9):
        return True, [0.0] * 3, 0.0
    lnp = [0.0] * 3
    for i, c in enumerate(counts):
        if c == 0:
            continue
        if c == K:
            lnp[i] = lnT
        elif K <= KS:
            P, T = aux
            lnp[i] = math.log(T + P[c:].sum())
        else:
            lnrow, _, ln_s = aux
            if ln_s == 0.0:
                m = max(lnrow[c:].max(), lnT)
                lnp[i] = m + math.log(np.exp(lnrow[c:] - m).sum() + math.exp(lnT - m))
            else:
                lnp[i] = tail_K(p, q, c)[0]
    return False, lnp, min(lnKm1, lnT)


def host_finish(counts, bonf, sig, decided_insig, lnp, ln_floor):
    """long double finishing on the host (lofreq_b200/csrc/host_finish.c), mirrors
    snpcaller.c:1144-1196: returns pvalues (longdouble[3]) and status[3]."""
    pv = np.full(3, LDBL_MAX, dtype=LD)
    st = np.ones(3, np.uint8)
    counts = [int(c) for c in counts]
    K = max(counts)
    if K == 0 or decided_insig:
        return pv, st
    lnT = lnp[counts.index(K)]
    pK = LDBL_MIN if lnT < LN_LDBL_MIN else np.exp(LD(lnT))
    if pK * LD(float(bonf)) > LD(sig):
        return pv, st
    for i, c in enumerate(counts):
        if c == 0:
            continue
        t = lnp[i]
        flagged = t < LN_LDBL_MIN
        if c < K and t - ln_floor > -LN_DBL_MIN:
            flagged = True
        pi = np.exp(LD(t))
        if flagged:
            if pi < LD(DBL_EPS):
                pv[i], st[i] = LDBL_MIN, 2
            else:
                pv[i], st[i] = LDBL_MAX, 1
        else:
            pv[i], st[i] = pi, 0
    return pv, st


def model_snpcaller(ep, counts, bonf, sig):
    if max(counts) == 0:
        return np.full(3, LDBL_MAX, dtype=LD), np.ones(3, np.uint8), [0.0] * 3
    d, lnp, fl = device_column(ep, counts, bonf, sig)
    pv, st = host_finish(counts, bonf, sig, d, lnp, fl)
    return pv, st, lnp


# ------------------------------------------------------------------------------------------------
# binom(): the arithmetic of lofreq_b200/csrc/binom.cu (k_binom), restated in Python.
# cdflib's cdfbin(which=1) (dcdflib.c:1727) reduces the binomial to the incomplete beta function (TOMS 708); the
# device instead sums the probability mass function on the short side of the mode: one accurately computed term
# (Loader's saddle-point form: Stirling-series error term + the deviance bd0) and the exact term ratio
# pmf(k-1)/pmf(k) = k/(n-k+1) * q/p from there on, all terms positive.
# ------------------------------------------------------------------------------------------------
LN_2PI = 1.8378770664093454835606594728112


def stirlerr_table():
    """stirlerr(n) = ln n! - ln(sqrt(2 pi n) (n/e)^n) for n = 0..15, in long double like the host does"""
    ld = np.longdouble
    out = [0.0]
    for n in range(1, 16):
        fact = ld(1)
        for i in range(2, n + 1):
            fact *= ld(i)
        v = np.log(fact) - (ld(n) + ld(0.5)) * np.log(ld(n)) + ld(n) - ld(0.5) * np.log(ld(2) * ld(np.pi))
        out.append(float(v))
    return out


_STIRL = None


def stirlerr(n):
    global _STIRL
    if _STIRL is None:
        _STIRL = stirlerr_table()
    if n <= 15:
        return _STIRL[n]
    S0, S1, S2, S3, S4 = 1.0 / 12, 1.0 / 360, 1.0 / 1260, 1.0 / 1680, 1.0 / 1188
    x = float(n)
    nn = x * x
    if n > 500:
        return (S0 - S1 / nn) / x
    if n > 80:
        return (S0 - (S1 - S2 / nn) / nn) / x
    if n > 35:
        return (S0 - (S1 - (S2 - S3 / nn) / nn) / nn) / x
    return (S0 - (S1 - (S2 - (S3 - S4 / nn) / nn) / nn) / nn) / x


def bd0(x, np_):
    """x ln(x/np) + np - x, without cancellation when x is close to np"""
    if abs(x - np_) < 0.1 * (x + np_):
        v = (x - np_) / (x + np_)
        s = (x - np_) * v
        ej = 2.0 * x * v
        v = v * v
        for j in range(1, 1000):
            ej *= v
            s1 = s + ej / (2 * j + 1)
            if s1 == s:
                return s1
            s = s1
    return x * math.log(x / np_) + np_ - x


def ln_binom_pmf(x, n, p, q):
    """ln of the binomial mass at x, 0 < p < 1"""
    if x == 0:
        return (-bd0(n, n * q) - n * p) if p < 0.1 else n * math.log(q)
    if x == n:
        return (-bd0(n, n * p) - n * q) if q < 0.1 else n * math.log(p)
    lc = stirlerr(n) - stirlerr(x) - stirlerr(n - x) - bd0(x, n * p) - bd0(n - x, n * q)
    lf = LN_2PI + math.log(x) + math.log1p(-x / n)
    return lc - 0.5 * lf


def model_binom(num_trials, num_success, pr):
    """(status, cdf, sf) like binom() (binom.c:52-93): P(X <= s), P(X > s) for X ~ Binomial(n, pr)"""
    n, s = int(num_trials), int(num_success)
    if not n > 0:
        return -5, None, None              # dcdflib.c:1879
    if s < 0 or s > n:
        return -4, None, None              # :1889
    if pr < 0.0 or pr > 1.0:
        return -6, None, None              # :1904
    if not s < n:
        return 0, 1.0, 0.0                 # cumbin, dcdflib.c:5017-5026
    q = 1.0 - pr
    if pr == 0.0:
        return 0, 1.0, 0.0
    if q == 0.0:
        return 0, 0.0, 1.0
    mode = math.floor((n + 1) * pr)
    if s < mode:
        # left tail: terms k = s, s-1, ... 0 shrink
        lt = ln_binom_pmf(s, n, pr, q)
        ratio = q / pr
        t, tot, k = 1.0, 1.0, s
        while k > 0:
            t *= (k / (n - k + 1.0)) * ratio
            tot += t
            k -= 1
            if t < 1e-18 * tot:
                break
        cum = math.exp(lt + math.log(tot))
        return 0, cum, 1.0 - cum
    lt = ln_binom_pmf(s + 1, n, pr, q)
    ratio = pr / q
    t, tot, k = 1.0, 1.0, s + 1
    while k < n:
        t *= ((n - k) / (k + 1.0)) * ratio
        tot += t
        k += 1
        if t < 1e-18 * tot:
            break
    ccum = math.exp(lt + math.log(tot))
    return 0, 1.0 - ccum, ccum


# ------------------------------------------------------------------------------------------------
# the fused recurrence kernel (lofreq_b200/csrc/dp_fused.cu: k_dp): what its tilt and lock-step form add to the arithmetic above
# ------------------------------------------------------------------------------------------------
def newton_tilt_fp32(p, q, K):
    """k_dp's root finder (group_tilt; there on a 64-bucket histogram of the probabilities): the sums that steer it are taken in fp32 (the tolerance is |e| sqrt(d) < 0.5),
    o = p/q per read, ln s capped at 60; one accepted step ends the iteration."""
    n = int(np.sum(np.ones_like(p)))
    kt = min(float(K), n - 0.5)
    lam = float(np.sum(p))
    s0 = kt * max(n - lam, 1e-300) / (max(lam, 1e-300) * max(n - kt, 0.5))
    lo, hi = 0.0, 60.0
    ls = min(math.log(max(s0, 1.0)), hi)
    o32 = (p / q).astype(np.float32)
    for _ in range(40):
        a = o32 * np.float32(math.exp(ls))
        w = a / (np.float32(1.0) + a)
        g = float(np.sum(w, dtype=np.float64)) - kt
        d = float(np.sum(w * (np.float32(1.0) - w), dtype=np.float64))
        step = g / d if d > 0 else 0.0
        if d > 0 and abs(step) * math.sqrt(max(d, 1.0)) < 0.5:
            return min(max(ls - step, 0.0), 60.0)
        if g > 0:
            hi = min(hi, ls)
        else:
            lo = max(lo, ls)
        nl = ls - step if d > 0 else 0.5 * (lo + hi)
        if not (lo < nl < hi):
            nl = 0.5 * (lo + hi)
        ls = nl
    return ls


def packed_column(ep, counts, bonf, sig, chunk=32):
    """One column the way k_dp evaluates it (8 < K <= 2048).  Returns
    (dead, lnp[3], ln_floor, blocks_run): dead = the conservative early exit fired, i.e. the column is insignificant."""
    counts = [int(c) for c in counts]
    K = max(counts)
    p, q = guards(ep)
    n = len(p)
    lam = float(np.sum(p))
    sum_lq = float(np.sum(np.log(q)))
    ln_s = newton_tilt_fp32(p, q, K) if chernoff_exponent(K, lam) > 300.0 else 0.0
    s = math.exp(ln_s) if ln_s != 0.0 else 1.0
    rq = 1.0 / q
    o = p * rq * s
    E = np.zeros(K); E[0] = 1.0
    T, e2 = 0.0, 0
    thr_ln = math.log(sig * (1.0 + 1e-9) / float(bonf)) - sum_lq + K * ln_s
    dead, blocks = False, 0
    for c0 in range(0, n, chunk):
        for j in range(c0, min(c0 + chunk, n)):
            top = E[K - 1]
            E[1:] = E[1:] + E[:-1] * o[j]
            T = T * rq[j] + top * o[j]
        blocks += 1
        m = max(E.max(), T)
        _, ex = math.frexp(m)
        ex -= 1                                   # the kernel takes the exponent field: m in [2^ex, 2^(ex+1))
        if ex > 200 or ex < -200:
            E = np.ldexp(E, -ex); T = math.ldexp(T, -ex); e2 += ex
        # early exit on the exponent of T alone: floor(log2 T) + e2 is a lower bound of log2 of the scaled tail, and
        # exp(sum_lq) (ALL reads) a lower bound of the product of q over the reads seen so far
        if T > 0 and ((math.frexp(T)[1] - 1) + e2) * math.log(2.0) > thr_ln:
            dead = True
            break
    if dead:
        return True, [0.0] * 3, 0.0, blocks
    base = e2 * math.log(2.0) + sum_lq
    lnT = math.log(T) + base - K * ln_s
    lnKm1 = math.log(E[K - 1]) + base - (K - 1) * ln_s
    lnp = [0.0] * 3
    invs = math.exp(-ln_s)
    for i, c in enumerate(counts):
        if c == 0:
            continue
        if c == K:
            lnp[i] = lnT
            continue
        # the other alleles off the same (possibly tilted) row: P(X >= c) = sum_{k >= c} E[k] s^-(k-c) + T s^-(K-c), times s^-c
        f, acc = 1.0, 0.0
        for k in range(c, K):
            acc += E[k] * f
            f *= invs
        acc += T * math.exp(-(K - c) * ln_s)
        lnp[i] = math.log(acc) + base - c * ln_s
    return False, lnp, min(lnT, lnKm1), blocks
