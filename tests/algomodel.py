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
