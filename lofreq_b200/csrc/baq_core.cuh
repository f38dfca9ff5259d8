// The banded glocal profile HMM behind BAQ (kpa_ext_glocal, kprobaln_ext.c:80-277; called once per read by
// bam_prob_realn_core_ext, bam_md_ext.c:407), one read per thread, written so that every floating-point operation happens in
// the reference's order and precision: posterior state and quality per base come out identical.
//
//   states per (query base i, reference base k): M (match), I (insertion), D (deletion); a band of half-width bw around
//   the diagonal; forward rows scaled to sum 1 (s[i]), backward rows scaled by the same factors; the posterior of the
//   best M / I cell of a row gives state[i] and q[i] = phred(1 - posterior).
//
// What differs from the reference is where the numbers live: the reference callocs two (l_query + 1) x (6 bw + 9) matrices
// per read and relies on their zeros outside the band; here a cell outside a row's band is never read (ld_* return the zero
// the reference would find), the forward matrix goes to a scratch buffer interleaved over the reads of a launch (cell c of
// row i of read t at ((i * W3 + c) * stride + t: the threads of a warp touch consecutive addresses), and the backward
// pass keeps two rows and folds the posterior pass into itself, so only one matrix is ever stored.
//
// The same source is compiled for the host by tests/test_baq_core.py (plain g++, -ffp-contract=off) to pin the arithmetic
// against the compiled reference without a GPU; the product only ever runs the device instance.
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define KPA_HD __host__ __device__ __forceinline__
#else
#define KPA_HD static inline
#endif

#if defined(__CUDA_ARCH__)
#define KPA_MUL(a, b) __dmul_rn((a), (b))
#define KPA_ADD(a, b) __dadd_rn((a), (b))
#define KPA_SUB(a, b) __dsub_rn((a), (b))
#define KPA_DIV(a, b) __ddiv_rn((a), (b))
#define KPA_FSUB(a, b) __fsub_rn((a), (b))
#define KPA_FDIV(a, b) __fdiv_rn((a), (b))
#else
#define KPA_MUL(a, b) ((a) * (b))
#define KPA_ADD(a, b) ((a) + (b))
#define KPA_SUB(a, b) ((a) - (b))
#define KPA_DIV(a, b) ((a) / (b))
#define KPA_FSUB(a, b) ((float)(a) - (float)(b))
#define KPA_FDIV(a, b) ((float)(a) / (float)(b))
#endif

namespace lfb {

constexpr double KPA_EI = .25;                 // emission of an inserted base (kprobaln_ext.c:41)
constexpr double KPA_EM = .33333333333;        // emission of a mismatch, per alternative base (:42)

// a base whose quality needs the host's logarithm: -4.343 ln(x) + .499 lies within the guard band of an integer
struct KpaFix {
    long long base;        // index into state[] / q[]
    double x;              // 1 - posterior
};

// band of row i: reference positions [beg, end], and x = max(i - bw, 0), the position the row's storage starts below
struct KpaBand {
    int beg, end, x;
};

KPA_HD KpaBand kpa_band(int i, int bw, int l_ref)
{
    KpaBand b;
    b.x = i - bw > 0 ? i - bw : 0;
    b.beg = 1 > i - bw ? 1 : i - bw;
    b.end = l_ref < i + bw ? l_ref : i + bw;
    return b;
}

// storage index of state s of reference position k in a row whose storage starts at x (set_u, kprobaln_ext.c:46)
KPA_HD int kpa_u(int x, int k, int s) { return (k - x + 1) * 3 + s; }

KPA_HD double kpa_emit(int r, int qy, double ql)
{
    // (ref[k] > 3 || query[i] > 3) ? 1 : ref[k] == query[i] ? 1 - qual[i] : qual[i] * EM   (kprobaln_ext.c:140,159)
    return (r > 3 || qy > 3) ? 1. : r == qy ? KPA_SUB(1., ql) : KPA_MUL(ql, KPA_EM);
}

// Mem: F(i, c) -> double& forward cell c of row i; B(p, c) -> double& backward cell c of the row buffer p (0 / 1);
// S(i) -> double& scaling factor i (0 .. l_query + 1).
// Returns the band half-width used; state / q get l_query entries; *n_fix counts guard-band bases appended to fix[].
template <class Mem>
KPA_HD int kpa_glocal_core(const uint8_t *ref0, int l_ref, const uint8_t *qry0, int l_query, const uint8_t *iqual, float cd, float ce,
                           int cbw, const float *q2p, Mem &mem, int *state, uint8_t *q, long long base0, KpaFix *fix, int fix_cap,
                           unsigned *n_fix)
{
    if (l_ref <= 0 || l_query <= 0) return 0;
    const uint8_t *ref = ref0 - 1, *query = qry0 - 1;          // 1-based like the reference
    int bw = l_ref > l_query ? l_ref : l_query;
    if (bw > cbw) bw = cbw;
    const int dl = l_ref > l_query ? l_ref - l_query : l_query - l_ref;
    if (bw < dl) bw = dl;
    // transition probabilities (kprobaln_ext.c:124-129): the terms in c->d / c->e are float arithmetic there
    const double sM = KPA_DIV(1., (double)(2 * l_query + 2)), sI = sM;
    const float one_d = KPA_FSUB(1.f, cd), one_e = KPA_FSUB(1.f, ce);
    const double oms = KPA_SUB(1., sM);
    const double m0 = KPA_MUL((double)KPA_FSUB(one_d, cd), oms), m1 = KPA_MUL((double)cd, oms), m2 = m1;
    const double m3 = KPA_MUL((double)one_e, KPA_SUB(1., sI)), m4 = KPA_MUL((double)ce, KPA_SUB(1., sI));
    const double m6 = (double)one_e, m8 = (double)ce;
    const double bM = (double)KPA_FDIV(one_d, (float)l_ref), bI = (double)KPA_FDIV(cd, (float)l_ref);
    auto qual = [&](int i) -> double { return (double)q2p[iqual ? iqual[i - 1] : 30]; };

    // ---- forward
    mem.S(0) = 1.;
    {   // row 1: from the start state
        const KpaBand r1 = kpa_band(1, bw, l_ref);
        const int end = l_ref < bw + 1 ? l_ref : bw + 1;
        const double q1 = qual(1);
        double sum = 0.;
        for (int k = 1; k <= end; ++k) {
            const double em = KPA_MUL(kpa_emit(ref[k], query[1], q1), bM), ei = KPA_MUL(KPA_EI, bI);
            mem.F(1, kpa_u(r1.x, k, 0)) = em;
            mem.F(1, kpa_u(r1.x, k, 1)) = ei;
            mem.F(1, kpa_u(r1.x, k, 2)) = 0.;
            sum = KPA_ADD(sum, KPA_ADD(em, ei));
        }
        mem.S(1) = sum;
        for (int k = 1; k <= end; ++k)
            for (int s = 0; s < 3; ++s) {
                double &c = mem.F(1, kpa_u(r1.x, k, s));
                c = KPA_DIV(c, sum);
            }
    }
    for (int i = 2; i <= l_query; ++i) {
        const KpaBand r = kpa_band(i, bw, l_ref), p = kpa_band(i - 1, bw, l_ref);
        // row 1 holds positions 1 .. min(l_ref, bw + 1) (its own rule), later rows their band
        const int pend = i - 1 == 1 ? (l_ref < bw + 1 ? l_ref : bw + 1) : p.end, pbeg = i - 1 == 1 ? 1 : p.beg;
        const double qli = qual(i);
        const int qyi = query[i];
        double sum = 0., dM = 0., dD = 0.;                  // M and D of position k - 1 of this row (zero below the band)
        for (int k = r.beg; k <= r.end; ++k) {
            double a0 = 0., a1 = 0., a2 = 0., b0 = 0., b1 = 0.;
            if (k - 1 >= pbeg && k - 1 <= pend) {
                a0 = mem.F(i - 1, kpa_u(p.x, k - 1, 0));
                a1 = mem.F(i - 1, kpa_u(p.x, k - 1, 1));
                a2 = mem.F(i - 1, kpa_u(p.x, k - 1, 2));
            }
            if (k >= pbeg && k <= pend) {
                b0 = mem.F(i - 1, kpa_u(p.x, k, 0));
                b1 = mem.F(i - 1, kpa_u(p.x, k, 1));
            }
            const double e = kpa_emit(ref[k], qyi, qli);
            const double fm = KPA_MUL(e, KPA_ADD(KPA_ADD(KPA_MUL(m0, a0), KPA_MUL(m3, a1)), KPA_MUL(m6, a2)));
            const double fi_ = KPA_MUL(KPA_EI, KPA_ADD(KPA_MUL(m1, b0), KPA_MUL(m4, b1)));
            const double fd = KPA_ADD(KPA_MUL(m2, dM), KPA_MUL(m8, dD));
            mem.F(i, kpa_u(r.x, k, 0)) = fm;
            mem.F(i, kpa_u(r.x, k, 1)) = fi_;
            mem.F(i, kpa_u(r.x, k, 2)) = fd;
            sum = KPA_ADD(sum, KPA_ADD(KPA_ADD(fm, fi_), fd));
            dM = fm;
            dD = fd;
        }
        mem.S(i) = sum;
        const double inv = KPA_DIV(1., sum);
        for (int k = r.beg; k <= r.end; ++k)
            for (int s = 0; s < 3; ++s) {
                double &c = mem.F(i, kpa_u(r.x, k, s));
                c = KPA_MUL(c, inv);
            }
    }
    const KpaBand rl = kpa_band(l_query, bw, l_ref);
    const int lbeg = l_query == 1 ? 1 : rl.beg, lend = l_query == 1 ? (l_ref < bw + 1 ? l_ref : bw + 1) : rl.end;
    {   // into the end state
        double sum = 0.;
        for (int k = lbeg; k <= lend; ++k)
            sum = KPA_ADD(sum, KPA_ADD(KPA_MUL(mem.F(l_query, kpa_u(rl.x, k, 0)), sM), KPA_MUL(mem.F(l_query, kpa_u(rl.x, k, 1)), sI)));
        mem.S(l_query + 1) = sum;
    }

    // ---- backward, with the posterior of a row taken as soon as the row exists
    auto posterior = [&](int i, int par, const KpaBand &r, int rbeg, int rend) {
        double sum = 0., mx = 0.;
        int max_k = -1;
        for (int k = rbeg; k <= rend; ++k) {
            double z = KPA_MUL(mem.F(i, kpa_u(r.x, k, 0)), mem.B(par, kpa_u(r.x, k, 0)));
            if (z > mx) { mx = z; max_k = (k - 1) << 2 | 0; }
            sum = KPA_ADD(sum, z);
            z = KPA_MUL(mem.F(i, kpa_u(r.x, k, 1)), mem.B(par, kpa_u(r.x, k, 1)));
            if (z > mx) { mx = z; max_k = (k - 1) << 2 | 1; }
            sum = KPA_ADD(sum, z);
        }
        mx = KPA_DIV(mx, sum);
        state[i - 1] = max_k;
        // q = (int)(-4.343 log(1 - max) + .499), 99 above 100 (kprobaln_ext.c:262).  log(0) = -inf makes the conversion
        // overflow, which the reference's x86 build turns into INT_MIN and then into byte 0.
        const double x = KPA_SUB(1., mx);
        int kq;
        if (!(x > 0.)) {
            kq = 0;
        } else {
            const double v = KPA_ADD(KPA_MUL(-4.343, log(x)), .499);
            kq = (int)v;
            const double fr = v - floor(v);
            if ((fr < 1e-9 || fr > 1. - 1e-9) && fix) {         // the device's logarithm may round the other way: host decides
                const unsigned slot =
#if defined(__CUDA_ARCH__)
                    atomicAdd(n_fix, 1u);
#else
                    (*n_fix)++;
#endif
                if ((int)slot < fix_cap) { fix[slot].base = base0 + i - 1; fix[slot].x = x; }
            }
        }
        q[i - 1] = (uint8_t)(kq > 100 ? 99 : kq);
    };
    int par = 0;
    {   // row l_query: from the end state
        const double sl = mem.S(l_query), sl1 = mem.S(l_query + 1);
        const double vM = KPA_DIV(KPA_DIV(sM, sl), sl1), vI = KPA_DIV(KPA_DIV(sI, sl), sl1);
        for (int k = lbeg; k <= lend; ++k) {
            mem.B(par, kpa_u(rl.x, k, 0)) = vM;
            mem.B(par, kpa_u(rl.x, k, 1)) = vI;
            mem.B(par, kpa_u(rl.x, k, 2)) = 0.;
        }
        posterior(l_query, par, rl, lbeg, lend);
    }
    for (int i = l_query - 1; i >= 1; --i) {
        const KpaBand r = kpa_band(i, bw, l_ref), n = kpa_band(i + 1, bw, l_ref);
        // positions row i + 1 holds (the last row was written over [lbeg, lend], which is its band)
        const int nbeg = (i + 1 == l_query) ? lbeg : n.beg, nend = (i + 1 == l_query) ? lend : n.end;
        const double y = i > 1 ? 1. : 0., qli1 = qual(i + 1);
        const int qyi1 = query[i + 1];
        const int np = par ^ 1;
        double dD = 0.;                                      // D of position k + 1 of this row (zero above the band)
        for (int k = r.end; k >= r.beg; --k) {
            double c11 = 0., c10i = 0.;
            if (k + 1 >= nbeg && k + 1 <= nend) c11 = mem.B(par, kpa_u(n.x, k + 1, 0));
            if (k >= nbeg && k <= nend) c10i = mem.B(par, kpa_u(n.x, k, 1));
            const double em = k >= l_ref ? 0. : kpa_emit(ref[k + 1], qyi1, qli1);
            const double e = KPA_MUL(em, c11);
            const double bm = KPA_ADD(KPA_ADD(KPA_MUL(e, m0), KPA_MUL(KPA_MUL(KPA_EI, m1), c10i)), KPA_MUL(m2, dD));
            const double bi = KPA_ADD(KPA_MUL(e, m3), KPA_MUL(KPA_MUL(KPA_EI, m4), c10i));
            const double bd = KPA_MUL(KPA_ADD(KPA_MUL(e, m6), KPA_MUL(m8, dD)), y);
            mem.B(np, kpa_u(r.x, k, 0)) = bm;
            mem.B(np, kpa_u(r.x, k, 1)) = bi;
            mem.B(np, kpa_u(r.x, k, 2)) = bd;
            dD = bd;
        }
        const double inv = KPA_DIV(1., mem.S(i));
        const int rbeg = i == 1 ? 1 : r.beg, rend = i == 1 ? (l_ref < bw + 1 ? l_ref : bw + 1) : r.end;
        for (int k = r.beg; k <= r.end; ++k)
            for (int s = 0; s < 3; ++s) {
                double &c = mem.B(np, kpa_u(r.x, k, s));
                c = KPA_MUL(c, inv);
            }
        par = np;
        posterior(i, par, r, r.beg, r.end);
        (void)rbeg; (void)rend;
    }
    return bw;
}

}  // namespace lfb
