// The banded glocal profile HMM behind BAQ (kpa_ext_glocal, kprobaln_ext.c:80-277; called once per read by
// bam_prob_realn_core_ext, bam_md_ext.c:407), one read per thread, written so that every floating-point operation happens in
// the reference's order and precision: posterior state and quality per base come out identical.
//
//   states per (query base i, reference base k): M (match), I (insertion), D (deletion); a band of half-width bw around
//   the diagonal; forward rows scaled to sum 1 (s[i]), backward rows scaled by the same factors; the posterior of the
//   best M / I cell of a row gives state[i] and q[i] = phred(1 - posterior).
//
// What differs from the reference is where the numbers live: the reference callocs two (l_query + 1) x (6 bw + 9) matrices
// per read and relies on their zeros outside the band; here a cell outside a row's band is never read (the zero the
// reference would find is supplied instead), the forward matrix goes to a scratch buffer interleaved over the 32 reads of
// a warp (baq.cu: KpaDevMem — the threads of a warp touch consecutive addresses),
// rows are stored unscaled and scaled when read (no rewrite pass), and the backward pass keeps two rows and folds the
// posterior pass into itself, so only one matrix is ever stored.
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

    // Rows are stored unscaled; the reference's rescaling (row 1: a division by its sum, later rows: a multiplication by the
    // inverse of their sum, kprobaln_ext.c:148,170; backward rows: by the inverse of the forward sum, :222) is applied when a
    // cell is read — the same operation on the same operands gives the same bits, and a row is written once and read once
    // instead of being rewritten in place.
    auto fscale = [&](int i, double c, double inv) -> double { return i == 1 ? KPA_DIV(c, inv) : KPA_MUL(c, inv); };   // inv: s[1] for row 1

    // ---- forward
    mem.S(0) = 1.;
    double inv_prev;                                          // scale operand of the row before the current one
    {   // row 1: from the start state
        const KpaBand r1 = kpa_band(1, bw, l_ref);
        const double q1 = qual(1);
        double sum = 0.;
        for (int k = r1.beg; k <= r1.end; ++k) {
            const double em = KPA_MUL(kpa_emit(ref[k], query[1], q1), bM), ei = KPA_MUL(KPA_EI, bI);
            mem.F(1, kpa_u(r1.x, k, 0)) = em;
            mem.F(1, kpa_u(r1.x, k, 1)) = ei;
            mem.F(1, kpa_u(r1.x, k, 2)) = 0.;
            sum = KPA_ADD(sum, KPA_ADD(em, ei));
        }
        mem.S(1) = sum;
        inv_prev = sum;
    }
    for (int i = 2; i <= l_query; ++i) {
        const KpaBand r = kpa_band(i, bw, l_ref), p = kpa_band(i - 1, bw, l_ref);
        const double qli = qual(i);
        const int qyi = query[i];
        double sum = 0., dM = 0., dD = 0.;                  // M and D of position k - 1 of this row (zero below the band)
        // M, I, D of positions k - 1 and k of the row above: carried from one position to the next, and the cells of
        // position k + 1 are requested before the cells of position k of this row are stored (the rows live in one buffer,
        // so the compiler would not move the loads across the stores itself)
        // (cells outside the band of the row above come back as zeros and stay zeros under the scaling: a row's sum is
        // positive — its insertion cells are — so the factor is finite)
        auto above = [&](int k, double &v0, double &v1, double &v2) {
            v0 = v1 = v2 = 0.;
            if (k >= p.beg && k <= p.end) {
                v0 = mem.F(i - 1, kpa_u(p.x, k, 0));
                v1 = mem.F(i - 1, kpa_u(p.x, k, 1));
                v2 = mem.F(i - 1, kpa_u(p.x, k, 2));
            }
        };
        double a0, a1, a2, b0, b1, b2;
        above(r.beg - 1, a0, a1, a2);
        above(r.beg, b0, b1, b2);
        a0 = fscale(i - 1, a0, inv_prev); a1 = fscale(i - 1, a1, inv_prev); a2 = fscale(i - 1, a2, inv_prev);
        for (int k = r.beg; k <= r.end; ++k) {
            double n0, n1, n2;
            above(k + 1, n0, n1, n2);
            b0 = fscale(i - 1, b0, inv_prev); b1 = fscale(i - 1, b1, inv_prev); b2 = fscale(i - 1, b2, inv_prev);
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
            a0 = b0; a1 = b1; a2 = b2;
            b0 = n0; b1 = n1; b2 = n2;
        }
        mem.S(i) = sum;
        inv_prev = KPA_DIV(1., sum);
    }
    const KpaBand rl = kpa_band(l_query, bw, l_ref);
    {   // into the end state
        double sum = 0.;
        for (int k = rl.beg; k <= rl.end; ++k)
            sum = KPA_ADD(sum, KPA_ADD(KPA_MUL(fscale(l_query, mem.F(l_query, kpa_u(rl.x, k, 0)), inv_prev), sM),
                                       KPA_MUL(fscale(l_query, mem.F(l_query, kpa_u(rl.x, k, 1)), inv_prev), sI)));
        mem.S(l_query + 1) = sum;
    }

    // ---- backward, with the posterior of a row taken as soon as the row exists
    // binv: the factor the backward row is to be read with (1 / s[i]; the last row is stored as it is: scaled == false)
    auto posterior = [&](int i, int par, const KpaBand &r, bool scaled, double binv) {
        const double finv = i == 1 ? mem.S(1) : KPA_DIV(1., mem.S(i));
        double sum = 0., mx = 0.;
        int max_k = -1;
        // (the four cells of position k + 1 are requested while position k is evaluated)
        double f0 = mem.F(i, kpa_u(r.x, r.beg, 0)), f1 = mem.F(i, kpa_u(r.x, r.beg, 1));
        double b0 = mem.B(par, kpa_u(r.x, r.beg, 0)), b1 = mem.B(par, kpa_u(r.x, r.beg, 1));
        for (int k = r.beg; k <= r.end; ++k) {
            double nf0 = 0., nf1 = 0., nb0 = 0., nb1 = 0.;
            if (k < r.end) {
                nf0 = mem.F(i, kpa_u(r.x, k + 1, 0)); nf1 = mem.F(i, kpa_u(r.x, k + 1, 1));
                nb0 = mem.B(par, kpa_u(r.x, k + 1, 0)); nb1 = mem.B(par, kpa_u(r.x, k + 1, 1));
            }
            double z = KPA_MUL(fscale(i, f0, finv), scaled ? KPA_MUL(b0, binv) : b0);
            if (z > mx) { mx = z; max_k = (k - 1) << 2 | 0; }
            sum = KPA_ADD(sum, z);
            z = KPA_MUL(fscale(i, f1, finv), scaled ? KPA_MUL(b1, binv) : b1);
            if (z > mx) { mx = z; max_k = (k - 1) << 2 | 1; }
            sum = KPA_ADD(sum, z);
            f0 = nf0; f1 = nf1; b0 = nb0; b1 = nb1;
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
        for (int k = rl.beg; k <= rl.end; ++k) {
            mem.B(par, kpa_u(rl.x, k, 0)) = vM;
            mem.B(par, kpa_u(rl.x, k, 1)) = vI;
            mem.B(par, kpa_u(rl.x, k, 2)) = 0.;
        }
        posterior(l_query, par, rl, false, 1.);
    }
    bool nscaled = false;                                    // does the row below (i + 1) carry a deferred factor?
    double ninv = 1.;
    for (int i = l_query - 1; i >= 1; --i) {
        const KpaBand r = kpa_band(i, bw, l_ref), n = kpa_band(i + 1, bw, l_ref);
        const double y = i > 1 ? 1. : 0., qli1 = qual(i + 1);
        const int qyi1 = query[i + 1];
        const int np = par ^ 1;
        double dD = 0.;                                      // D of position k + 1 of this row (zero above the band)
        for (int k = r.end; k >= r.beg; --k) {
            double c11 = 0., c10i = 0.;
            if (k + 1 >= n.beg && k + 1 <= n.end) {
                c11 = mem.B(par, kpa_u(n.x, k + 1, 0));
                if (nscaled) c11 = KPA_MUL(c11, ninv);
            }
            if (k >= n.beg && k <= n.end) {
                c10i = mem.B(par, kpa_u(n.x, k, 1));
                if (nscaled) c10i = KPA_MUL(c10i, ninv);
            }
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
        par = np;
        nscaled = true;
        ninv = KPA_DIV(1., mem.S(i));
        posterior(i, par, r, true, ninv);
    }
    return bw;
}

}  // namespace lfb
