// k_packed: the O(depth*K) recurrence for columns with 8 < K <= 256, several columns per warp.
//
// The recurrence over reads is strictly serial (snpcaller.c:865-966: row n needs row n-1), so a column can use
// at most K lanes x registers of parallelism, and a warp that owns one K = 40 column spends most of its issue
// slots on the per-read bookkeeping (parameter load, lane-boundary shuffle, loop) instead of on cells.  Here a
// warp owns 32/G columns, G = 4, 8, 16 or 32 lanes each, PK_R = 8 cells per lane, and all its columns advance
// read by read in lock step: one parameter load + one shuffle + PK_R DFMAs per step serve every column of the warp.
//
// Two kernels, because the two halves want different occupancies:
// k_pk_prep (one warp per column, few registers, many warps per SM — every step is a latency-bound sweep):
//   1. walks the quality bytes of the column with 16-byte loads: plp_to_errprobs per read (snpcaller.c:399-491),
//      merged probabilities to an L2-resident scratch row (-1 = filtered out), and the sums the later steps need
//      (reads kept, sum p, sum ln q);
//   2. when the tail is further out than the untilted fp64 range (Chernoff exponent > 300 nats) solves the
//      saddlepoint equation for the tilt s by Newton on ln s.
// k_packed (fp64-pipe-bound), per warp task:
//      and leaves the per-read parameters (o, 1/q), o = p*s/q, in the row.
//   3. the recurrence in odds form, E[k] += E[k-1]*o, absorbing state T = T/q + E[K-1]*o, exact power-of-two
//      rescaling per group every 32 reads; the parameters of the next 32 reads of every column travel from the
//      rows to shared memory with cp.async while the current block runs;
//   4. ln P(X >= K), and the tails of the other alleles from the same (possibly tilted) row:
//      P(X >= c) = sum_{k >= c} E[k] s^-k — all terms positive.
// Columns the packed form cannot finish (parameters above 2^20, tail outside the untilted range after the fact,
// a low cell of a strongly tilted row lost to underflow) are appended to the fallback list, which k_heavy<8> takes next.
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "internal.h"
#include "dev_common.cuh"

namespace lfb {

template <int G>
__device__ __forceinline__ double group_sum(double v)
{
#pragma unroll
    for (int m = G / 2; m >= 1; m >>= 1) v += __shfl_xor_sync(FULL, v, m);
    return v;
}

template <int G>
__device__ __forceinline__ int group_sum_i(int v)
{
#pragma unroll
    for (int m = G / 2; m >= 1; m >>= 1) v += __shfl_xor_sync(FULL, v, m);
    return v;
}

template <int G>
__device__ __forceinline__ int group_max_i(int v)
{
#pragma unroll
    for (int m = G / 2; m >= 1; m >>= 1) v = max(v, __shfl_xor_sync(FULL, v, m));
    return v;
}

template <int G>
__device__ __forceinline__ int group_min_i(int v)
{
#pragma unroll
    for (int m = G / 2; m >= 1; m >>= 1) v = min(v, __shfl_xor_sync(FULL, v, m));
    return v;
}

// first task of every packed list in processing order (deepest bin first, widest group first): entry i belongs to
// list PK_NL-1-i; `per_task` columns per task (0: as many as the list's group width allows)
__device__ __forceinline__ void list_bases(const Workspace &ws, unsigned *tbase, bool per_column)
{
    unsigned acc = 0;
    for (int i = 0; i < PK_NL; ++i) {
        tbase[i] = acc;
        const int li = PK_NL - 1 - i;
        const unsigned nj = min(ws.counters->n_pjobs[li], (unsigned)ws.pcap);
        const unsigned ncol = per_column ? 1u : (8u >> (li & 3));
        acc += (nj + ncol - 1) / ncol;
    }
    tbase[PK_NL] = acc;
}

// saddlepoint tilt: ln s with sum_n o_n s/(1 + o_n s) = min(K, N - 1/2), o_n = p_n/q_n (whole warp, one column).
// The sums steer a root finder whose tolerance is coarse (see below), so they are taken in fp32.
__device__ double warp_newton(const double2 *row, int n, int K, int N, double lam)
{
    const int lane = lane_id();
    const double kt = fmin((double)K, (double)N - 0.5);
    const double s0 = kt * fmax((double)N - lam, 1e-300) / (fmax(lam, 1e-300) * fmax((double)N - kt, 0.5));
    double lo = 0.0, hi = 60.0;
    double ls = fmin(log(fmax(s0, 1.0)), hi);
    for (int it = 0; it < 40; ++it) {
        const float sf = (float)exp(ls);
        float g0 = 0.f, g1 = 0.f, d0 = 0.f, d1 = 0.f;
        int pos = lane;
        for (; pos + 32 < n; pos += 64) {
            const float a = (float)row[pos].x * sf, c = (float)row[pos + 32].x * sf;
            const float wa = __fdividef(a, 1.f + a), wc = __fdividef(c, 1.f + c);      // o s/(1 + o s); filtered reads have o = 0
            g0 += wa; g1 += wc;
            d0 = fmaf(wa, 1.f - wa, d0); d1 = fmaf(wc, 1.f - wc, d1);                   // derivative with respect to ln s
        }
        if (pos < n) {
            const float a = (float)row[pos].x * sf;
            const float wa = __fdividef(a, 1.f + a);
            g0 += wa;
            d0 = fmaf(wa, 1.f - wa, d0);
        }
        const double gsum = __shfl_sync(FULL, warp_sum((double)g0 + (double)g1), 0) - kt;
        const double d = __shfl_sync(FULL, warp_sum((double)d0 + (double)d1), 0);
        // an error e in ln s costs about d*e^2/2 nats of head-room (of ~700): stop once that is negligible
        const double step = d > 0.0 ? gsum / d : 0.0;
        if (d > 0.0 && fabs(step) * sqrt(fmax(d, 1.0)) < 0.5) {
            ls = fmin(fmax(ls - step, 0.0), 60.0);
            break;
        }
        if (gsum > 0.0) hi = fmin(hi, ls); else lo = fmax(lo, ls);
        double nl = d > 0.0 ? ls - step : 0.5 * (lo + hi);
        if (!(nl > lo && nl < hi)) nl = 0.5 * (lo + hi);
        ls = nl;
    }
    return ls;
}

// running sums of k_pk_prep's first sweep
struct PrepAcc {
    int N;
    int hi_o, hi_rq;       // largest o and 1/q seen, as the high words of the doubles (positive: ordered like the values)
    double lam, lq, qp;
};

__device__ __forceinline__ void prep_take(PrepAcc &a, double jp, bool ok, bool may_be_one, double2 &out)
{
    out = make_double2(0.0, 1.0);                      // neutral step: a filtered read leaves the row unchanged
    if (!ok) return;
    double p, q;
    if (may_be_one) {
        guard_pq(jp, p, q);
    } else {                                           // jp <= 0.96: only the p-guard of snpcaller.c:872-881 can fire
        p = jp < DEPS ? DEPS : jp;
        q = 1.0 - jp;
    }
    const double rq = 1.0 / q;
    const double o = p * rq;
    out = make_double2(o, rq);
    a.lam += p;
    ++a.N;
    a.qp *= q;
    a.hi_o = max(a.hi_o, __double2hiint(o));
    a.hi_rq = max(a.hi_rq, __double2hiint(rq));
}

constexpr int PREP_WARPS = 8;

__global__ void __launch_bounds__(32 * PREP_WARPS, 4) k_pk_prep(const __grid_constant__ DevConf cf, const __grid_constant__ DevBatch b,
                                                                const Lut *lut, const Workspace ws)
{
    __shared__ double s_lut[768];
    __shared__ unsigned s_tbase[PK_NL + 1];
    load_lut(s_lut, lut);
    if (threadIdx.x == 0) list_bases(ws, s_tbase, true);
    __syncthreads();
    const unsigned total = s_tbase[PK_NL];
    const int lane = lane_id();
    // the configuration treats reference and alt reads alike (the defaults): no per-read position bookkeeping;
    // bytes outside the column are zeroed, and bq 0 is filtered out
    const bool uniform = cf.min_bq >= 1 && cf.min_alt_bq <= cf.min_bq && cf.alt_bq_mode == 0 && !cf.def_alt_jq_on && !cf.jq_filters;
    const bool general = cf.use_baq | cf.use_sq;
    // jobs are dealt out statically (they cost about the same, and two atomics per column were a quarter of this
    // kernel's time); the row in the scratch pool was reserved by k_finalize
    const unsigned nwarps = gridDim.x * PREP_WARPS;
    for (unsigned t = blockIdx.x * PREP_WARPS + (threadIdx.x >> 5); t < total; t += nwarps) {
        const unsigned hit = __ballot_sync(FULL, s_tbase[lane] <= t && t < s_tbase[lane + 1]);
        const int oi = __ffs(hit) - 1;
        const int li = PK_NL - 1 - oi;
        const unsigned slot = t - s_tbase[oi];
        const long long c = ws.pjobs[(long long)li * ws.pcap + slot];
        PkInfo *info = ws.pinfo + (long long)li * ws.pcap + slot;
        Geom g;
        int cov;
        load_geom(b, c, g, cov);
        if (cf.alt_bq_mode == 1) g.alt_bp = cf.alt_bq_prob;
        const int K = max(ws.cnt6[6 * c], max(ws.cnt6[6 * c + 1], ws.cnt6[6 * c + 2]));
        const int npad = (g.n + 31) & ~31;             // rows are padded to 32 reads
        const long long off = info->scr_off;
        bool fb = false;
        double ln_s = 0.0, sum_lq = 0.0;
        if (!fb) {
            double2 *row = ws.pk_scratch + off;
            PrepAcc a;
            a.N = 0; a.lam = 0.0; a.lq = 0.0; a.qp = 1.0; a.hi_o = 0; a.hi_rq = 0;
            const long long abase = g.off & ~15ll;
            const int lead = (int)(g.off - abase);
            const int nch = (lead + g.n + 15) >> 4;
            for (int i = lane; i < nch; i += 32) {
                Chunk16 ch;
                load_chunk(cf, b, abase + 16ll * i, ch);
                const int pos0 = 16 * i - lead;
                const bool inside = pos0 >= 0 && pos0 + 16 <= g.n;
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    const unsigned wbq = w == 0 ? ch.bq.x : w == 1 ? ch.bq.y : w == 2 ? ch.bq.z : ch.bq.w;
                    const unsigned wmq = w == 0 ? ch.mq.x : w == 1 ? ch.mq.y : w == 2 ? ch.mq.z : ch.mq.w;
                    const unsigned wbaq = w == 0 ? ch.baq.x : w == 1 ? ch.baq.y : w == 2 ? ch.baq.z : ch.baq.w;
                    const unsigned wsq = w == 0 ? ch.sq.x : w == 1 ? ch.sq.y : w == 2 ? ch.sq.z : ch.sq.w;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int pos = pos0 + 4 * w + j;
                        const bool in_col = inside || (pos >= 0 && pos < g.n);
                        const int bq = (wbq >> (8 * j)) & 0xff, mq = (wmq >> (8 * j)) & 0xff;
                        double jp = 0.0;
                        bool ok;
                        if (uniform) {
                            ok = in_col && bq >= cf.min_bq;
                            const double bp = s_lut[bq];
                            const double mp = cf.use_mq ? s_lut[256 + mq] : 0.0;
                            if (general)
                                jp = merge4(cf.use_sq ? s_lut[512 + ((wsq >> (8 * j)) & 0xff)] : 0.0, mp,
                                            cf.use_baq ? s_lut[512 + ((wbaq >> (8 * j)) & 0xff)] : 0.0, bp);
                            else
                                jp = __dadd_rn(mp, __dmul_rn(__dsub_rn(1.0, mp), bp));    // sp = bap = 0: the dropped terms are exact
                        } else {
                            bool is_alt;
                            int slot_;
                            ok = in_col && eval_read<true>(cf, s_lut, g, pos, bq, mq, (wbaq >> (8 * j)) & 0xff, (wsq >> (8 * j)) & 0xff,
                                                           is_alt, slot_, jp);
                        }
                        double2 e;
                        prep_take(a, jp, ok, general || !uniform, e);
                        if (in_col) row[pos] = e;
                    }
                    if (a.qp < 1e-200) {               // four reads shrink the product by at most 2^-208
                        a.lq += log(a.qp);
                        a.qp = 1.0;
                    }
                }
            }
            for (int pos = g.n + lane; pos < npad; pos += 32) row[pos] = make_double2(0.0, 1.0);
            a.lq += log(a.qp);
            const int N = __reduce_add_sync(FULL, a.N);
            const double lam = __shfl_sync(FULL, warp_sum(a.lam), 0);
            sum_lq = __shfl_sync(FULL, warp_sum(a.lq), 0);
            // upper bounds of the largest o and 1/q of the column
            const double max_o = __hiloint2double(__reduce_max_sync(FULL, a.hi_o), (int)0xffffffff);
            const double max_rq = __hiloint2double(__reduce_max_sync(FULL, a.hi_rq), (int)0xffffffff);
            __syncwarp();                              // the row is read by other lanes from here on
            // Chernoff exponent of the tail: beyond ~300 nats the untilted cells of interest drift out of fp64 range
            const double cher = ((double)K > lam) ? ((double)K * log((double)K / lam) - (double)K + lam) : 0.0;
            double s = 1.0;
            if (cher > 300.0) {
                ln_s = warp_newton(row, g.n, K, N, lam);
                s = exp(ln_s);
                if (ln_s == 0.0) s = 1.0;
            }
            // between two rescalings (32 reads) a cell may grow by (1 + o)^32 and the absorbing state by (1/q)^32
            if (max_o * s > 1048576.0 || max_rq > 1048576.0) {
                fb = true;
            } else if (s != 1.0) {
#pragma unroll 4
                for (int pos = lane; pos < g.n; pos += 32) row[pos].x *= s;
            }
        }
        if (lane == 0) {
            if (fb) {
                info->scr_off = -1;
                const unsigned sl = atomicAdd(&ws.counters->n_jobs[CLS_FALLBACK], 1u);
                ws.jobs[(long long)CLS_FALLBACK * ws.cap_cols + sl] = (int)c;
            } else {
                info->ln_s = ln_s;
                info->sum_lq = sum_lq;
            }
        }
    }
}

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}

__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <int G>
__device__ void packed_task(const DevConf &cf, const DevBatch &b, const Workspace &ws, const int *list, const PkInfo *infos,
                            unsigned j0, unsigned nj, double2 *par)
{
    constexpr int NCOL = 32 / G;
    constexpr int R = PK_R;
    constexpr int PBUF = 33 * 8;                       // one parameter buffer: 8 rows of 32 reads, padded (banks)
    const int lane = lane_id(), grp = lane / G, gl = lane % G;
    const int last = grp * G + G - 1;                  // the lane that owns the top cells and the absorbing state
    bool have = j0 + grp < nj;
    const long long c = have ? list[j0 + grp] : -1;
    int npad = 0;
    int cnt[3] = {0, 0, 0};
    long long bonf = 1;
    double ln_s = 0.0, sum_lq = 0.0;
    const double2 *myrow = ws.pk_scratch;
    if (have) {
        const PkInfo inf = infos[j0 + grp];
        if (inf.scr_off < 0) {
            have = false;                              // k_pk_prep sent it to the per-column kernels
        } else {
            const int4 nc = reinterpret_cast<const int4 *>(b.nt_cnt)[c];
            npad = (nc.x + nc.y + nc.z + nc.w + 31) & ~31;
            myrow += inf.scr_off;
            ln_s = inf.ln_s;
            sum_lq = inf.sum_lq;
#pragma unroll
            for (int i = 0; i < 3; ++i) cnt[i] = ws.cnt6[6 * c + i];
            bonf = ws.bonf_used[c];
        }
    }
    const int K = max(cnt[0], max(cnt[1], cnt[2]));

    // ---- 3. the recurrence, all columns of the warp in lock step
    const int k0 = K - G * R + gl * R;                 // cell of register 0 (k < 0: padding, stays 0)
    double E[R], T = 0.0;
#pragma unroll
    for (int r = 0; r < R; ++r) E[r] = (have && k0 + r == 0) ? 1.0 : 0.0;
    int e2 = 0;
    const int nmax = __reduce_max_sync(FULL, npad);
    // parameters (o, 1/q) of 32 reads of every column: cp.async from the rows k_pk_prep wrote, double-buffered
    auto stage = [&](int n0, double2 *buf) {
#pragma unroll
        for (int ci = 0; ci < NCOL; ++ci) {
            const int n_c = __shfl_sync(FULL, npad, ci * G);
            const double2 *row = reinterpret_cast<const double2 *>(__shfl_sync(FULL, (unsigned long long)myrow, ci * G));
            double2 *dst = buf + ci * 33 + lane;
            if (n0 < n_c) cp_async16(dst, row + n0 + lane);
            else *dst = make_double2(0.0, 1.0);        // this column has no reads left: neutral steps
        }
    };
    // Early exit (the reference's, snpcaller.c:916-958, in a form that needs no running sum of ln q): the tail over the
    // reads seen so far can only grow, and the product of the q of ALL reads is a lower bound of the product over the
    // reads seen, so  T * 2^e2 * exp(sum_lq) * s^-K  <=  P(X >= K).  Once that bound passes sig / bonf the column is
    // insignificant whatever follows.  Tested on the exponent of T alone (floor(log2 T) <= log2 T), once per block.
    const double thr_ln = log(cf.sig * (1.0 + 1e-9) / (double)bonf) - sum_lq + (double)K * ln_s;
    bool dead = !have;
    stage(0, par);
    int cur = 0;
    for (int n0 = 0; n0 < nmax; n0 += 32, cur ^= 1) {
        cp_async_wait_all();
        __syncwarp();
        if (n0 + 32 < nmax) stage(n0 + 32, par + (cur ^ 1) * PBUF);
        const double2 *pp = par + cur * PBUF + grp * 33;
        // software-pipelined: the parameters of read j+1 and the boundary cell for read j+1 (the top cell right
        // after its own update) are requested before the remaining R-1 cells of read j are updated
        double2 c_next = pp[0];
        double in_next = __shfl_up_sync(FULL, E[R - 1], 1);
#pragma unroll 4
        for (int j = 0; j < 32; ++j) {
            const double2 cc = c_next;
            const double in = gl == 0 ? 0.0 : in_next;
            c_next = pp[(j + 1) & 31];
            const double top = E[R - 1];
            T = fma(top, cc.x, T * cc.y);
            E[R - 1] = fma(E[R - 2], cc.x, top);
            in_next = __shfl_up_sync(FULL, E[R - 1], 1);
#pragma unroll
            for (int r = R - 2; r >= 1; --r) E[r] = fma(E[r - 1], cc.x, E[r]);
            E[0] = fma(in, cc.x, E[0]);
        }
        // exact power-of-two rescaling, per group
        int hi = 0;
#pragma unroll
        for (int r = 0; r < R; ++r) hi = max(hi, __double2hiint(E[r]));
        if (lane == last) hi = max(hi, __double2hiint(T));
        hi = group_max_i<G>(hi);
        const int ex = (hi >> 20) - 1023;
        if (hi > 0 && (ex > 200 || ex < -200)) {
            const double f = __hiloint2double((1023 - ex) << 20, 0);
#pragma unroll
            for (int r = 0; r < R; ++r) E[r] *= f;
            T *= f;
            e2 += ex;
        }
        const bool over = lane == last && (double)(((__double2hiint(T) >> 20) - 1023) + e2) * LN2 > thr_ln;
        const bool over_g = __shfl_sync(FULL, (int)over, last) != 0;     // every lane takes part, dead or not
        dead = dead || over_g;
        __syncwarp();                                  // this buffer is refilled during the next block
        if (__all_sync(FULL, dead || n0 + 32 >= npad)) break;      // nothing left to decide in this warp
    }
    cp_async_wait_all();
    bool fb = false;

    // ---- 4. tails
    int hiE = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) hiE = max(hiE, __double2hiint(E[r]));
    hiE = group_max_i<G>(hiE);
    const double Tl = __shfl_sync(FULL, T, last);
    const double topl = __shfl_sync(FULL, E[R - 1], last);
    const int hiT = __double2hiint(Tl);
    const int peak = max(hiE, hiT) >> 20;
    const int gap = peak - (hiT >> 20);
    if ((gap > 580 && ln_s == 0.0) || gap > 900) fb = true;     // needs the tilt after all / out of range: k_heavy<R> decides
    const double base = (double)e2 * LN2 + sum_lq;
    const double lnT = log(Tl) + base - (double)K * ln_s;
    const double lnKm1 = log(topl) + base - (double)(K - 1) * ln_s;
    bool site = have && !fb && !dead;
    if (site && lnT > -700.0 && exp(lnT) * (double)bonf > cf.sig * (1.0 + 1e-9)) site = false;   // snpcaller.c:1155
    double lnp[3] = {0.0, 0.0, 0.0};
    const double invs = (ln_s == 0.0) ? 1.0 : exp(-ln_s);
#pragma unroll 1
    for (int i = 0; i < 3; ++i) {
        const int ci = cnt[i];
        if (!__any_sync(FULL, site && ci > 0 && ci < K)) {
            if (ci == K) lnp[i] = lnT;
            continue;
        }
        const bool mine = site && ci > 0 && ci < K;
        // P(X >= ci) = sum_{k >= ci} E[k] s^-(k-ci) + T s^-(K-ci), times s^-ci and the common scale
        double acc = 0.0;
        int hc = 0x7fffffff;
        if (mine) {
            const int kk = max(k0, ci);
            double f = (ln_s == 0.0) ? 1.0 : exp(-(double)(kk - ci) * ln_s);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (k0 + r >= ci) {
                    acc = fma(E[r], f, acc);
                    f *= invs;
                }
                if (k0 + r == ci) hc = __double2hiint(E[r]);
            }
            if (lane == last) acc = fma(T, (ln_s == 0.0) ? 1.0 : exp(-(double)(K - ci) * ln_s), acc);
        }
        acc = group_sum<G>(acc);
        hc = group_min_i<G>(hc);
        if (mine) {
            // the leading cell must have stayed a normal number all along (rescaling keeps the peak within 2^+-200
            // at the block boundaries and below 2^840 inside a block)
            if ((hc >> 20) < 64 || peak - (hc >> 20) > 850) fb = true;
            lnp[i] = log(acc) + base - (double)ci * ln_s;
        } else if (ci == K) {
            lnp[i] = lnT;
        }
    }
    if (dead) fb = false;                              // ruled out for good: nothing for the fallback list to decide
    if (fb) site = false;
    if (have && gl == 0) {
        if (fb) {
            const unsigned slot = atomicAdd(&ws.counters->n_jobs[CLS_FALLBACK], 1u);
            ws.jobs[(long long)CLS_FALLBACK * ws.cap_cols + slot] = (int)c;
        } else if (site) {
            Cand cd;
            cd.col = c;
            cd.bonf = bonf;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                cd.lnp[i] = cnt[i] > 0 ? lnp[i] : 0.0;
                cd.cnt[i] = cnt[i];
                cd.raw[i] = ws.cnt6[6 * c + 3 + i];
            }
            cd.ln_floor = fmin(lnT, lnKm1);
            cd.flags = 0;
            cd.pad = 0;
            ws.cand[atomicAdd(&ws.counters->n_cand, 1u)] = cd;
            mark_cand(ws, c);
        }
    }
    __syncwarp();
}

constexpr int PK_WARPS = 4;

__global__ void __launch_bounds__(32 * PK_WARPS) k_packed(const __grid_constant__ DevConf cf, const __grid_constant__ DevBatch b,
                                                          const Workspace ws)
{
    __shared__ double2 s_par[PK_WARPS][2 * 33 * 8];
    __shared__ unsigned s_tbase[PK_NL + 1];
    if (threadIdx.x == 0) list_bases(ws, s_tbase, false);
    __syncthreads();
    const unsigned total = s_tbase[PK_NL];
    const int lane = lane_id(), wib = threadIdx.x >> 5;
    // the first task of every warp is dealt out statically — with about as many tasks as resident warps, a race for
    // them leaves some SMs with six tasks per sub-partition and others with three — the rest dynamically
    const unsigned nwarps = gridDim.x * PK_WARPS;
    unsigned t = blockIdx.x * PK_WARPS + wib;
    for (;; ) {
        if (t >= total) break;
        const unsigned hit = __ballot_sync(FULL, s_tbase[lane] <= t && t < s_tbase[lane + 1]);
        const int oi = __ffs(hit) - 1;
        const int li = PK_NL - 1 - oi;
        const int gi = li & 3;
        const unsigned nj = min(ws.counters->n_pjobs[li], (unsigned)ws.pcap);
        const unsigned j0 = (t - s_tbase[oi]) * (8u >> gi);
        const int *list = ws.pjobs + (long long)li * ws.pcap;
        const PkInfo *infos = ws.pinfo + (long long)li * ws.pcap;
        switch (gi) {
            case 0: packed_task<4>(cf, b, ws, list, infos, j0, nj, s_par[wib]); break;
            case 1: packed_task<8>(cf, b, ws, list, infos, j0, nj, s_par[wib]); break;
            case 2: packed_task<16>(cf, b, ws, list, infos, j0, nj, s_par[wib]); break;
            default: packed_task<32>(cf, b, ws, list, infos, j0, nj, s_par[wib]); break;
        }
        if (lane == 0) t = nwarps + atomicAdd(&ws.counters->next_ptask, 1u);
        t = __shfl_sync(FULL, t, 0);
    }
}

constexpr int PK_CTAS_PER_SM = 6;

void launch_packed(const LaunchState &ls, const DevConf &cf, const DevBatch &b, const Lut *lut, const Workspace &ws, cudaStream_t st)
{
    if (b.n_cols <= 0 || !ws.pk_scratch) return;
    k_pk_prep<<<ls.sms * 4, 32 * PREP_WARPS, 0, st>>>(cf, b, lut, ws);
    k_packed<<<ls.sms * PK_CTAS_PER_SM, 32 * PK_WARPS, 0, st>>>(cf, b, ws);
}

}  // namespace lfb
