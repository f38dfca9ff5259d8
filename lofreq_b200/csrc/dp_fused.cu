// k_dp: the O(depth*K) recurrence of pruned_calc_prob_dist (snpcaller.c:830-971) for every column with
// 8 < K <= 2048, straight from the quality bytes — plp_to_errprobs (snpcaller.c:345-498), the tilt, the recurrence and
// the per-allele tails in ONE kernel, several columns per warp.
//
// The recurrence over reads is strictly serial (row n needs row n-1), so a column can use at most K lanes x registers
// of parallelism, and a warp that owns one K = 40 column would spend its issue slots on per-read bookkeeping instead
// of on cells.  A warp therefore owns 32/G columns, G = 4, 8, 16 or 32 lanes each, R = 8, 16, 32 or 64 cells per
// lane (G*R >= K), all of them advancing read by read in lock step: one parameter load + one shuffle + R DFMAs per
// step serve every column of the warp.  Job lists per (G, R, depth bin) keep the columns of a warp alike.
//
// Per warp task:
//   1. pre-pass over the column's bytes with 16-byte loads: reads kept and lambda = sum of the merged probabilities
//      (G lanes per column).  Chernoff exponent of the tail > 300 nats -> the untilted cells of interest would leave
//      the fp64 range -> saddlepoint tilt s: a histogram of the merged probabilities (160 buckets of 0.75 dB) built by
//      the whole warp in shared memory, Newton on ln s over the buckets (the tolerance of the tilt is coarse: an
//      error e costs about var*e^2/2 nats of ~700).
//   2. the quality bytes of the next 64 reads of every column of the warp travel from HBM to shared memory with
//      1-D bulk TMA copies (cp.async.bulk + mbarrier, double-buffered: stage k+1 is in flight while stage k is
//      consumed); for each block of 32 reads the lanes of a column turn its bytes into the step parameters
//      (o, 1/q), o = p*s/q, with the shared-memory copy of the glibc-pow LUT and the reference's merge order.
//   3. the recurrence in odds form, E[k] += E[k-1]*o (one DFMA per cell), absorbing state T = T/q + E[K-1]*o,
//      exact power-of-two rescaling per column every 32 reads, and the reference's early exit (snpcaller.c:916-958)
//      on a lower bound of ln P(X >= K among the reads seen) built from the exponents of T and of the running
//      product of q.
//   4. ln P(X >= K), and the tails of the other alleles from the same (possibly tilted) row:
//      P(X >= c) = sum_{k >= c} E[k] s^-k — all terms positive.
// Columns this form cannot finish (a parameter above 2^20, a tail that needs the tilt after all, a low cell of a
// strongly tilted row lost to underflow) go to the per-column fallback lists that k_heavy_all takes afterwards.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <math.h>
#include <stdlib.h>
#define LFB_EVAL_INLINE __noinline__      // one copy of the general read evaluation per kernel: see dlog below
#include "internal.h"
#include "dev_common.cuh"
#include "screen_common.cuh"

namespace lfb {

// One copy of the logarithm and the exponential for everything outside the recurrence: k_dp is four (G, R) instances of a long
// task, and a warp runs most of a task's code once — what it executes is bounded by instruction fetch, so the code the
// instances can share is kept out of line.
static __device__ __noinline__ double dlog(double x) { return log(x); }
static __device__ __noinline__ double dexp(double x) { return exp(x); }

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
    if constexpr (G == 32) {
        return __reduce_max_sync(FULL, v);
    } else {
#pragma unroll
        for (int m = G / 2; m >= 1; m >>= 1) v = max(v, __shfl_xor_sync(FULL, v, m));
        return v;
    }
}

template <int G>
__device__ __forceinline__ int group_min_i(int v)
{
#pragma unroll
    for (int m = G / 2; m >= 1; m >>= 1) v = min(v, __shfl_xor_sync(FULL, v, m));
    return v;
}

// distribution truncated at KSM = 16 (P[k] = P(k errors), k < 16; T = P(>= 16 errors)) for the small counts of the other
// alleles of a heavy column (a few sequencing errors beside the variant): one read folded in ...
constexpr int KSM = 16;
template <int KV>
__device__ __forceinline__ void small_update(double (&P)[KV], double &T, double p, double q)
{
    T = fma(P[KV - 1], p, T);
#pragma unroll
    for (int k = KV - 1; k >= 1; --k) P[k] = fma(P[k - 1], p, P[k] * q);
    P[0] = P[0] * q;
}

// ... and the G per-lane distributions of a column merged by truncated convolution (butterfly inside the group)
template <int G, int KV>
__device__ __forceinline__ void small_merge(double (&P)[KV], double &T)
{
#pragma unroll 1
    for (int m = 1; m < G; m <<= 1) {
        // c[k] = sum_i P[i] b[k-i]; the partner's cells are fetched one at a time inside the convolution
        double cc[KV];
        const double tb = __shfl_xor_sync(FULL, T, m);
        double sum_a = 0.0, sum_b = 0.0;
#pragma unroll
        for (int k = 0; k < KV; ++k) {
            cc[k] = 0.0;
            sum_a += P[k];
        }
        double t = 0.0, asuf = 0.0;
#pragma unroll
        for (int j = 0; j < KV; ++j) {
            const double bj = __shfl_xor_sync(FULL, P[j], m);
            sum_b += bj;
#pragma unroll
            for (int k = j; k < KV; ++k) cc[k] = fma(P[k - j], bj, cc[k]);
            if (j >= 1) {
                asuf += P[KV - j];
                t = fma(bj, asuf, t);
            }
        }
        t += T * (sum_b + tb) + tb * sum_a;
#pragma unroll
        for (int k = 0; k < KV; ++k) P[k] = cc[k];
        T = t;
    }
}

// ------------------------------------------------------------------------------------------------
// mbarrier + 1-D bulk TMA (cp.async.bulk), PTX ISA 8.x
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_inval(unsigned long long *bar)
{
    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_arrive(unsigned long long *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned; completion is signalled on `bar`
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gmem_src, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// layout of the dynamic shared memory
// ------------------------------------------------------------------------------------------------
constexpr int DP_WARPS = 4;
constexpr int DP_NCOLMAX = 8;               // columns per warp at G = 4
// A warp works in superblocks of 8 G reads of each of its 32 / G columns (32 reads at G = 4 .. 256 at G = 32): 256
// (column, read) pairs, 8 per lane, whatever the class.  The step parameters of a superblock are made in one go (the
// latency of a parameter's dependent chain is paid once per 8 reads of a lane), the recurrence then runs over it in blocks
// of 32 reads.  One TMA stage holds one superblock (two at G = 4).
constexpr int DP_PAR_ENTRIES = 264;         // parameter rows: 32 / G rows of 8 G + 1 entries
constexpr int DP_STAGE_ROWS = 640;          // bytes per stage and plane: 32 / G rows of (reads per stage + 16) bytes, at most 8 * 80
__host__ __device__ constexpr int dp_superblock(int G) { return 8 * G; }
__host__ __device__ constexpr int dp_stage_reads(int G) { return G == 4 ? 64 : 8 * G; }

struct DpWarpSmem {
    unsigned long long bar[2];
    union {
        double2 par[DP_PAR_ENTRIES];                                  // step parameters of the current superblock
        // before the recurrence starts the same bytes serve the set-up of the task:
        unsigned char hist_bytes[DP_NCOLMAX * 512];                   // ColHist per column of the warp (tilt)
        int median_hist[256];                                         // def_alt_bq == -1 (warp_ref_median)
    } u;
    // followed by the byte stages: [2][planes][DP_STAGE_ROWS]
};

__host__ __device__ constexpr size_t dp_warp_bytes(int planes)
{
    return ((sizeof(DpWarpSmem) + 15) & ~(size_t)15) + (size_t)2 * planes * DP_STAGE_ROWS;
}

// processing order of the job lists (longest tasks first): the unbinned lists, then bin by bin from the deepest, and
// inside a bin the classes with the most columns per warp first.  tbase[i] = first task of entry i, tbase[DP_NL] = all
__device__ __forceinline__ int dp_entry_list(int i) { return (i % DP_NCLS) * DP_NBIN1 + (DP_NBIN1 - 1 - i / DP_NCLS); }
__device__ __forceinline__ int dp_cols_per_task(int cls) { return cls == 0 ? 8 : cls == 1 ? 4 : cls == 2 ? 2 : 1; }

// whole CTA: every thread counts the tasks of one entry, thread 0 sums them up
__device__ void dp_list_bases(const Workspace &ws, unsigned *tbase, int cls_lo, int cls_hi)
{
    for (int i = threadIdx.x; i < DP_NL; i += blockDim.x) {
        const int li = dp_entry_list(i);
        const int cls = li / DP_NBIN1;
        unsigned tasks = 0;
        if (cls >= cls_lo && cls <= cls_hi) {
            const bool unbinned = (li % DP_NBIN1) == DP_NBIN;
            const unsigned nj = unbinned ? ws.counters->n_pjobs[li] : min(ws.counters->n_pjobs[li], (unsigned)ws.pcap);
            const unsigned ncol = dp_cols_per_task(cls);
            tasks = (nj + ncol - 1) / ncol;
        }
        tbase[i + 1] = tasks;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned acc = 0;
        tbase[0] = 0;
        for (int i = 1; i <= DP_NL; ++i) {
            acc += tbase[i];
            tbase[i] = acc;
        }
    }
    __syncthreads();
}

__device__ __forceinline__ const int *dp_list_ptr(const Workspace &ws, int li)
{
    const int cls = li / DP_NBIN1, bin = li % DP_NBIN1;
    return bin == DP_NBIN ? ws.ujobs + (long long)cls * ws.cap_cols : ws.pjobs + ((long long)cls * DP_NBIN + bin) * ws.pcap;
}

// ------------------------------------------------------------------------------------------------
// tilt: saddlepoint equation sum_n p_n s/(q_n + p_n s) = min(K, N - 1/2), solved on a histogram of the merged
// probabilities of the column (64 buckets, two per binade: inside a bucket p varies by at most 41 %, and the bucket
// mean stands for it — the tolerance of the tilt is coarse: an error e in ln s costs about var * e^2 / 2 nats of ~700)
// ------------------------------------------------------------------------------------------------
constexpr int DP_NB = 64;
struct ColHist {
    float sum[DP_NB];
    int cnt[DP_NB];
};

__device__ __forceinline__ int hist_bucket(double p)
{
    const int bk = (0x3ff00000 - __double2hiint(p)) >> 19;
    return min(max(bk, 0), DP_NB - 1);
}

// Newton on ln s, the G lanes of a column over its histogram; columns that need no tilt idle along (need == false)
template <int G>
__device__ __forceinline__ double group_tilt(const ColHist &h, bool need, int K, int N, double lam, double scale)
{
    const int gl = lane_id() % G;
    const double kt = fmin((double)K, (double)N - 0.5);
    const double s0 = kt * fmax((double)N - lam, 1e-300) / (fmax(lam, 1e-300) * fmax((double)N - kt, 0.5));
    double lo = 0.0, hi = 60.0;
    double ls = fmin(dlog(fmax(s0, 1.0)), hi);
    bool conv = !need;
#pragma unroll 1
    for (int it = 0; it < 40; ++it) {
        if (__all_sync(FULL, conv)) break;
        const float sf = (float)dexp(ls);
        float gs = 0.f, ds = 0.f;
        if (!conv) {
#pragma unroll 1
            for (int k = gl; k < DP_NB; k += G) {
                const int c = h.cnt[k];
                if (c == 0) continue;
                const float pb = fminf(h.sum[k] / (float)c, 1.f);
                const float ps = pb * sf;
                const float w = __fdividef(ps, fmaxf(1.f - pb, 1e-30f) + ps);        // p s / (q + p s)
                gs = fmaf((float)c, w, gs);
                ds = fmaf((float)c * w, 1.f - w, ds);                                  // derivative with respect to ln s
            }
        }
        const double gsum = group_sum<G>((double)gs) * scale - kt;       // the histogram holds a sample: scale = reads / sampled reads
        const double d = group_sum<G>((double)ds) * scale;
        if (conv) continue;
        // an error e in ln s costs about d*e^2/2 nats of head-room (of ~700): stop once that is negligible
        const double step = d > 0.0 ? gsum / d : 0.0;
        if (d > 0.0 && fabs(step) * sqrt(fmax(d, 1.0)) < 0.5) {
            ls = fmin(fmax(ls - step, 0.0), 60.0);
            conv = true;
            continue;
        }
        if (gsum > 0.0) hi = fmin(hi, ls); else lo = fmax(lo, ls);
        double nl = d > 0.0 ? ls - step : 0.5 * (lo + hi);
        if (!(nl > lo && nl < hi)) nl = 0.5 * (lo + hi);
        ls = nl;
    }
    return need ? ls : 0.0;
}

// P(X >= c) for the counts c0, c1, c2 (<= KSM) of one column, exactly: every lane of the column's group folds its 16-byte
// chunks of reads into a distribution truncated at KSM, the G distributions are merged by truncated convolution.
// Whole warp (the shuffles of the merge); lanes with want == false only take part.  Its own register allocation: the
// recurrence's registers are not live here.
template <int G, int KV>
__device__ __noinline__ void small_tails_sweep(const DevConf &cf, const DevBatch &b, const EvalMode &em, const double *s_lut, const Geom &g,
                                               int lead, bool want, int c0, int c1, int c2, double (&tails)[3])
{
    const int gl = lane_id() % G;
    double P[KV], TS = 0.0;
#pragma unroll
    for (int k = 0; k < KV; ++k) P[k] = (k == 0) ? 1.0 : 0.0;
    const long long abase = g.off & ~15ll;
    const int nch = want ? (lead + g.n + 15) >> 4 : 0;
#pragma unroll 1
    for (int i = gl; i < nch; i += G) {
        Chunk16 ch;
        load_chunk(cf, b, abase + 16ll * i, ch);
        const int pos0 = 16 * i - lead;
#pragma unroll 1
        for (int wd = 0; wd < 4; ++wd) {
            const unsigned wbq = wd == 0 ? ch.bq.x : wd == 1 ? ch.bq.y : wd == 2 ? ch.bq.z : ch.bq.w;
            const unsigned wmq = wd == 0 ? ch.mq.x : wd == 1 ? ch.mq.y : wd == 2 ? ch.mq.z : ch.mq.w;
            const unsigned wbaq = wd == 0 ? ch.baq.x : wd == 1 ? ch.baq.y : wd == 2 ? ch.baq.z : ch.baq.w;
            const unsigned wsq = wd == 0 ? ch.sq.x : wd == 1 ? ch.sq.y : wd == 2 ? ch.sq.z : ch.sq.w;
#pragma unroll 1
            for (int j = 0; j < 4; ++j) {
                const int pos = pos0 + 4 * wd + j;
                double jp = 0.0;
                if (pos < 0 || pos >= g.n ||
                    !dp_eval(cf, em, s_lut, g, pos, (wbq >> (8 * j)) & 0xff, (wmq >> (8 * j)) & 0xff, (wbaq >> (8 * j)) & 0xff, (wsq >> (8 * j)) & 0xff, jp))
                    continue;
                double p, q;
                guard_pq(jp, p, q);
                small_update<KV>(P, TS, p, q);
            }
        }
    }
    small_merge<G, KV>(P, TS);
    double t0 = TS, t1 = TS, t2 = TS;
#pragma unroll
    for (int k = KV - 1; k >= 0; --k) {      // small terms first
        if (k >= c0) t0 += P[k];
        if (k >= c1) t1 += P[k];
        if (k >= c2) t2 += P[k];
    }
    tails[0] = t0;
    tails[1] = t1;
    tails[2] = t2;
}

// lower bound of ln(x) for a positive normal double from its bits: x = m 2^e, ln m >= (m - 1) ln 2 on [1, 2)
__device__ __forceinline__ double ln_lower(double x)
{
    const int hi = __double2hiint(x);
    const int e = (hi >> 20) - 1023;
    const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(x));
    return ((double)e + (m - 1.0)) * LN2;
}

// ------------------------------------------------------------------------------------------------
// one warp task: 32/G columns in lock step
// ------------------------------------------------------------------------------------------------
template <int G, int R>
__device__ void dp_task(const DevConf &cf, const DevBatch &b, const Workspace &ws, const double *s_lut, const Lut *lut, DpWarpSmem &sm,
                        unsigned char *stage_bytes, int planes, const int *list, unsigned j0, unsigned nj)
{
    constexpr int NCOL = 32 / G;
    const int lane = lane_id(), grp = lane / G, gl = lane % G;
#ifdef LFB_DP_PROF
    long long pf_t0 = clock64(), pf_par = 0, pf_loop = 0, pf_chk = 0, pf_a, pf_setup, pf_pre, pf_tilt, pf_main, pf_w = 0, pf_e = 0, pf_b;
#endif
    const int last = grp * G + G - 1;                   // the lane that owns the top cells and the absorbing state
    bool have = j0 + grp < nj;
    const long long c = have ? list[j0 + grp] : -1;
    Geom g;
    g.off = 0; g.n = 0; g.b1 = g.b2 = g.b3 = 0; g.ref_idx = -1; g.alt_bp = 0.0;
    int cnt[3] = {0, 0, 0};
    long long bonf = 1;
    if (have) {
        int cov;
        load_geom(b, c, g, cov);
#pragma unroll
        for (int i = 0; i < 3; ++i) cnt[i] = ws.cnt6[6 * c + i];
        bonf = bonf_of(cf, ws.counters->bonf_start_used, col_rank(ws, c));
    }
    const int K = max(cnt[0], max(cnt[1], cnt[2]));
    // alt-base quality override (snpcaller.c:431-441)
    if (cf.alt_bq_mode == 1) {
        g.alt_bp = cf.alt_bq_prob;
    } else if (cf.alt_bq_mode == 2) {
        // median of the reference-base qualities: whole warp, one column at a time
#pragma unroll 1
        for (int ci = 0; ci < NCOL; ++ci) {
            Geom gg;
            gg.off = __shfl_sync(FULL, g.off, ci * G);
            gg.n = __shfl_sync(FULL, g.n, ci * G);
            gg.b1 = __shfl_sync(FULL, g.b1, ci * G);
            gg.b2 = __shfl_sync(FULL, g.b2, ci * G);
            gg.b3 = __shfl_sync(FULL, g.b3, ci * G);
            gg.ref_idx = __shfl_sync(FULL, g.ref_idx, ci * G);
            gg.alt_bp = 0.0;
            if (!__shfl_sync(FULL, (int)have, ci * G)) continue;
            setup_alt_bq(cf, b, s_lut, gg, sm.u.median_hist);
            if (grp == ci) g.alt_bp = gg.alt_bp;
        }
    }

    // the bytes of reads [ks*ST, (ks+1)*ST) of this group's column -> stage buffer ks & 1
    constexpr int SB = dp_superblock(G), ST = dp_stage_reads(G), ROW = ST + 16;     // ROW: + the lead of an unaligned column
    static_assert(NCOL * ROW <= DP_STAGE_ROWS && NCOL * (SB + 1) <= DP_PAR_ENTRIES, "shared memory layout");
    const int n_mine = have ? g.n : 0;
    const int nmax = __reduce_max_sync(FULL, n_mine);
    const int lead = (int)(g.off & 15ll);
    auto issue = [&](int ks) {
        if (gl == 0) {
            unsigned long long *bar = &sm.bar[ks & 1];
            const int r0 = ks * ST;
            if (r0 < n_mine) {
                const long long a0 = (g.off + r0) & ~15ll;
                const unsigned nbytes = (unsigned)((lead + min(ST, n_mine - r0) + 15) & ~15);
                mbar_arrive_expect_tx(bar, nbytes * (unsigned)planes);
                unsigned char *dst = stage_bytes + (size_t)(ks & 1) * planes * DP_STAGE_ROWS + grp * ROW;
                int pl = 0;
                tma_load_1d(dst + (size_t)(pl++) * DP_STAGE_ROWS, b.bq + a0, nbytes, bar);
                if (cf.use_mq) tma_load_1d(dst + (size_t)(pl++) * DP_STAGE_ROWS, b.mq + a0, nbytes, bar);
                if (cf.use_baq) tma_load_1d(dst + (size_t)(pl++) * DP_STAGE_ROWS, b.baq + a0, nbytes, bar);
                if (cf.use_sq) tma_load_1d(dst + (size_t)(pl++) * DP_STAGE_ROWS, b.sq + a0, nbytes, bar);
            } else {
                mbar_arrive(bar);
            }
        }
    };
    if (lane == 0) {
        mbar_init(&sm.bar[0], NCOL);
        mbar_init(&sm.bar[1], NCOL);
        fence_barrier_init();
    }
    __syncwarp();
    // the first stage travels while the pre-pass runs
    int issued = 0, waited = 0;
    if (nmax > 0) { issue(0); issued = 1; }

    const EvalMode em = eval_mode(cf);
#ifdef LFB_DP_PROF
    pf_setup = clock64();
#endif

    // ---- 1. pre-pass, G lanes per column, 16-byte loads: what the tilt needs — the number of reads kept, lambda and the
    // histogram of the merged probabilities — estimated from a sample of 8 .. 32 chunks of 16 reads spread evenly over the
    // column (the reads are grouped by base, and error reads are of lower quality than the rest: a prefix would not do).
    // The tilt is a free parameter of the recurrence — the results do not depend on it, only the numeric range the cells
    // stay in — so an estimate is enough: a relative error e of lambda costs about K e^2 / 2 nats of ~700, and a row that
    // leaves the range all the same is detected (gap tests below) and handed to the per-column kernel.
    int N = 0;
    double lam = 0.0, scale = 1.0;
    ColHist &hist = *reinterpret_cast<ColHist *>(sm.u.hist_bytes + grp * sizeof(ColHist));
    for (int i = gl; i < DP_NB; i += G) {
        hist.sum[i] = 0.f;
        hist.cnt[i] = 0;
    }
    __syncwarp();
    {
        // runs of reads that fall into the same bucket (uniform qualities: all of them) are summed in registers first
        int run_b = -1, run_c = 0;
        float run_s = 0.f;
        auto flush = [&]() {
            if (run_b >= 0) {
                atomicAdd(&hist.sum[run_b], run_s);
                atomicAdd(&hist.cnt[run_b], run_c);
            }
            run_b = -1;
        };
        constexpr int SAMPLE = G >= 8 ? G : 8;           // chunks per column
        const long long abase = g.off & ~15ll;
        const int nch = have ? (lead + g.n + 15) >> 4 : 0;
        const int m = min(nch, SAMPLE);
        int seen = 0;                                    // positions of the column the sample has looked at
#pragma unroll 1
        for (int k = gl; k < m; k += G) {
            const int i = nch <= SAMPLE ? k : (int)(((long long)k * nch) / SAMPLE);
            Chunk16 ch;
            load_chunk(cf, b, abase + 16ll * i, ch);
            const int pos0 = 16 * i - lead;
            const bool inside = pos0 >= 0 && pos0 + 16 <= g.n;
            seen += min(pos0 + 16, g.n) - max(pos0, 0);
#pragma unroll 1
            for (int wd = 0; wd < 4; ++wd) {
                const unsigned wbq = wd == 0 ? ch.bq.x : wd == 1 ? ch.bq.y : wd == 2 ? ch.bq.z : ch.bq.w;
                const unsigned wmq = wd == 0 ? ch.mq.x : wd == 1 ? ch.mq.y : wd == 2 ? ch.mq.z : ch.mq.w;
                const unsigned wbaq = wd == 0 ? ch.baq.x : wd == 1 ? ch.baq.y : wd == 2 ? ch.baq.z : ch.baq.w;
                const unsigned wsq = wd == 0 ? ch.sq.x : wd == 1 ? ch.sq.y : wd == 2 ? ch.sq.z : ch.sq.w;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int pos = pos0 + 4 * wd + j;
                    double jp = 0.0;
                    const bool ok = (inside || (pos >= 0 && pos < g.n)) &&
                                    dp_eval(cf, em, s_lut, g, pos, (wbq >> (8 * j)) & 0xff, (wmq >> (8 * j)) & 0xff, (wbaq >> (8 * j)) & 0xff,
                                            (wsq >> (8 * j)) & 0xff, jp);
                    if (!ok) continue;
                    const double p = jp < DEPS ? DEPS : jp;
                    lam += p;
                    ++N;
                    const int bk = hist_bucket(p);
                    if (bk != run_b) {
                        flush();
                        run_b = bk;
                        run_s = 0.f;
                        run_c = 0;
                    }
                    run_s += (float)p;
                    ++run_c;
                }
            }
        }
        flush();
        N = group_sum_i<G>(N);
        lam = group_sum<G>(lam);
        seen = group_sum_i<G>(seen);
        if (have && seen > 0 && seen < g.n) {
            scale = (double)g.n / (double)seen;
            lam *= scale;
            N = (int)((double)N * scale);
        }
    }
    __syncwarp();
#ifdef LFB_DP_PROF
    pf_pre = clock64();
#endif
    // Chernoff exponent of the tail: beyond ~300 nats the untilted cells of interest drift out of fp64 range
    const double cher = (have && (double)K > lam) ? ((double)K * dlog((double)K / fmax(lam, 1e-300)) - (double)K + lam) : 0.0;
    const double ln_s = group_tilt<G>(hist, have && cher > 300.0, K, N, lam, scale);
    __syncwarp();                                       // the histograms give way to the step parameters
    const double s = (ln_s == 0.0) ? 1.0 : dexp(ln_s);

#ifdef LFB_DP_PROF
    pf_tilt = clock64();
#endif
    // ---- 2./3. the recurrence, all columns of the warp in lock step
    const int k0 = K - G * R + gl * R;                  // cell of register 0 (k < 0: padding, stays 0)
    double E[R], T = 0.0;
#pragma unroll
    for (int r = 0; r < R; ++r) E[r] = (have && k0 + r == 0) ? 1.0 : 0.0;
    int e2 = 0;
    const double thr_ln = dlog(cf.sig * (1.0 + 1e-9) / (double)bonf) + (double)K * ln_s;
    bool dead = !have, fb = false;
    double lq_acc = 0.0, qprod = 1.0;                   // ln of the product of q over this lane's reads = lq_acc + ln(qprod)
    double lq_sb = 0.0;                                 // lower bound of the column's sum of ln q up to the end of the current superblock
    double2 *const prow = sm.u.par + grp * (SB + 1);
    // step parameters of superblock sb -> sm.par (single buffer: written between two superblocks)
    auto make_params = [&](int sb) {
        const int rbase = sb * SB;
#ifdef LFB_DP_PROF
        pf_b = clock64();
#endif
        if ((rbase % ST) == 0) {
            const int ks = rbase / ST;
            // (the buffer stage ks+1 overwrites held stage ks-1: every lane left it before the last __syncwarp)
            if ((ks + 1) * ST < nmax) {
                issue(ks + 1);
                issued = ks + 2;
            }
            mbar_wait(&sm.bar[ks & 1], (unsigned)(ks >> 1) & 1u);
            waited = ks + 1;
        }
#ifdef LFB_DP_PROF
        pf_w += clock64() - pf_b; pf_b = clock64();
#endif
        const int ks = rbase / ST;
        const unsigned char *src = stage_bytes + (size_t)(ks & 1) * planes * DP_STAGE_ROWS + grp * ROW + lead + (rbase - ks * ST);
        bool bad = false;
        if (em.uniform && em.plain_merge) {
            // The default configuration (bq and mq merged, reference and alt reads alike), straight-line: 1 / q of a (bq, mq)
            // pair is the product of two table entries, so nothing here waits for a division, and the reads of a lane are
            // independent chains the scheduler interleaves.  (The relative error of that product, ~4e-16, enters T once per
            // read; a q of zero — probability 1 — gives inf and sends the column to the per-column kernel like any step too
            // large to hold between two rescalings.)
#pragma unroll
            for (int i0 = 0; i0 < 8; i0 += 4) {
                int bqv[4], mqv[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int t = gl + G * (i0 + u);
                    bqv[u] = src[t];
                    mqv[u] = cf.use_mq ? src[DP_STAGE_ROWS + t] : 255;
                }
                double bpv[4], mpv[4], rqv[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    bpv[u] = s_lut[bqv[u]];
                    mpv[u] = s_lut[256 + mqv[u]];
                    rqv[u] = __ldg(&lut->rbq[bqv[u]]) * __ldg(&lut->rmq[mqv[u]]);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int t = gl + G * (i0 + u);
                    const bool ok = !dead && rbase + t < n_mine && bqv[u] >= cf.min_bq;
                    const double jp = __dadd_rn(mpv[u], __dmul_rn(__dsub_rn(1.0, mpv[u]), bpv[u]));
                    double p, q;
                    guard_pq(jp, p, q);
                    const double rq = rqv[u];
                    const double o = p * s * rq;
                    // between two rescalings (32 reads) a cell may grow by (1 + o)^32 and the absorbing state by (1/q)^32
                    bad |= ok && !(o <= 1048576.0 && rq <= 1048576.0);
                    qprod *= ok ? q : 1.0;
                    prow[t] = ok ? make_double2(o, rq) : make_double2(0.0, 1.0);     // neutral step: the row stays as it is
                }
            }
        } else if (em.uniform) {
            // the same with alignment and / or source qualities merged in (the reference's default carries BAQ): 1 / q is the
            // product of up to four table entries, the merged probability keeps the reference's association order
            const int o_baq = (cf.use_mq ? 2 : 1) * DP_STAGE_ROWS, o_sq = o_baq + (cf.use_baq ? DP_STAGE_ROWS : 0);
#pragma unroll
            for (int i0 = 0; i0 < 8; i0 += 2) {
                int bqv[2], mqv[2], baqv[2], sqv[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int t = gl + G * (i0 + u);
                    bqv[u] = src[t];
                    mqv[u] = cf.use_mq ? src[DP_STAGE_ROWS + t] : 255;
                    baqv[u] = cf.use_baq ? src[o_baq + t] : 255;
                    sqv[u] = cf.use_sq ? src[o_sq + t] : 255;
                }
                double jpv[2], rqv[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    jpv[u] = merge4(s_lut[512 + sqv[u]], s_lut[256 + mqv[u]], s_lut[512 + baqv[u]], s_lut[bqv[u]]);
                    rqv[u] = (__ldg(&lut->rbq[bqv[u]]) * __ldg(&lut->rmq[mqv[u]])) * (__ldg(&lut->raq[baqv[u]]) * __ldg(&lut->raq[sqv[u]]));
                }
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int t = gl + G * (i0 + u);
                    const bool ok = !dead && rbase + t < n_mine && bqv[u] >= cf.min_bq;
                    double p, q;
                    guard_pq(jpv[u], p, q);
                    const double rq = rqv[u];
                    const double o = p * s * rq;
                    bad |= ok && !(o <= 1048576.0 && rq <= 1048576.0);
                    qprod *= ok ? q : 1.0;
                    prow[t] = ok ? make_double2(o, rq) : make_double2(0.0, 1.0);
                }
            }
        } else {
#pragma unroll 1
            for (int i = 0; i < 8; ++i) {
                const int t = gl + G * i;                   // read of this superblock
                const int pos = rbase + t;
                double2 e = make_double2(0.0, 1.0);
                if (!dead && pos < n_mine) {
                    int pl = 0;
                    const int bq = src[(size_t)(pl++) * DP_STAGE_ROWS + t];
                    const int mq = cf.use_mq ? src[(size_t)(pl++) * DP_STAGE_ROWS + t] : 0;
                    const int baq = cf.use_baq ? src[(size_t)(pl++) * DP_STAGE_ROWS + t] : 0;
                    const int sq = cf.use_sq ? src[(size_t)(pl++) * DP_STAGE_ROWS + t] : 0;
                    double jp;
                    if (dp_eval(cf, em, s_lut, g, pos, bq, mq, baq, sq, jp)) {
                        double p, q;
                        guard_pq(jp, p, q);
                        const double rq = 1.0 / q;
                        const double o = p * s * rq;
                        e = make_double2(o, rq);
                        qprod *= q;
                        bad |= (o > 1048576.0 || rq > 1048576.0);
                    }
                }
                prow[t] = e;
            }
        }
#ifdef LFB_DP_PROF
        pf_e += clock64() - pf_b;
#endif
        if (qprod < 1e-200) {
            lq_acc += dlog(qprod);
            qprod = 1.0;
        }
        if (__any_sync(FULL, bad)) {
            // this column cannot be held between two rescalings: the per-column kernel rescales after every read
            const unsigned bm = __ballot_sync(FULL, bad);
            const unsigned gm = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (grp * G));
            if (bm & gm) { fb = true; dead = true; }
            __syncwarp();
            if (fb) {
#pragma unroll
                for (int i = 0; i < 8; ++i) prow[gl + G * i] = make_double2(0.0, 1.0);
            }
        }
        // The early exit below wants a lower bound of the sum of ln q over the reads seen; the sum up to the end of this
        // superblock is one (every term is negative), and costs one reduction per superblock instead of one per block.
        lq_sb = group_sum<G>(lq_acc + ln_lower(qprod));
        __syncwarp();
    };
#ifdef LFB_DP_PROF
    pf_a = clock64();
#endif
    if (nmax > 0) make_params(0);
#ifdef LFB_DP_PROF
    pf_par += clock64() - pf_a;
#endif
    for (int n0 = 0; n0 < nmax; n0 += 32) {
#ifdef LFB_DP_PROF
        pf_a = clock64();
#endif
        const double2 *pp = prow + (n0 % SB);
        // software-pipelined: the parameters of read j+1 and the boundary cell for read j+1 (the top cell right
        // after its own update) are requested before the remaining R-1 cells of read j are updated
        double2 c_next = pp[0];
        // the boundary cell travels as two 32-bit halves; the first lane of a column takes zeros instead (selected on the
        // halves, so that the pair the DFMA reads is formed by the selects themselves)
        int in_hi = __shfl_up_sync(FULL, __double2hiint(E[R - 1]), 1), in_lo = __shfl_up_sync(FULL, __double2loint(E[R - 1]), 1);
#pragma unroll 4
        for (int j = 0; j < 32; ++j) {
            const double2 cc = c_next;
            const double in = __hiloint2double(gl == 0 ? 0 : in_hi, gl == 0 ? 0 : in_lo);
            c_next = pp[j + 1];                          // (entry SB of the row is padding)
            const double top = E[R - 1];
            T = fma(top, cc.x, T * cc.y);
            E[R - 1] = fma(E[R - 2], cc.x, top);
            in_hi = __shfl_up_sync(FULL, __double2hiint(E[R - 1]), 1);
            in_lo = __shfl_up_sync(FULL, __double2loint(E[R - 1]), 1);
#pragma unroll
            for (int r = R - 2; r >= 1; --r) E[r] = fma(E[r - 1], cc.x, E[r]);
            E[0] = fma(in, cc.x, E[0]);
        }
#ifdef LFB_DP_PROF
        pf_loop += clock64() - pf_a; pf_a = clock64();
#endif
        // exact power-of-two rescaling, per column
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
        // Early exit (the reference's, snpcaller.c:916-958): the tail over the reads seen so far can only grow, so once a
        // lower bound of ln P(X >= K among them) = ln T + e2 ln 2 + sum ln q - K ln s passes ln(sig / bonf) the column is
        // insignificant whatever follows.  Lower bounds of the logarithms from the bits of T and of the q products.
        const bool over = lane == last && T > 1e-300 && ln_lower(T) + (double)e2 * LN2 + lq_sb > thr_ln;
        const bool over_g = __shfl_sync(FULL, (int)over, last) != 0;     // every lane takes part, dead or not
        dead = dead || over_g;
#ifdef LFB_DP_PROF
        pf_chk += clock64() - pf_a; pf_a = clock64();
#endif
        if (__all_sync(FULL, dead || n0 + 32 >= n_mine)) break;      // nothing left to decide in this warp
        if (((n0 + 32) % SB) == 0) {
            __syncwarp();                              // sm.par is rewritten now
            make_params((n0 + 32) / SB);
        }
#ifdef LFB_DP_PROF
        pf_par += clock64() - pf_a;
#endif
    }
#ifdef LFB_DP_PROF
    pf_main = clock64();
#endif
    // a prefetched stage may still be in flight: it must land before the buffers and barriers are reused
    if (issued > waited) mbar_wait(&sm.bar[(issued - 1) & 1], (unsigned)((issued - 1) >> 1) & 1u);
    __syncwarp();
    if (lane == 0) {
        mbar_inval(&sm.bar[0]);
        mbar_inval(&sm.bar[1]);
    }
    const double sum_lq = group_sum<G>(lq_acc + dlog(qprod));

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
    const bool ruled_out = dead && !fb;                 // early exit fired: insignificant for good
    if (!ruled_out && ((gap > 580 && ln_s == 0.0) || gap > 900)) fb = true;     // needs the tilt after all / out of range
    const double base = (double)e2 * LN2 + sum_lq;
    const double lnT = dlog(Tl) + base - (double)K * ln_s;
    const double lnKm1 = dlog(topl) + base - (double)(K - 1) * ln_s;
    bool site = have && !fb && !dead;
    if (site && lnT > -700.0 && dexp(lnT) * (double)bonf > cf.sig * (1.0 + 1e-9)) site = false;   // snpcaller.c:1155
    double lnp0 = 0.0, lnp1 = 0.0, lnp2 = 0.0;
    const double invs = (ln_s == 0.0) ? 1.0 : dexp(-ln_s);
    bool sweep_for[3] = {false, false, false};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int ci = cnt[i];
        double &out = i == 0 ? lnp0 : i == 1 ? lnp1 : lnp2;
        if (ci == K) out = lnT;
        const bool mine = site && ci > 0 && ci < K;
        if (!__any_sync(FULL, mine)) continue;
        // P(X >= ci) = sum_{k >= ci} E[k] s^-(k-ci) + T s^-(K-ci), times s^-ci and the common scale
        double acc = 0.0;
        int hc = 0x7fffffff;
        if (mine) {
            const int kk = max(k0, ci);
            double f = (ln_s == 0.0) ? 1.0 : dexp(-(double)(kk - ci) * ln_s);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (k0 + r >= ci) {
                    acc = fma(E[r], f, acc);
                    f *= invs;
                }
                if (k0 + r == ci) hc = __double2hiint(E[r]);
            }
            if (lane == last) acc = fma(T, (ln_s == 0.0) ? 1.0 : dexp(-(double)(K - ci) * ln_s), acc);
        }
        acc = group_sum<G>(acc);
        hc = group_min_i<G>(hc);
        if (mine) {
            // the leading cell must have stayed a normal number all along (rescaling keeps the peak within 2^+-200
            // at the block boundaries and below 2^840 inside a block); a small count whose cells were lost on a strongly
            // tilted row is computed exactly below, a large one sends the column to the per-column fallback
            if ((hc >> 20) < 64 || peak - (hc >> 20) > 850) {
                if (ci <= KSM) sweep_for[i] = true; else fb = true;
            }
            out = dlog(acc) + base - (double)ci * ln_s;
        }
    }
    // Alleles with a count of at most KSM (a few sequencing errors beside the variant) that could not be read off the
    // row: their tail exactly, from the distribution truncated at KSM.  After the row has been read (its registers are
    // free now); rare on shallow data, the rule on deep noisy columns.
    const bool want_small = !fb && (sweep_for[0] || sweep_for[1] || sweep_for[2]);
    if (__any_sync(FULL, want_small)) {
        double t3[3];
        int maxc = 0;
#pragma unroll
        for (int i = 0; i < 3; ++i) if (sweep_for[i]) maxc = max(maxc, cnt[i]);
        if (__reduce_max_sync(FULL, want_small ? maxc : 0) <= 8) small_tails_sweep<G, 8>(cf, b, em, s_lut, g, lead, want_small, cnt[0], cnt[1], cnt[2], t3);
        else small_tails_sweep<G, KSM>(cf, b, em, s_lut, g, lead, want_small, cnt[0], cnt[1], cnt[2], t3);
        if (want_small && sweep_for[0]) lnp0 = dlog(t3[0]);
        if (want_small && sweep_for[1]) lnp1 = dlog(t3[1]);
        if (want_small && sweep_for[2]) lnp2 = dlog(t3[2]);
    }
    if (ruled_out) fb = false;
    if (fb) site = false;
    if (have && gl == 0) {
        if (fb) {
            const int cls = max(3, class_of(K));          // k_heavy_all: R = 8 .. 64 cells per lane
            const unsigned slot = atomicAdd(&ws.counters->n_jobs[cls], 1u);
            ws.jobs[(long long)cls * ws.cap_cols + slot] = (int)c;
        } else if (site) {
            Cand cd;
            cd.col = c;
            cd.bonf = bonf;
            cd.lnp[0] = cnt[0] > 0 ? lnp0 : 0.0;
            cd.lnp[1] = cnt[1] > 0 ? lnp1 : 0.0;
            cd.lnp[2] = cnt[2] > 0 ? lnp2 : 0.0;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
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
#ifdef LFB_DP_PROF
    if (lane == 0 && (blockIdx.x % 5) == 0)
        printf("dp_task G=%d R=%d blk %d nmax %d: setup %lld pre %lld tilt %lld main %lld (par %lld loop %lld chk %lld) tails %lld total %lld start %lld parwait %lld pareval %lld\n", G, R, blockIdx.x, nmax,
               pf_setup - pf_t0, pf_pre - pf_setup, pf_tilt - pf_pre, pf_main - pf_tilt, pf_par, pf_loop, pf_chk, (long long)clock64() - pf_main, (long long)clock64() - pf_t0, pf_t0, pf_w, pf_e);
#endif
}

// One kernel per register budget: RC = 0: R = 8 cells per lane (classes 0..3, K <= 256, the bulk of the columns; 6 CTAs
// per SM), RC = 1: R = 16 and 32 (classes 4, 5, K <= 1024), RC = 2: R = 64 (class 6, K <= 2048).
template <int RC>
__global__ void __launch_bounds__(32 * DP_WARPS, RC == 0 ? 6 : RC == 1 ? 4 : 1)
k_dp(const __grid_constant__ DevConf cf, const __grid_constant__ DevBatch b, const Lut *lut, const Workspace ws, int planes)
{
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    __shared__ double s_lut[768];
    __shared__ unsigned s_tbase[DP_NL + 1];
    constexpr int CLS_LO = RC == 0 ? 0 : RC == 1 ? 4 : 6, CLS_HI = RC == 0 ? 3 : RC == 1 ? 5 : 6;
    dp_list_bases(ws, s_tbase, CLS_LO, CLS_HI);
    const unsigned total = s_tbase[DP_NL];
    if (total == 0) return;
    load_lut(s_lut, lut);
    const int lane = lane_id(), wib = threadIdx.x >> 5;
    unsigned char *wbase = dyn_smem + (size_t)wib * dp_warp_bytes(planes);
    DpWarpSmem &sm = *reinterpret_cast<DpWarpSmem *>(wbase);
    unsigned char *stage_bytes = wbase + ((sizeof(DpWarpSmem) + 15) & ~(size_t)15);
    // the first task of every warp is dealt out statically — with about as many tasks as resident warps, a race for
    // them leaves some SMs with six tasks per sub-partition and others with three — the rest dynamically
    const unsigned nwarps = gridDim.x * DP_WARPS;
    unsigned t = blockIdx.x * DP_WARPS + wib;
    unsigned *next = &ws.counters->next_ptask[RC];
    for (;;) {
        if (t >= total) break;
        // entry of the processing order that holds task t
        int lo = 0, hi = DP_NL - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (s_tbase[mid] <= t) lo = mid; else hi = mid - 1;
        }
        const int li = dp_entry_list(lo);
        const int cls = li / DP_NBIN1;
        const bool unbinned = (li % DP_NBIN1) == DP_NBIN;
        const unsigned nj = unbinned ? ws.counters->n_pjobs[li] : min(ws.counters->n_pjobs[li], (unsigned)ws.pcap);
        const unsigned j0 = (t - s_tbase[lo]) * (unsigned)dp_cols_per_task(cls);
        const int *list = dp_list_ptr(ws, li);
        if (RC == 2) {
            dp_task<32, 64>(cf, b, ws, s_lut, lut, sm, stage_bytes, planes, list, j0, nj);
        } else if (RC == 1) {
            if (cls == 4) dp_task<32, 16>(cf, b, ws, s_lut, lut, sm, stage_bytes, planes, list, j0, nj);
            else dp_task<32, 32>(cf, b, ws, s_lut, lut, sm, stage_bytes, planes, list, j0, nj);
        } else {
            switch (cls) {
                case 0: dp_task<4, 8>(cf, b, ws, s_lut, lut, sm, stage_bytes, planes, list, j0, nj); break;
                case 1: dp_task<8, 8>(cf, b, ws, s_lut, lut, sm, stage_bytes, planes, list, j0, nj); break;
                case 2: dp_task<16, 8>(cf, b, ws, s_lut, lut, sm, stage_bytes, planes, list, j0, nj); break;
                default: dp_task<32, 8>(cf, b, ws, s_lut, lut, sm, stage_bytes, planes, list, j0, nj); break;
            }
        }
        if (lane == 0) t = nwarps + atomicAdd(next, 1u);
        t = __shfl_sync(FULL, t, 0);
    }
}

static size_t dp_smem_bytes(int planes) { return (size_t)DP_WARPS * dp_warp_bytes(planes); }

int dp_smem_optin()
{
    // the largest configuration: all four quality planes
    const int bytes = (int)dp_smem_bytes(4);
    if (cudaFuncSetAttribute(k_dp<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess) return 1;
    if (cudaFuncSetAttribute(k_dp<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess) return 1;
    if (cudaFuncSetAttribute(k_dp<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess) return 1;
    return 0;
}

// k_dp<0> on `st`; k_dp<1> (256 < K <= 1024) and k_dp<2> (K > 1024) beside it
void launch_dp(const LaunchState &ls, const DevConf &cf, const DevBatch &b, const Lut *lut, const Workspace &ws, cudaStream_t st,
               cudaStream_t st1, cudaStream_t st2)
{
    if (b.n_cols <= 0) return;
    const int planes = 1 + (cf.use_mq ? 1 : 0) + (cf.use_baq ? 1 : 0) + (cf.use_sq ? 1 : 0);
    const size_t smem = dp_smem_bytes(planes);
    static const int dp0_ctas = getenv("LFB200_DP0_CTAS") ? atoi(getenv("LFB200_DP0_CTAS")) : 6;
    k_dp<0><<<ls.sms * dp0_ctas, 32 * DP_WARPS, smem, st>>>(cf, b, lut, ws, planes);
    k_dp<1><<<ls.sms * 4, 32 * DP_WARPS, smem, st1>>>(cf, b, lut, ws, planes);
    k_dp<2><<<ls.sms, 32 * DP_WARPS, smem, st2>>>(cf, b, lut, ws, planes);
}

}  // namespace lfb
