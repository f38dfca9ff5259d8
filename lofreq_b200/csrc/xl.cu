// k_xl: the O(depth*K) recurrence for columns with K > 2048 (deep amplicons: depth >= ~4100 at AF 50 %), one CTA per
// column, the row of K cells spread over up to 8 warps that run as a WAVEFRONT over the reads.
//
// The recurrence is serial in the reads, but cell k at read n needs only cell k-1 at read n-1, so warp w (cells
// [k0 + w*32R, k0 + (w+1)*32R)) can work on reads [32b, 32b+32) as soon as warp w-1 has finished that block: the only
// thing that crosses a warp boundary is the top cell of warp w-1 before each of its 32 updates, left in shared memory
// together with the power-of-two exponent the producing warp used for the block.  No CTA-wide barrier per read (the
// earlier kernel had one), no barrier at all in steady state: producer and consumer meet through a block counter in
// shared memory, two blocks of slack.  Every warp keeps its own exact power-of-two scale, so the row may span far more
// than the fp64 range across warps.
//
// R = 16, 32 or 64 cells per lane is chosen so that the K cells occupy as many of the 8 warps as possible (all four
// sub-partitions of the SM busy): K <= 4096 -> R = 16, <= 8192 -> 32, <= 16384 -> 64.
// Per column: pre-pass (reads kept, lambda, largest step parameters, exact small tails for alleles with count <= 8),
// tilt from a histogram (as in k_dp), then one wavefront run for the largest count and one more per further allele
// with a count above 8.  Columns with a step parameter above 2^20 go to k_heavy_xl (rescaling after every read);
// alt counts above 16384 are reported as unsupported (never silently skipped).
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "internal.h"
#include "dev_common.cuh"
#include "screen_common.cuh"

namespace lfb {

constexpr int XLW = 8;                     // warps per CTA
constexpr int XLT = 32 * XLW;
constexpr int XL_NHIST = 160;

struct XlEdge {
    double v[32];       // top cell of the producing warp before each of the 32 updates of the block
    int e2;             // its power-of-two exponent during the block
    int pad;
};

constexpr int XL_SORT_MAX = 512;

struct XlSm {
    double lut[768];
    double2 par[XLW][32];
    XlEdge edge[XLW][2];
    int done[XLW];      // blocks finished by warp w (of the current run)
    int stop;           // early exit: the tail has passed sig / bonf
    double red[XLW];
    int redi[XLW];
    double small_P[XLW][KS];
    double small_T[XLW];
    float hsum[XL_NHIST];
    int hcnt[XL_NHIST];
    int med_hist[256];
    double bcast[4];
    unsigned job;
    // longest-first order of the job list (a few hundred columns of very different cost on 148 CTAs: taken in arrival order
    // the last CTA to start a deep column decides the kernel's duration)
    float cost[XL_SORT_MAX];
    unsigned short order[XL_SORT_MAX];
};

__device__ __forceinline__ double xl_block_sum(double v, XlSm &sh)
{
    v = warp_sum(v);
    __syncthreads();
    if (lane_id() == 0) sh.red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < XLW; ++w) t += sh.red[w];
    return t;
}

__device__ __forceinline__ int xl_block_max(int v, XlSm &sh)
{
    v = __reduce_max_sync(FULL, v);
    __syncthreads();
    if (lane_id() == 0) sh.redi[threadIdx.x >> 5] = v;
    __syncthreads();
    int t = sh.redi[0];
#pragma unroll
    for (int w = 1; w < XLW; ++w) t = max(t, sh.redi[w]);
    return t;
}

__device__ __forceinline__ int ld_volatile(const int *p) { return *reinterpret_cast<const volatile int *>(p); }
__device__ __forceinline__ void st_volatile(int *p, int v) { *reinterpret_cast<volatile int *>(p) = v; }

__device__ __forceinline__ double xl_ln_lower(double x)
{
    const int hi = __double2hiint(x);
    const int e = (hi >> 20) - 1023;
    const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(x));
    return ((double)e + (m - 1.0)) * LN2;
}

struct XlOut {
    double lnT, lnKm1;
    bool ruled_out;      // the early exit fired: insignificant whatever follows
    int gap;             // exponent distance between the largest cell and the absorbing state
};

// One wavefront run: ln P(X >= K) and ln P(X = K-1) with tilt exp(ln_s).  thr_ln: early-exit threshold on
// ln T + e2 ln2 + sum ln q (already including K ln s); +inf switches the early exit off.
template <int R>
__device__ XlOut xl_run(const DevConf &cf, const DevBatch &b, const Geom &g, int K, double ln_s, double thr_ln, XlSm &sh)
{
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int W = (K + 32 * R - 1) / (32 * R);         // active warps
    const bool active = w < W;
    const bool is_last = w == W - 1;
    const double s = (ln_s == 0.0) ? 1.0 : exp(ln_s);
    const int k0 = K - W * 32 * R + (w * 32 + lane) * R;   // cell of register 0 (k < 0: padding, stays 0)
    double E[R], T = 0.0;
#pragma unroll
    for (int r = 0; r < R; ++r) E[r] = (active && k0 + r == 0) ? 1.0 : 0.0;
    int e2 = 0;
    double lq_acc = 0.0, qprod = 1.0;
    __syncthreads();
    if (tid < XLW) sh.done[tid] = 0;
    if (tid == 0) sh.stop = 0;
    __syncthreads();
    const int n = g.n;
    const int nblk = (n + 31) >> 5;
    bool ruled_out = false;
    if (active) {
        // bytes of the read this lane turns into step parameters, one block ahead
        int nbq = 0, nmq = 0, nbaq = 0, nsq = 0;
        auto fetch = [&](int blk) {
            const int pos = blk * 32 + lane;
            if (pos < n) {
                const long long a = g.off + pos;
                nbq = b.bq[a];
                if (cf.use_mq) nmq = b.mq[a];
                if (cf.use_baq) nbaq = b.baq[a];
                if (cf.use_sq) nsq = b.sq[a];
            }
        };
        fetch(0);
        for (int blk = 0; blk < nblk; ++blk) {
            // ---- wait: input from the warp below, room in the warp above
            // (every lane polls the same word: the warp stays converged, which the shuffles of the recurrence rely on —
            // a one-lane spin loop leaves the warp split and every later shuffle takes the slow divergent path)
            if (w > 0)
                while (ld_volatile(&sh.done[w - 1]) <= blk && !ld_volatile(&sh.stop)) {}
            if (!is_last)
                while (ld_volatile(&sh.done[w + 1]) < blk - 1 && !ld_volatile(&sh.stop)) {}
            __syncwarp();
            if (ld_volatile(&sh.stop)) break;
            __threadfence_block();
            // ---- step parameters of this block (every warp its own copy: 32 reads against 32 x 32R cells)
            {
                const int pos = blk * 32 + lane;
                double2 e = make_double2(0.0, 1.0);
                if (pos < n) {
                    bool is_alt;
                    int slot;
                    double jp;
                    if (eval_read<true>(cf, sh.lut, g, pos, nbq, nmq, nbaq, nsq, is_alt, slot, jp)) {
                        double p, q;
                        guard_pq(jp, p, q);
                        const double rq = 1.0 / q;
                        e = make_double2(p * s * rq, rq);
                        qprod *= q;
                        if (qprod < 1e-200) {
                            lq_acc += log(qprod);
                            qprod = 1.0;
                        }
                    }
                }
                sh.par[w][lane] = e;
                if (blk + 1 < nblk) fetch(blk + 1);
            }
            // the boundary cells of this block arrive in the producing warp's scale: bring them to ours, and if they are
            // far above our own cells (the bulk of the distribution is still below this warp) adopt their scale
            const XlEdge *ein = &sh.edge[w > 0 ? w - 1 : 0][blk & 1];
            XlEdge *eout = &sh.edge[w][blk & 1];
            double f_in = 0.0;
            if (w > 0) {
                const int e_in = ein->e2;
                if (e_in - e2 > 200) {
                    const int sh_ = min(e_in - e2, 2000);
                    const double f = sh_ > 1000 ? 0.0 : __hiloint2double((1023 - sh_) << 20, 0);
#pragma unroll
                    for (int r = 0; r < R; ++r) E[r] *= f;
                    T *= f;
                    e2 = e_in;
                }
                const int d = e_in - e2;               // <= 200
                f_in = d < -1000 ? 0.0 : __hiloint2double((1023 + d) << 20, 0);
            }
            __syncwarp();
            const double2 *pp = sh.par[w];
            double2 c_next = pp[0];
            double in_next = __shfl_up_sync(FULL, E[R - 1], 1);
            double edge_next = (w > 0 && lane == 0) ? ein->v[0] : 0.0;
#pragma unroll 2
            for (int j = 0; j < 32; ++j) {
                const double2 cc = c_next;
                const double in = lane == 0 ? edge_next * f_in : in_next;
                c_next = pp[(j + 1) & 31];
                if (w > 0 && lane == 0) edge_next = ein->v[(j + 1) & 31];
                const double top = E[R - 1];
                if (!is_last && lane == 31) eout->v[j] = top;
                if (is_last) T = fma(top, cc.x, T * cc.y);        // warp-uniform: only the last warp owns the absorbing state
                E[R - 1] = fma(E[R - 2], cc.x, top);
                in_next = __shfl_up_sync(FULL, E[R - 1], 1);
#pragma unroll
                for (int r = R - 2; r >= 1; --r) E[r] = fma(E[r - 1], cc.x, E[r]);
                E[0] = fma(in, cc.x, E[0]);
            }
            if (!is_last && lane == 31) eout->e2 = e2;
            // exact power-of-two rescaling, per warp
            int hi = 0;
#pragma unroll
            for (int r = 0; r < R; ++r) hi = max(hi, __double2hiint(E[r]));
            if (is_last && lane == 31) hi = max(hi, __double2hiint(T));
            hi = __reduce_max_sync(FULL, hi);
            const int ex = (hi >> 20) - 1023;
            if (hi > 0 && (ex > 200 || ex < -200)) {
                const double f = __hiloint2double((1023 - ex) << 20, 0);
#pragma unroll
                for (int r = 0; r < R; ++r) E[r] *= f;
                T *= f;
                e2 += ex;
            }
            // early exit (snpcaller.c:916-958), decided by the warp that owns the absorbing state
            if (is_last) {
                const double lq_lb = warp_sum(lq_acc + xl_ln_lower(qprod));
                const bool over = lane == 31 && T > 1e-300 && xl_ln_lower(T) + (double)e2 * LN2 + lq_lb > thr_ln;
                if (__shfl_sync(FULL, (int)over, 31)) {
                    ruled_out = true;
                    if (lane == 0) st_volatile(&sh.stop, 1);
                }
            }
            __syncwarp();
            __threadfence_block();
            if (lane == 0) st_volatile(&sh.done[w], blk + 1);
            if (ruled_out) break;
        }
    }
    __syncthreads();
    // results live in the last active warp
    XlOut out;
    int hiE = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) hiE = max(hiE, __double2hiint(E[r]));
    hiE = __reduce_max_sync(FULL, hiE);
    const double sum_lq = __shfl_sync(FULL, warp_sum(lq_acc + log(qprod)), 0);
    if (is_last && lane == 31) {
        const int hiT = __double2hiint(T);
        sh.bcast[0] = log(T) + (double)e2 * LN2 + sum_lq - (double)K * ln_s;
        sh.bcast[1] = log(E[R - 1]) + (double)e2 * LN2 + sum_lq - (double)(K - 1) * ln_s;
        sh.bcast[2] = (double)((max(hiE, hiT) >> 20) - (hiT >> 20));
        sh.bcast[3] = ruled_out ? 1.0 : 0.0;
    }
    __syncthreads();
    out.lnT = sh.bcast[0];
    out.lnKm1 = sh.bcast[1];
    out.gap = (int)sh.bcast[2];
    out.ruled_out = sh.bcast[3] != 0.0 || ld_volatile(&sh.stop) != 0;
    __syncthreads();
    return out;
}

__device__ XlOut xl_run_any(const DevConf &cf, const DevBatch &b, const Geom &g, int K, double ln_s, double thr_ln, XlSm &sh)
{
    if (K <= 8 * 32 * 16) return xl_run<16>(cf, b, g, K, ln_s, thr_ln, sh);
    if (K <= 8 * 32 * 32) return xl_run<32>(cf, b, g, K, ln_s, thr_ln, sh);
    return xl_run<64>(cf, b, g, K, ln_s, thr_ln, sh);
}

// tilt for count K from the CTA-wide histogram in sh.hsum / sh.hcnt (warp 0 solves, everybody gets the result)
__device__ double xl_tilt(int K, int N, double lam, XlSm &sh)
{
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        float cb[XL_NHIST / 32], pb[XL_NHIST / 32];
#pragma unroll
        for (int k = 0; k < XL_NHIST / 32; ++k) {
            const int c = sh.hcnt[lane + 32 * k];
            cb[k] = (float)c;
            pb[k] = c ? fminf(sh.hsum[lane + 32 * k] / (float)c, 1.f) : 0.f;
        }
        const double kt = fmin((double)K, (double)N - 0.5);
        const double s0 = kt * fmax((double)N - lam, 1e-300) / (fmax(lam, 1e-300) * fmax((double)N - kt, 0.5));
        double lo = 0.0, hi = 60.0;
        double ls = fmin(log(fmax(s0, 1.0)), hi);
        for (int it = 0; it < 40; ++it) {
            const float sf = (float)exp(ls);
            float gs = 0.f, ds = 0.f;
#pragma unroll
            for (int k = 0; k < XL_NHIST / 32; ++k) {
                const float ps = pb[k] * sf;
                const float wv = __fdividef(ps, fmaxf(1.f - pb[k], 1e-30f) + ps);
                gs = fmaf(cb[k], wv, gs);
                ds = fmaf(cb[k] * wv, 1.f - wv, ds);
            }
            const double gsum = __shfl_sync(FULL, warp_sum((double)gs), 0) - kt;
            const double d = __shfl_sync(FULL, warp_sum((double)ds), 0);
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
        if (lane == 0) sh.bcast[0] = ls;
    }
    __syncthreads();
    const double r = sh.bcast[0];
    __syncthreads();
    return r;
}

// ln P(X >= K) for one count: tilt when the Chernoff exponent asks for it, again when the untilted run shows it was needed
__device__ XlOut xl_tail(const DevConf &cf, const DevBatch &b, const Geom &g, int K, int N, double lam, double thr_base, XlSm &sh, int &flags)
{
    const double cher = ((double)K > lam) ? ((double)K * log((double)K / fmax(lam, 1e-300)) - (double)K + lam) : 0.0;
    double ln_s = cher > 300.0 ? xl_tilt(K, N, lam, sh) : 0.0;
    XlOut o = xl_run_any(cf, b, g, K, ln_s, thr_base + (double)K * ln_s, sh);
    if (!o.ruled_out && o.gap > 580 && ln_s == 0.0) {
        ln_s = xl_tilt(K, N, lam, sh);
        o = xl_run_any(cf, b, g, K, ln_s, thr_base + (double)K * ln_s, sh);
    }
    if (!o.ruled_out && o.gap > 900) flags |= CF_RANGE;
    return o;
}

__global__ void __launch_bounds__(XLT, 1) k_xl(const __grid_constant__ DevConf cf, const __grid_constant__ DevBatch b, const Lut *lut,
                                               const Workspace ws)
{
    extern __shared__ __align__(16) unsigned char xl_smem[];
    XlSm &sh = *reinterpret_cast<XlSm *>(xl_smem);
    const unsigned njobs = ws.counters->n_jobs[CLS_XL];
    if (njobs == 0) return;
    load_lut(sh.lut, lut);
    const int *jobs = ws.jobs + (long long)CLS_XL * ws.cap_cols;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    // every CTA works out the same longest-first order (cost ~ reads x cells); a long list balances by itself
    const bool sorted = njobs <= XL_SORT_MAX;
    if (sorted) {
        for (unsigned j = tid; j < njobs; j += XLT) {
            const long long c = jobs[j];
            const int4 nt = __ldg(reinterpret_cast<const int4 *>(b.nt_cnt) + c);
            const int k = max(ws.cnt6[6 * c], max(ws.cnt6[6 * c + 1], ws.cnt6[6 * c + 2]));
            sh.cost[j] = (float)(nt.x + nt.y + nt.z + nt.w) * (float)k;
        }
        __syncthreads();
        for (unsigned j = tid; j < njobs; j += XLT) {
            const float mine = sh.cost[j];
            unsigned r = 0;
            for (unsigned i = 0; i < njobs; ++i) {
                const float o = sh.cost[i];
                r += (o > mine || (o == mine && i < j)) ? 1u : 0u;
            }
            sh.order[r] = (unsigned short)j;
        }
    }
    for (;;) {
        __syncthreads();
        if (tid == 0) sh.job = atomicAdd(&ws.counters->next_job[CLS_XL], 1u);
        __syncthreads();
        const unsigned j = sh.job;
        if (j >= njobs) break;
        const long long c = jobs[sorted ? sh.order[j] : j];
        Geom g;
        int cov;
        load_geom(b, c, g, cov);
        if (cf.alt_bq_mode) {
            if (tid < 32) setup_alt_bq(cf, b, sh.lut, g, sh.med_hist);
            if (tid == 0) sh.bcast[0] = g.alt_bp;
            __syncthreads();
            g.alt_bp = sh.bcast[0];
            __syncthreads();
        }
        int cnt[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) cnt[i] = ws.cnt6[6 * c + i];
        const int K = max(cnt[0], max(cnt[1], cnt[2]));
        const long long bonf = bonf_of(cf, ws.counters->bonf_start_used, col_rank(ws, c));
        Cand cd;
        cd.flags = 0;
        cd.lnp[0] = cd.lnp[1] = cd.lnp[2] = 0.0;
        cd.ln_floor = 0.0;
        bool site = true, handed_on = false;
        if (K > XLW * 32 * 64) {
            // no kernel of this build takes an alt count this large: the column is reported as a site whose alleles carry
            // the status "unsupported" (never silently skipped), every other column of the batch is unaffected
            cd.flags = CF_UNSUPPORTED;
        } else {
            // ---- pre-pass: reads kept, lambda, largest step parameters, histogram for the tilt, small tails
            for (int i = tid; i < XL_NHIST; i += XLT) { sh.hsum[i] = 0.f; sh.hcnt[i] = 0; }
            __syncthreads();
            bool use_small[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) use_small[i] = cnt[i] > 0 && cnt[i] <= KS && cnt[i] < K;
            const bool want_small = use_small[0] || use_small[1] || use_small[2];
            double P8[KS], T8 = 0.0;
#pragma unroll
            for (int k = 0; k < KS; ++k) P8[k] = (k == 0) ? 1.0 : 0.0;
            int Nl = 0, hi_o = 0, hi_rq = 0;
            double laml = 0.0;
            const long long abase = g.off & ~15ll;
            const int lead = (int)(g.off - abase);
            const int nch = (lead + g.n + 15) >> 4;
            for (int i = tid; i < nch; i += XLT) {
                Chunk16 ch;
                load_chunk(cf, b, abase + 16ll * i, ch);
#pragma unroll 1
                for (int k = 0; k < 16; ++k) {
                    const int pos = 16 * i - lead + k;
                    if (pos < 0 || pos >= g.n) continue;
                    bool is_alt;
                    int slot;
                    double jp;
                    if (!eval_read<true>(cf, sh.lut, g, pos, byte_of(ch.bq, k), byte_of(ch.mq, k), byte_of(ch.baq, k), byte_of(ch.sq, k), is_alt, slot, jp))
                        continue;
                    double p, q;
                    guard_pq(jp, p, q);
                    laml += p;
                    ++Nl;
                    const double rq = 1.0 / q;
                    hi_o = max(hi_o, __double2hiint(p * rq));
                    hi_rq = max(hi_rq, __double2hiint(rq));
                    int bk = (0x3ff00000 - __double2hiint(p)) >> 18;
                    bk = min(max(bk, 0), XL_NHIST - 1);
                    atomicAdd(&sh.hsum[bk], (float)p);
                    atomicAdd(&sh.hcnt[bk], 1);
                    if (want_small) {
                        T8 = fma(P8[KS - 1], p, T8);
#pragma unroll
                        for (int kk = KS - 1; kk >= 1; --kk) P8[kk] = fma(P8[kk - 1], p, P8[kk] * q);
                        P8[0] = P8[0] * q;
                    }
                }
            }
            const int N = (int)(xl_block_sum((double)Nl, sh) + 0.5);
            const double lam = xl_block_sum(laml, sh);
            const double max_o = __hiloint2double(xl_block_max(hi_o, sh), (int)0xffffffff);
            const double max_rq = __hiloint2double(xl_block_max(hi_rq, sh), (int)0xffffffff);
            double small_tail[3] = {0.0, 0.0, 0.0};
            if (want_small) {
                // merge the 256 per-thread distributions: butterfly inside each warp, then warp 0 folds the 8 warp results
#pragma unroll 1
                for (int m = 1; m < 32; m <<= 1) {
                    double bb[KS], cc[KS];
                    const double tb = __shfl_xor_sync(FULL, T8, m);
                    double sum_a = 0.0, sum_b = 0.0;
#pragma unroll
                    for (int k = 0; k < KS; ++k) { bb[k] = __shfl_xor_sync(FULL, P8[k], m); sum_a += P8[k]; sum_b += bb[k]; }
#pragma unroll
                    for (int k = 0; k < KS; ++k) {
                        double acc = 0.0;
#pragma unroll
                        for (int i = 0; i <= k; ++i) acc = fma(P8[i], bb[k - i], acc);
                        cc[k] = acc;
                    }
                    double t = T8 * (sum_b + tb) + tb * sum_a, asuf = 0.0;
#pragma unroll
                    for (int jj = 1; jj < KS; ++jj) { asuf += P8[KS - jj]; t = fma(bb[jj], asuf, t); }
#pragma unroll
                    for (int k = 0; k < KS; ++k) P8[k] = cc[k];
                    T8 = t;
                }
                if (lane == 0) {
#pragma unroll
                    for (int k = 0; k < KS; ++k) sh.small_P[w][k] = P8[k];
                    sh.small_T[w] = T8;
                }
                __syncthreads();
                if (tid == 0) {
                    double A[KS], TA = sh.small_T[0];
                    for (int k = 0; k < KS; ++k) A[k] = sh.small_P[0][k];
                    for (int ww = 1; ww < XLW; ++ww) {
                        double C[KS], sa = 0.0, sb = 0.0;
                        const double *B = sh.small_P[ww];
                        const double TB = sh.small_T[ww];
                        for (int k = 0; k < KS; ++k) { sa += A[k]; sb += B[k]; }
                        for (int k = 0; k < KS; ++k) {
                            double acc = 0.0;
                            for (int i = 0; i <= k; ++i) acc = fma(A[i], B[k - i], acc);
                            C[k] = acc;
                        }
                        double t = TA * (sb + TB) + TB * sa, asuf = 0.0;
                        for (int jj = 1; jj < KS; ++jj) { asuf += A[KS - jj]; t = fma(B[jj], asuf, t); }
                        for (int k = 0; k < KS; ++k) A[k] = C[k];
                        TA = t;
                    }
                    for (int k = 0; k < KS; ++k) sh.small_P[0][k] = A[k];
                    sh.small_T[0] = TA;
                }
                __syncthreads();
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    double tl = sh.small_T[0];
                    for (int k = KS - 1; k >= 0; --k)
                        if (k >= cnt[i]) tl += sh.small_P[0][k];
                    small_tail[i] = tl;
                }
                __syncthreads();
            }
            // the largest tilt any run of this column will use is the one for K: parameters must stay below 2^20 under it
            const double cherK = ((double)K > lam) ? ((double)K * log((double)K / fmax(lam, 1e-300)) - (double)K + lam) : 0.0;
            (void)cherK;
            const double s_max = exp(xl_tilt(K, N, lam, sh));
            if (max_o * s_max > 1048576.0 || max_rq > 1048576.0) {
                // rescaling every 32 reads cannot hold this column: the CTA-per-column kernel that rescales after every read
                if (tid == 0) {
                    const unsigned slot = atomicAdd(&ws.counters->n_jobs[CLS_XLFB], 1u);
                    ws.jobs[(long long)CLS_XLFB * ws.cap_cols + slot] = (int)c;
                }
                handed_on = true;
            } else {
                const double thr_base = log(cf.sig * (1.0 + 1e-9) / (double)bonf);
                const XlOut main_t = xl_tail(cf, b, g, K, N, lam, thr_base, sh, cd.flags);
                if (main_t.ruled_out) site = false;
                else if (main_t.lnT > -700.0 && exp(main_t.lnT) * (double)bonf > cf.sig * (1.0 + 1e-9)) site = false;
                if (site) {
                    cd.ln_floor = fmin(main_t.lnT, main_t.lnKm1);
#pragma unroll 1
                    for (int i = 0; i < 3; ++i) {
                        const int cc = cnt[i];
                        if (cc == 0) continue;
                        if (cc == K) { cd.lnp[i] = main_t.lnT; continue; }
                        if (use_small[i]) { cd.lnp[i] = log(small_tail[i]); continue; }
                        if (i == 2 && cc == cnt[1]) { cd.lnp[i] = cd.lnp[1]; continue; }
                        if (i >= 1 && cc == cnt[0]) { cd.lnp[i] = cd.lnp[0]; continue; }
                        // a further allele with a count above 8: its own run, early exit off
                        const XlOut t2 = xl_tail(cf, b, g, cc, N, lam, INFINITY, sh, cd.flags);
                        cd.lnp[i] = t2.lnT;
                    }
                }
            }
        }
        if (handed_on || !site) continue;
        if (tid == 0) {
            cd.col = c;
            cd.bonf = bonf;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                cd.cnt[i] = cnt[i];
                cd.raw[i] = ws.cnt6[6 * c + 3 + i];
            }
            cd.pad = 0;
            const unsigned slot = atomicAdd(&ws.counters->n_cand, 1u);
            ws.cand[slot] = cd;
            mark_cand(ws, c);
        }
    }
}

int xl_smem_optin()
{
    return cudaFuncSetAttribute(k_xl, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(XlSm)) != cudaSuccess;
}

void launch_xl(const LaunchState &ls, const DevConf &cf, const DevBatch &b, const Lut *lut, const Workspace &ws, cudaStream_t st)
{
    if (b.n_cols <= 0) return;
    k_xl<<<ls.sms, XLT, sizeof(XlSm), st>>>(cf, b, lut, ws);
}

}  // namespace lfb
