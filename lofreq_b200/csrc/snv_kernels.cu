// sm_100a kernels of the per-pileup-column SNV test.
//
// What the reference does per column (lofreq_call.c:734-879 -> snpcaller.c):
//   plp_to_errprobs (snpcaller.c:345-498)  quality bytes -> merged error probabilities, alt counts
//   qsort                                   (order only changes rounding and where pruning fires)
//   snpcaller -> poissbin -> pruned_calc_prob_dist (snpcaller.c:830-1204)
//                                           Poisson-binomial DP in log space, tail p-value per allele
//
// What runs here (DESIGN.md has the whole picture; dp_fused.cu, poissbin.cu, mailbox.cu, binom.cu, fisher.cu hold the rest):
//   k_front (front.cu)
//               gates and alt counts (only the reads showing a non-reference base are looked at), tested flags, the
//               place of a column in the running Bonferroni count (lofreq_call.c:794-800), the first stage of the
//               reference's early exit one lane per column, job lists for everything that survives.
//   k_prune2    second stage of the early exit with the exact factor.
//   k_mid       K <= 8 survivors: the exact distribution truncated at K, evaluated in linear space on 32 disjoint read
//               subsets and merged by truncated convolution over warp shuffles.
//   k_dp (dp_fused.cu), k_xl (xl.cu)
//               the O(depth*K) recurrence for 8 < K <= 2048 (several columns per warp) and K > 2048 (one CTA per column).
//   k_heavy_all / k_heavy_xl
//               the per-column fallbacks for what those hand back (one warp / one CTA per column): K cells tiled over
//               lanes x R registers, one shuffle per read for the lane boundary, odds form E[k] += E[k-1]*o (one DFMA
//               per cell), rescaling after every read if need be, exponential tilting after the fact.
//   k_emit_sites  status / called / QUAL per site on the device, records in column order into pinned host memory.
//   k_rank_cands  the sites in column order (a permutation), so that the host needs no sort.
//   k_prob_jobs the same routines fed with ready-made double error probabilities (snpcaller() symbol).
//
// All sums are over positive terms, so the relative error of a tail is O(depth * 2^-53); the reference's
// log-space arithmetic (exp + log1p per cell) is not reproduced, its results are (tests/test_parity_gpu.py).
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <float.h>
#include "internal.h"
#include "dev_common.cuh"
#include "screen_common.cuh"

namespace lfb {
// ------------------------------------------------------------------------------------------------
// distribution truncated at K (K <= KS): P[k] = P(k errors), k < K; T = P(>= K errors)
// ------------------------------------------------------------------------------------------------
template <int K>
__device__ __forceinline__ void lane_update(double (&P)[K], double &T, double p, double q)
{
    T = fma(P[K - 1], p, T);
#pragma unroll
    for (int k = K - 1; k >= 1; --k) P[k] = fma(P[k - 1], p, P[k] * q);
    P[0] = P[0] * q;
}

// butterfly merge of the 32 per-lane distributions; every lane ends with the distribution of all reads
template <int K>
__device__ __forceinline__ void tree_merge(double (&P)[K], double &T)
{
#pragma unroll 1
    for (int m = 1; m < 32; m <<= 1) {
        double b[K], c[K];
        const double tb = __shfl_xor_sync(FULL, T, m);
        double sum_a = 0.0, sum_b = 0.0;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            b[k] = __shfl_xor_sync(FULL, P[k], m);
            sum_a += P[k];
            sum_b += b[k];
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
            double acc = 0.0;
#pragma unroll
            for (int i = 0; i <= k; ++i) acc = fma(P[i], b[k - i], acc);
            c[k] = acc;
        }
        double t = T * (sum_b + tb) + tb * sum_a;
        double asuf = 0.0;
#pragma unroll
        for (int j = 1; j < K; ++j) {
            asuf += P[K - j];
            t = fma(b[j], asuf, t);
        }
#pragma unroll
        for (int k = 0; k < K; ++k) P[k] = c[k];
        T = t;
    }
}

// Same merge for wide distributions (K up to 32): the partner's cells are fetched one at a time inside the
// convolution instead of being held in registers.
template <int K>
__device__ __forceinline__ void tree_merge_stream(double (&P)[K], double &T)
{
#pragma unroll 1
    for (int m = 1; m < 32; m <<= 1) {
        double c[K];
        const double tb = __shfl_xor_sync(FULL, T, m);
        double sum_a = 0.0, sum_b = 0.0;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            c[k] = 0.0;
            sum_a += P[k];
        }
        double t = 0.0, asuf = 0.0;
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const double bj = __shfl_xor_sync(FULL, P[j], m);
            sum_b += bj;
#pragma unroll
            for (int k = j; k < K; ++k) c[k] = fma(P[k - j], bj, c[k]);
            if (j >= 1) {
                asuf += P[K - j];
                t = fma(bj, asuf, t);
            }
        }
        t += T * (sum_b + tb) + tb * sum_a;
#pragma unroll
        for (int k = 0; k < K; ++k) P[k] = c[k];
        T = t;
    }
}

// The distribution was truncated at KV >= K = max count.  tails[i] = P(X >= cnt[i]) (0 when cnt[i] == 0),
// tails[3] = min(P(X = K-1), P(X >= K)).
template <int KV>
__device__ __forceinline__ void small_tails(const double (&P)[KV], double T, const int (&cnt)[3], int K, double (&tails)[4])
{
    double t0 = T, t1 = T, t2 = T, tk = T, pk1 = 0.0;
#pragma unroll
    for (int k = KV - 1; k >= 0; --k) {      // small terms first
        if (k >= cnt[0]) t0 += P[k];
        if (k >= cnt[1]) t1 += P[k];
        if (k >= cnt[2]) t2 += P[k];
        if (k >= K) tk += P[k];
        if (k == K - 1) pk1 = P[k];
    }
    tails[0] = cnt[0] > 0 ? t0 : 0.0;
    tails[1] = cnt[1] > 0 ? t1 : 0.0;
    tails[2] = cnt[2] > 0 ? t2 : 0.0;
    tails[3] = fmin(pk1, tk);
}

// ------------------------------------------------------------------------------------------------
// read evaluation shared by k_mid and the fallbacks
// ------------------------------------------------------------------------------------------------
// merged error probability of a read that shows the reference base and passed the bq filter
// (merge_srcq_mapq_baq_and_bq with the terms of absent planes dropped: x*0, +0 and *1 are exact)
__device__ __forceinline__ double ref_read_prob(const DevConf &cf, const double *lut, int bq, int mq, int baq, int sq)
{
    const double bp = lut[bq];
    if (cf.use_baq | cf.use_sq)
        return merge4(cf.use_sq ? lut[512 + sq] : 0.0, cf.use_mq ? lut[256 + mq] : 0.0, cf.use_baq ? lut[512 + baq] : 0.0, bp);
    if (!cf.use_mq) return bp;
    const double mp = lut[256 + mq];
    return __dadd_rn(mp, __dmul_rn(__dsub_rn(1.0, mp), bp));
}

__device__ __forceinline__ double lds_f64(unsigned saddr)
{
    double v;
    asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(saddr));
    return v;
}

// What the per-read loop needs from DevConf, hoisted into registers once per column.
struct ReadRules {
    int min_bq, min_alt;           // bq thresholds for reference / alt reads
    double skip, skip_alt;         // merged-probability cut-offs (+inf when the jq filters are off)
    bool use_mq, general_merge, use_baq, use_sq, alt_bq, alt_jq;
    double alt_bp, alt_jp;
};

// Fold the reads of one 16-byte chunk into the lane's truncated distribution.
// UNIFORM: the configuration treats reference and alt reads alike in this sweep (defaults: min_alt_bq <=
// min_bq, no def_alt_bq / def_alt_jq, no jq filters, min_bq >= 1) — bytes outside the column were zeroed by
// the caller, so the bq filter drops them and no position bookkeeping is needed.  With bq >= 1 every
// probability is <= 0.96, so the q-guard of snpcaller.c:877 cannot fire and the p-guard is max(p, DBL_EPSILON).
template <int KV, bool UNIFORM>
__device__ __forceinline__ void fold_chunk(const ReadRules &rr, unsigned lut_sa, int n, int ref_lo, int ref_hi, int pos0,
                                           const Chunk16 &ch, double (&P)[KV], double &T)
{
#pragma unroll 1
    for (int w = 0; w < 4; ++w) {
        unsigned wbq = w == 0 ? ch.bq.x : w == 1 ? ch.bq.y : w == 2 ? ch.bq.z : ch.bq.w;
        unsigned wmq = w == 0 ? ch.mq.x : w == 1 ? ch.mq.y : w == 2 ? ch.mq.z : ch.mq.w;
        if (UNIFORM) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int bq = (wbq >> (8 * j)) & 0xff;
                if (bq < rr.min_bq) continue;
                const double bp = lds_f64(lut_sa + 8 * bq);
                double jp = bp;
                if (rr.use_mq) {
                    const double mp = lds_f64(lut_sa + 2048 + 8 * ((wmq >> (8 * j)) & 0xff));
                    jp = __dadd_rn(mp, __dmul_rn(__dsub_rn(1.0, mp), bp));   // sp = bap = 0: the dropped terms are exact
                }
                const double q = 1.0 - jp;
                if (jp < DEPS) jp = DEPS;
                lane_update<KV>(P, T, jp, q);
            }
        } else {
            unsigned wbaq = w == 0 ? ch.baq.x : w == 1 ? ch.baq.y : w == 2 ? ch.baq.z : ch.baq.w;
            unsigned wsq = w == 0 ? ch.sq.x : w == 1 ? ch.sq.y : w == 2 ? ch.sq.z : ch.sq.w;
#pragma unroll 1
            for (int j = 0; j < 4; ++j, wbq >>= 8, wmq >>= 8, wbaq >>= 8, wsq >>= 8) {
                const int pos = pos0 + 4 * w + j;
                const int bq = wbq & 0xff;
                const bool is_alt = pos < ref_lo || pos >= ref_hi;
                if (pos < 0 || pos >= n || bq < (is_alt ? rr.min_alt : rr.min_bq)) continue;
                double bp = lds_f64(lut_sa + 8 * bq);
                if (is_alt && rr.alt_bq) bp = rr.alt_bp;
                const double mp = rr.use_mq ? lds_f64(lut_sa + 2048 + 8 * (wmq & 0xff)) : 0.0;
                double jp;
                if (rr.general_merge)
                    jp = merge4(rr.use_sq ? lds_f64(lut_sa + 4096 + 8 * (wsq & 0xff)) : 0.0, mp,
                                rr.use_baq ? lds_f64(lut_sa + 4096 + 8 * (wbaq & 0xff)) : 0.0, bp);
                else
                    jp = __dadd_rn(mp, __dmul_rn(__dsub_rn(1.0, mp), bp));
                if (jp >= (is_alt ? rr.skip_alt : rr.skip)) continue;
                if (is_alt && rr.alt_jq) jp = rr.alt_jp;
                double q = 1.0 - jp;
                if (jp < DEPS || fabs(q) < DEPS) guard_pq(jp, jp, q);
                lane_update<KV>(P, T, jp, q);
            }
        }
    }
}

// zero the bytes of a chunk that lie outside [0, n) of the column (UNIFORM path: bq 0 is filtered out)
__device__ __forceinline__ void mask_chunk(uint4 &v, int pos0, int n)
{
    if (pos0 >= 0 && pos0 + 16 <= n) return;
    unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        unsigned m = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int pos = pos0 + 4 * k + j;
            if (pos >= 0 && pos < n) m |= 0xffu << (8 * j);
        }
        w[k] &= m;
    }
    v = make_uint4(w[0], w[1], w[2], w[3]);
}

// Full evaluation of a column with K <= KS: every lane folds the reads of its 16-byte chunks into a
// distribution truncated at KV >= K, then the 32 distributions are merged.  Only columns that survive the
// 32-read prune of k_prune2 get here (true low-frequency variants, and the first few columns of a run
// whose Bonferroni factor is still small).
template <int KV>
__device__ __noinline__ void screen_small(const DevConf &cf, const DevBatch &b, unsigned lut_sa, const Geom &g,
                                          const int (&cnt)[3], int K, double (&tails)[4])
{
    const int lane = lane_id();
    double P[KV], T = 0.0;
#pragma unroll
    for (int k = 0; k < KV; ++k) P[k] = (k == 0) ? 1.0 : 0.0;
    const long long abase = g.off & ~15ll;
    const int lead = (int)(g.off - abase);
    const int nchunks = (lead + g.n + 15) >> 4;
    int ref_lo, ref_hi;
    ref_range(g, ref_lo, ref_hi);
    ReadRules rr;
    rr.min_bq = cf.min_bq;
    rr.min_alt = max(cf.min_bq, cf.min_alt_bq);
    rr.skip = cf.skip_jp;
    rr.skip_alt = fmin(cf.skip_jp, cf.skip_alt_jp);
    rr.use_mq = cf.use_mq;
    rr.use_baq = cf.use_baq;
    rr.use_sq = cf.use_sq;
    rr.general_merge = cf.use_baq | cf.use_sq;
    rr.alt_bq = cf.alt_bq_mode != 0;
    rr.alt_bp = g.alt_bp;
    rr.alt_jq = cf.def_alt_jq_on != 0;
    rr.alt_jp = cf.def_alt_jq_prob;
    const bool uniform = cf.min_bq >= 1 && cf.min_alt_bq <= cf.min_bq && !rr.alt_bq && !rr.alt_jq && !cf.jq_filters &&
                         !rr.general_merge;
#pragma unroll 1
    for (int i = lane; i < nchunks; i += 32) {
        Chunk16 ch;
        load_chunk(cf, b, abase + 16ll * i, ch);
        const int pos0 = 16 * i - lead;
        if (uniform) {
            mask_chunk(ch.bq, pos0, g.n);
            fold_chunk<KV, true>(rr, lut_sa, g.n, ref_lo, ref_hi, pos0, ch, P, T);
        } else {
            fold_chunk<KV, false>(rr, lut_sa, g.n, ref_lo, ref_hi, pos0, ch, P, T);
        }
    }
    if (KV > 8) tree_merge_stream<KV>(P, T); else tree_merge<KV>(P, T);
    small_tails<KV>(P, T, cnt, K, tails);
}

// ------------------------------------------------------------------------------------------------
// running Bonferroni: prefix sum over tested flags, then the significance screen
// ------------------------------------------------------------------------------------------------

// exclusive scan of per-tile counts (the candidate marks: k_rank_cands), one block; the counts are zeroed for the next batch
__global__ void __launch_bounds__(1024) k_scan_blocks(unsigned int *tilecount, long long *blocksum, int nb, unsigned long long *total)
{
    __shared__ long long s_warp[32];
    __shared__ long long s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = lane_id(), w = threadIdx.x >> 5;
    for (int base = 0; base < nb; base += 1024) {
        const int i = base + threadIdx.x;
        const long long v = (i < nb) ? (long long)tilecount[i] : 0;
        if (i < nb) tilecount[i] = 0;
        long long x = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const long long y = __shfl_up_sync(FULL, x, d);
            if (lane >= d) x += y;
        }
        if (lane == 31) s_warp[w] = x;
        __syncthreads();
        if (w == 0) {
            long long z = s_warp[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const long long y = __shfl_up_sync(FULL, z, d);
                if (lane >= d) z += y;
            }
            s_warp[lane] = z;
        }
        __syncthreads();
        const long long carry = s_carry;
        const long long incl = x + (w ? s_warp[w - 1] : 0) + carry;
        if (i < nb) blocksum[i] = incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = incl;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total) *total = (unsigned long long)s_carry;
}

// (2) second stage of the prune: the columns k_front could not rule out within PRUNE_CAP1 reads, one lane each, up to
//     PRUNE_CAP reads (from the first read again: eight reads are cheaper to redo than to carry).  The survivors (true
//     low-frequency variants, the first columns of a run) join k_mid's job list.
__global__ void __launch_bounds__(128) k_prune2(const __grid_constant__ DevConf cf, const __grid_constant__ DevBatch b, const Lut *lut,
                                                const Workspace ws, const long long *bonf_start_dev)
{
    __shared__ double s_lut[768];
    // the exact factor this batch continues from: k_front used the caller's conf; region shards on other GPUs may have
    // added to it since (lfb200_comm_exchange leaves the sum in device memory, no host round trip)
    const long long start = bonf_start_dev ? *bonf_start_dev : cf.bonf_start;
    const unsigned njobs = ws.counters->n_jobs[CLS_PRUNE2];
    if (njobs == 0) return;
    load_lut(s_lut, lut);
    const int *jobs = ws.jobs + (long long)CLS_PRUNE2 * ws.cap_cols;
    for (unsigned j0 = blockIdx.x * blockDim.x; j0 < njobs; j0 += gridDim.x * blockDim.x) {
        const unsigned j = j0 + threadIdx.x;
        bool live = j < njobs;
        const long long c = live ? jobs[j] : 0;
        Geom mg;
        mg.off = 0; mg.b1 = mg.b2 = mg.b3 = mg.n = 0; mg.ref_idx = -1; mg.alt_bp = cf.alt_bq_prob;
        int K = 1;
        double limit = 0.0;
        if (live) {
            int cov;
            load_geom(b, c, mg, cov);
            mg.alt_bp = cf.alt_bq_prob;
            const int2 *in = reinterpret_cast<const int2 *>(ws.cnt6 + 6 * c);
            const int2 a0 = in[0], a1 = in[1];
            K = max(a0.x, max(a0.y, a1.x));
            limit = cf.sig * (1.0 + 1e-9) / (double)bonf_of(cf, start, col_rank(ws, c));
        }
        live = lane_prune<KS>(cf, b, s_lut, mg, K, limit, PRUNE_CAP, live, nullptr, true);
        if (live) {
            const unsigned slot = atomicAdd(&ws.counters->n_jobs[0], 1u);
            ws.jobs[slot] = (int)c;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// read sources for the O(depth*K) routines
// ------------------------------------------------------------------------------------------------
// Every pass over a source walks it tile by tile: `stage(t0)` makes reads [t0, t0 + TILE) available to get().
// The raw sources read global memory on every get() (TILE = everything, stage is a no-op); Staged<> keeps one
// tile of merged probabilities in shared memory, so the repeated passes of the O(depth*K) routines (lambda,
// Newton iterations, one or more runs of the recurrence, secondary alleles) pay the global-memory latency once
// per tile instead of once per 32 reads and pass.
constexpr int NO_TILE = 0x40000000;

struct ByteSrc {
    static constexpr int TILE = NO_TILE;
    static constexpr bool FILTERS = true;
    const DevConf *cf;
    const DevBatch *b;
    const double *lut;
    Geom g;
    __device__ __forceinline__ int size() const { return g.n; }
    __device__ __forceinline__ void stage(int) const {}
    __device__ __forceinline__ bool get(int pos, double &jp) const
    {
        const long long a = g.off + pos;
        bool is_alt;
        int slot;
        return eval_read<true>(*cf, lut, g, pos, b->bq[a], cf->use_mq ? b->mq[a] : 0, cf->use_baq ? b->baq[a] : 0,
                               cf->use_sq ? b->sq[a] : 0, is_alt, slot, jp);
    }
    // merged probabilities of reads [t0, t0 + m) into buf (swizzled, see Staged), -1 = filtered out.
    // 128-bit loads: lane i takes the i-th aligned 16-byte chunk of every plane.
    __device__ __forceinline__ void fill(int t0, int m, double *buf) const
    {
        const long long a0 = g.off + t0;
        const long long abase = a0 & ~15ll;
        const int lead = (int)(a0 - abase);
        const int nch = (lead + m + 15) >> 4;
        for (int i = lane_id(); i < nch; i += 32) {
            Chunk16 ch;
            load_chunk(*cf, *b, abase + 16ll * i, ch);
#pragma unroll 1
            for (int w = 0; w < 4; ++w) {
                const unsigned wbq = w == 0 ? ch.bq.x : w == 1 ? ch.bq.y : w == 2 ? ch.bq.z : ch.bq.w;
                const unsigned wmq = w == 0 ? ch.mq.x : w == 1 ? ch.mq.y : w == 2 ? ch.mq.z : ch.mq.w;
                const unsigned wbaq = w == 0 ? ch.baq.x : w == 1 ? ch.baq.y : w == 2 ? ch.baq.z : ch.baq.w;
                const unsigned wsq = w == 0 ? ch.sq.x : w == 1 ? ch.sq.y : w == 2 ? ch.sq.z : ch.sq.w;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int idx = 16 * i + 4 * w + j;          // chunk-relative slot
                    const int tp = idx - lead;                    // tile-relative read
                    if (tp < 0 || tp >= m) continue;
                    bool is_alt;
                    int slot;
                    double jp;
                    const bool ok = eval_read<true>(*cf, lut, g, t0 + tp, (wbq >> (8 * j)) & 0xff, (wmq >> (8 * j)) & 0xff,
                                                    (wbaq >> (8 * j)) & 0xff, (wsq >> (8 * j)) & 0xff, is_alt, slot, jp);
                    buf[idx + (idx >> 4)] = ok ? jp : -1.0;
                }
            }
        }
    }
    __device__ __forceinline__ int lead_of(int t0) const { return (int)((g.off + t0) & 15ll); }
};

struct ProbSrc {
    static constexpr int TILE = NO_TILE;
    static constexpr bool FILTERS = false;
    const double *ep;
    int n;
    __device__ __forceinline__ int size() const { return n; }
    __device__ __forceinline__ void stage(int) const {}
    __device__ __forceinline__ bool get(int pos, double &jp) const
    {
        jp = ep[pos];
        return true;
    }
    __device__ __forceinline__ void fill(int t0, int m, double *buf) const
    {
        for (int i = lane_id(); i < m; i += 32) buf[i + (i >> 4)] = ep[t0 + i];
    }
    __device__ __forceinline__ int lead_of(int) const { return 0; }
};

constexpr int STAGE_TILE = 1024;                                   // reads per tile
constexpr int STAGE_SLOTS = STAGE_TILE + 16;                       // + the lead bytes of an unaligned column
constexpr int STAGE_DOUBLES = STAGE_SLOTS + (STAGE_SLOTS >> 4) + 1;   // one pad per 16 slots: the 16-byte-chunk fill is bank-conflict free

template <class Raw>
struct Staged {
    static constexpr int TILE = STAGE_TILE;
    Raw raw;
    double *buf;        // STAGE_DOUBLES of shared memory owned by this warp
    mutable int tile0, lead;
    __device__ __forceinline__ void init(double *smem) { buf = smem; tile0 = -1; lead = 0; }
    __device__ __forceinline__ int size() const { return raw.size(); }
    __device__ __forceinline__ void stage(int t0) const
    {
        if (t0 == tile0) return;
        __syncwarp();
        raw.fill(t0, min(STAGE_TILE, raw.size() - t0), buf);
        tile0 = t0;
        lead = raw.lead_of(t0);
        __syncwarp();
    }
    __device__ __forceinline__ bool get(int pos, double &jp) const
    {
        const int idx = pos - tile0 + lead;
        jp = buf[idx + (idx >> 4)];
        return !Raw::FILTERS || jp >= 0.0;
    }
};

// distribution truncated at K <= KS from a generic source (secondary alleles of heavy columns, small
// stand-alone problems); lanes stride over the reads
template <int K, class Src>
__device__ __noinline__ void src_small(const Src &src, double (&P8)[KS], double &T)
{
    double P[K];
    T = 0.0;
#pragma unroll
    for (int k = 0; k < K; ++k) P[k] = (k == 0) ? 1.0 : 0.0;
    const int n = src.size();
    for (int t0 = 0; t0 < n; t0 += Src::TILE) {
        src.stage(t0);
        const int t1 = min(n - t0, Src::TILE) + t0;
        for (int pos = t0 + lane_id(); pos < t1; pos += 32) {
            double jp;
            if (!src.get(pos, jp)) continue;
            double p, q;
            guard_pq(jp, p, q);
            lane_update<K>(P, T, p, q);
        }
    }
    tree_merge<K>(P, T);
#pragma unroll
    for (int k = 0; k < KS; ++k) P8[k] = (k < K) ? P[k < K ? k : 0] : 0.0;
}

template <class Src>
__device__ void src_small_dispatch(const Src &src, int K, double (&P8)[KS], double &T)
{
    switch (K) {
        case 1: src_small<1>(src, P8, T); break;
        case 2: src_small<2>(src, P8, T); break;
        case 3: src_small<3>(src, P8, T); break;
        case 4: src_small<4>(src, P8, T); break;
        case 5: src_small<5>(src, P8, T); break;
        case 6: src_small<6>(src, P8, T); break;
        case 7: src_small<7>(src, P8, T); break;
        default: src_small<8>(src, P8, T); break;
    }
}

// ------------------------------------------------------------------------------------------------
// the O(depth*K) recurrence, K cells over 32 lanes x R registers
// ------------------------------------------------------------------------------------------------
template <int R>
struct Row {
    double E[R];       // odds-form cells: lane l, register r holds k = K - 32R + l*R + r (k < 0: padding, always 0)
    double T;          // absorbing state P(>= K) in the same scaling (valid on lane 31)
    int e2;            // power-of-two exponent taken out so far
    double sum_lq;     // sum over reads of ln q (whole warp)
    double ln_s;       // tilt
};

template <int R>
__device__ __forceinline__ void rescale(Row<R> &row, bool force)
{
    int hi = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) hi = max(hi, __double2hiint(row.E[r]));
    if (lane_id() == 31) hi = max(hi, __double2hiint(row.T));
    hi = __reduce_max_sync(FULL, hi);
    const int ex = (hi >> 20) - 1023;
    if (force ? (ex != 0) : (ex > 200 || ex < -200)) {
        const double f = __hiloint2double((1023 - ex) << 20, 0);
#pragma unroll
        for (int r = 0; r < R; ++r) row.E[r] *= f;
        row.T *= f;
        row.e2 += ex;
    }
}

// Runs the recurrence over every read of src with tilt exp(ln_s).  sm: 32 double2 of shared memory per warp.
template <int R, class Src>
__device__ __forceinline__ void dp_run(const Src &src, int K, double ln_s, double2 *sm, Row<R> &row)
{
    const int lane = lane_id();
    const double s = (ln_s == 0.0) ? 1.0 : exp(ln_s);
    const int k0 = K - 32 * R + lane * R;
#pragma unroll
    for (int r = 0; r < R; ++r) row.E[r] = (k0 + r == 0) ? 1.0 : 0.0;
    row.T = 0.0;
    row.e2 = 0;
    row.ln_s = ln_s;
    double lq_acc = 0.0, q_prod = 1.0;      // sum of ln q = lq_acc + ln(q_prod); the log is taken once per ~hundreds of reads
    const int n = src.size();
    const unsigned lt_mask = (1u << lane) - 1u;
    for (int n0 = 0; n0 < n; n0 += 32) {
        if (n0 % Src::TILE == 0) src.stage(n0);
        const int pos = n0 + lane;
        double jp = 0.0, o = 0.0, rq = 1.0;
        const bool ok = pos < n && src.get(pos, jp);
        if (ok) {
            double p, q;
            guard_pq(jp, p, q);
            rq = 1.0 / q;
            o = p * s * rq;
            q_prod *= q;
            if (q_prod < 1e-200) {
                lq_acc += log(q_prod);
                q_prod = 1.0;
            }
        }
        const unsigned m = __ballot_sync(FULL, ok);
        const int cnt = __popc(m);
        const bool slow = __any_sync(FULL, ok && (o > 1048576.0 || rq > 1048576.0));
        __syncwarp();
        if (ok) sm[__popc(m & lt_mask)] = make_double2(o, rq);
        __syncwarp();
        if (!slow) {
            // software-pipelined: the parameters of read j+1 and the boundary cell for read j+1 (the top cell
            // right after its own update) are requested before the remaining R-1 cells of read j are updated
            double2 c_next = sm[0];
            double in_next = __shfl_up_sync(FULL, row.E[R - 1], 1);
            for (int j = 0; j < cnt; ++j) {
                const double2 c = c_next;
                const double in = lane == 0 ? 0.0 : in_next;
                c_next = sm[(j + 1) & 31];
                const double top = row.E[R - 1];
                row.T = fma(top, c.x, row.T * c.y);
                if (R > 1) {
                    row.E[R - 1] = fma(row.E[R > 1 ? R - 2 : 0], c.x, top);
                    in_next = __shfl_up_sync(FULL, row.E[R - 1], 1);
#pragma unroll
                    for (int r = R - 2; r >= 1; --r) row.E[r] = fma(row.E[r - 1], c.x, row.E[r]);
                    row.E[0] = fma(in, c.x, row.E[0]);
                } else {
                    row.E[0] = fma(in, c.x, top);
                    in_next = __shfl_up_sync(FULL, row.E[0], 1);
                }
            }
            rescale<R>(row, false);
        } else {
            for (int j = 0; j < cnt; ++j) {
                const double2 c = sm[j];
                const double top = row.E[R - 1];
                double in = __shfl_up_sync(FULL, top, 1);
                if (lane == 0) in = 0.0;
                row.T = fma(top, c.x, row.T * c.y);
#pragma unroll
                for (int r = R - 1; r >= 1; --r) row.E[r] = fma(row.E[r - 1], c.x, row.E[r]);
                row.E[0] = fma(in, c.x, row.E[0]);
                rescale<R>(row, true);
            }
        }
    }
    lq_acc += log(q_prod);
    row.sum_lq = warp_sum(lq_acc);
}

// saddlepoint tilt: ln s with sum_n o_n/(1+o_n) = min(K, N-1/2), o_n = p_n s / q_n
template <class Src>
__device__ double newton_tilt(const Src &src, int K, int N, double lam)
{
    const int lane = lane_id();
    const double kt = fmin((double)K, (double)N - 0.5);
    const double s0 = kt * fmax((double)N - lam, 1e-300) / (fmax(lam, 1e-300) * ((double)N - kt));
    double ls = log(fmax(s0, 1.0));
    double lo = 0.0, hi = 60.0;
    ls = fmin(ls, hi);
    const int n = src.size();
    for (int it = 0; it < 40; ++it) {
        const double s = exp(ls);
        double g = 0.0, d = 0.0;
        for (int t0 = 0; t0 < n; t0 += Src::TILE) {
            src.stage(t0);
            const int t1 = min(n - t0, Src::TILE) + t0;
            for (int pos = t0 + lane; pos < t1; pos += 32) {
                double jp;
                if (!src.get(pos, jp)) continue;
                double p, q;
                guard_pq(jp, p, q);
                const double ps = p * s;
                const double w = ps / (q + ps);        // o/(1+o)
                g += w;
                d += w * (1.0 - w);                    // derivative with respect to ln s = variance of the tilted sum
            }
        }
        g = warp_sum(g) - kt;
        d = warp_sum(d);
        g = __shfl_sync(FULL, g, 0);
        d = __shfl_sync(FULL, d, 0);
        if (g > 0.0) hi = fmin(hi, ls); else lo = fmax(lo, ls);
        double nl = d > 0.0 ? ls - g / d : 0.5 * (lo + hi);
        if (!(nl > lo && nl < hi)) nl = 0.5 * (lo + hi);
        // an error e in ln s costs about d*e^2/2 nats of head-room (of ~700): stop once that is negligible
        const bool done = fabs(nl - ls) * sqrt(fmax(d, 1.0)) < 0.5;
        ls = nl;
        if (done) break;
    }
    return ls;
}

struct TailOut {
    double lnT;        // ln P(X >= K)
    double lnKm1;      // ln P(X == K-1)
    int flags;
};

// ln P(X >= K) for K > KS, choosing the tilt; leaves the final row in `row`
template <int R, class Src>
__device__ TailOut heavy_tail(const Src &src, int K, int N, double lam, double2 *sm, Row<R> &row)
{
    TailOut out;
    out.flags = 0;
    // Chernoff exponent of the tail: beyond ~300 nats the untilted cells of interest drift out of fp64 range
    const double cher = ((double)K > lam) ? ((double)K * log((double)K / lam) - (double)K + lam) : 0.0;
    double ln_s = 0.0;
    if (cher > 300.0) ln_s = newton_tilt(src, K, N, lam);
    for (int attempt = 0; attempt < 2; ++attempt) {
        dp_run<R>(src, K, ln_s, sm, row);
        // how far below the largest cell does the absorbing state sit?
        int hi = 0;
#pragma unroll
        for (int r = 0; r < R; ++r) hi = max(hi, __double2hiint(row.E[r]));
        hi = __reduce_max_sync(FULL, hi);
        const int hiT = __shfl_sync(FULL, __double2hiint(row.T), 31);
        const int gap = ((max(hi, hiT)) >> 20) - (hiT >> 20);
        if (gap > 580 && ln_s == 0.0 && attempt == 0) {
            ln_s = newton_tilt(src, K, N, lam);
            continue;
        }
        if (gap > 900) out.flags |= CF_RANGE;
        break;
    }
    const double T = __shfl_sync(FULL, row.T, 31);
    const double top = __shfl_sync(FULL, row.E[R - 1], 31);
    const double base = (double)row.e2 * LN2 + row.sum_lq;
    out.lnT = log(T) + base - (double)K * row.ln_s;
    out.lnKm1 = log(top) + base - (double)(K - 1) * row.ln_s;
    return out;
}

// sum over cells k >= c of an untilted final row plus the absorbing state -> ln P(X >= c)
template <int R>
__device__ double row_tail(const Row<R> &row, int K, int c)
{
    const int lane = lane_id();
    const int k0 = K - 32 * R + lane * R;
    double s = (lane == 31) ? row.T : 0.0;
#pragma unroll
    for (int r = R - 1; r >= 0; --r)
        if (k0 + r >= c) s += row.E[r];
    s = warp_sum(s);
    s = __shfl_sync(FULL, s, 0);
    return log(s) + (double)row.e2 * LN2 + row.sum_lq;
}

// The whole snpcaller() arithmetic for one column / problem with K = max count.
// Returns false when the column is clearly insignificant (no site).
template <int R, class Src>
__device__ bool run_problem(const Src &src, const int (&cnt)[3], long long bonf, double sig, double2 *sm, Cand &cd)
{
    const int lane = lane_id();
    const int K = max(cnt[0], max(cnt[1], cnt[2]));
    cd.flags = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i) cd.lnp[i] = 0.0;
    cd.ln_floor = 0.0;

    if (K <= KS) {
        double P[KS], T;
        src_small_dispatch(src, K, P, T);
        if (T * (double)bonf > sig * (1.0 + 1e-9)) return false;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            if (cnt[i] == 0) continue;
            double s = T;
#pragma unroll
            for (int k = KS - 1; k >= 0; --k)
                if (k >= cnt[i] && k < K) s += P[k];
            cd.lnp[i] = log(s);
        }
        double pk1 = P[0];
#pragma unroll
        for (int k = 1; k < KS; ++k)
            if (k == K - 1) pk1 = P[k];
        cd.ln_floor = log(fmin(pk1, T));
        return true;
    }

    // N = reads that survive the filters, lam = sum of their error probabilities
    int N = 0;
    double lam = 0.0;
    for (int t0 = 0; t0 < src.size(); t0 += Src::TILE) {
        src.stage(t0);
        const int t1 = min(src.size() - t0, Src::TILE) + t0;
        for (int pos = t0 + lane; pos < t1; pos += 32) {
            double jp;
            if (!src.get(pos, jp)) continue;
            double p, q;
            guard_pq(jp, p, q);
            lam += p;
            ++N;
        }
    }
    N = __reduce_add_sync(FULL, N);
    lam = __shfl_sync(FULL, warp_sum(lam), 0);

    Row<R> row;
    const TailOut main_t = heavy_tail<R>(src, K, N, lam, sm, row);
    cd.flags |= main_t.flags;
    if (main_t.lnT > -700.0 && exp(main_t.lnT) * (double)bonf > sig * (1.0 + 1e-9)) return false;
    cd.ln_floor = fmin(main_t.lnT, main_t.lnKm1);
    const bool untilted = (row.ln_s == 0.0);
    double sec[3] = {0.0, 0.0, 0.0};
    // alleles below K that can reuse the untilted row
    if (untilted) {
#pragma unroll
        for (int i = 0; i < 3; ++i)
            if (cnt[i] > 0 && cnt[i] < K) sec[i] = row_tail<R>(row, K, cnt[i]);
    }
#pragma unroll 1
    for (int i = 0; i < 3; ++i) {
        const int c = cnt[i];
        if (c == 0) continue;
        if (c == K) { cd.lnp[i] = main_t.lnT; continue; }
        if (untilted) { cd.lnp[i] = sec[i]; continue; }
        if (i == 2 && c == cnt[1]) { cd.lnp[i] = cd.lnp[1]; continue; }
        if (i >= 1 && c == cnt[0]) { cd.lnp[i] = cd.lnp[0]; continue; }
        if (c <= KS) {
            double P[KS], T;
            src_small_dispatch(src, c, P, T);
            cd.lnp[i] = log(T);
        } else {
            const TailOut t2 = heavy_tail<R>(src, c, N, lam, sm, row);
            cd.flags |= t2.flags;
            cd.lnp[i] = t2.lnT;
        }
    }
    return true;
}

// ------------------------------------------------------------------------------------------------
// k_mid: the columns with K <= 8 that the lane-per-column prune of k_front / k_prune2 could not rule out within
// PRUNE_CAP reads (true low-frequency variants, the first columns of a run).  At this size the recurrence over reads
// (depth serial steps of a short row) is slower than folding the reads in parallel — every lane its 16-byte chunks into
// a distribution truncated at 2, 4 or 8 — and merging the 32 distributions by truncated convolution.
// Untilted: if the tail leaves the fp64 range the column is handed to the per-column fallback (k_heavy_all).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_mid(const __grid_constant__ DevConf cf, const __grid_constant__ DevBatch b, const Lut *lut,
                                             const Workspace ws)
{
    __shared__ double s_lut[768];
    __shared__ int s_hist[4][256];
    if (ws.counters->n_jobs[0] == 0) return;                 // nothing listed: leave before the table is even loaded
    load_lut(s_lut, lut);
    unsigned lut_sa = (unsigned)__cvta_generic_to_shared(s_lut);
    const int lane = lane_id(), wib = threadIdx.x >> 5;
    const unsigned njobs = ws.counters->n_jobs[0];
    const int *jobs = ws.jobs;
    for (;;) {
        unsigned j = 0;
        if (lane == 0) j = atomicAdd(&ws.counters->next_job[0], 1u);
        j = __shfl_sync(FULL, j, 0);
        if (j >= njobs) break;
        const long long c = jobs[j];
        Geom g;
        int cov;
        load_geom(b, c, g, cov);
        setup_alt_bq(cf, b, s_lut, g, s_hist[wib]);
        int cnt[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) cnt[i] = ws.cnt6[6 * c + i];
        const int K = max(cnt[0], max(cnt[1], cnt[2]));
        const long long bonf = bonf_of(cf, ws.counters->bonf_start_used, col_rank(ws, c));
        double tails[4];
        if (K <= 2) screen_small<2>(cf, b, lut_sa, g, cnt, K, tails);
        else if (K <= 4) screen_small<4>(cf, b, lut_sa, g, cnt, K, tails);
        else screen_small<8>(cf, b, lut_sa, g, cnt, K, tails);
        double tK = cnt[0] == K ? tails[0] : cnt[1] == K ? tails[1] : tails[2];
        tK = __shfl_sync(FULL, tK, 0);
        const double fl = __shfl_sync(FULL, tails[3], 0);
        if (!(fl > 1e-280)) {
            // too far out for the untilted form: k_heavy<1> has the tilt
            if (lane == 0) ws.jobs[(long long)CLS_FALLBACK * ws.cap_cols + atomicAdd(&ws.counters->n_jobs[CLS_FALLBACK], 1u)] = (int)c;
            continue;
        }
        if (tK * (double)bonf > cf.sig * (1.0 + 1e-9)) continue;
        if (lane == 0) {
            Cand cd;
            cd.col = c;
            cd.bonf = bonf;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                cd.lnp[i] = cnt[i] > 0 ? log(tails[i]) : 0.0;
                cd.cnt[i] = cnt[i];
                cd.raw[i] = ws.cnt6[6 * c + 3 + i];
            }
            cd.ln_floor = log(tails[3]);
            cd.flags = 0;
            cd.pad = 0;
            ws.cand[atomicAdd(&ws.counters->n_cand, 1u)] = cd;
            mark_cand(ws, c);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// k_heavy_all: the per-column fallback — one warp per column, K cells tiled over 32 lanes x R registers, every
// repair the general routine knows (run_problem: tilt after the fact, rescaling after every read, separate runs for
// the other alleles).  Takes what k_dp and k_mid hand back; normally every list is empty and the kernel leaves at once.
// ------------------------------------------------------------------------------------------------
template <int R>
__device__ void heavy_list(const DevConf &cf, const DevBatch &b, const double *s_lut, const Workspace &ws, int cls, double2 *s_par,
                           int *s_hist, double *s_stage)
{
    const int lane = lane_id();
    const unsigned njobs = ws.counters->n_jobs[cls];
    const int *jobs = ws.jobs + (long long)cls * ws.cap_cols;
    for (;;) {
        unsigned j = 0;
        if (lane == 0) j = atomicAdd(&ws.counters->next_job[cls], 1u);
        j = __shfl_sync(FULL, j, 0);
        if (j >= njobs) break;
        const long long c = jobs[j];
        Staged<ByteSrc> src;
        src.init(s_stage);
        src.raw.cf = &cf;
        src.raw.b = &b;
        src.raw.lut = s_lut;
        int cov;
        load_geom(b, c, src.raw.g, cov);
        setup_alt_bq(cf, b, s_lut, src.raw.g, s_hist);
        int cnt[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) cnt[i] = ws.cnt6[6 * c + i];
        const long long bonf = bonf_of(cf, ws.counters->bonf_start_used, col_rank(ws, c));
        Cand cd;
        const bool site = run_problem<R>(src, cnt, bonf, cf.sig, s_par, cd);
        if (site && lane == 0) {
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

__global__ void __launch_bounds__(128) k_heavy_all(const __grid_constant__ DevConf cf, const __grid_constant__ DevBatch b, const Lut *lut,
                                                   const Workspace ws)
{
    __shared__ double s_lut[768];
    __shared__ double2 s_par[4][32];
    __shared__ int s_hist[4][256];
    extern __shared__ double s_stage[];          // [4][STAGE_DOUBLES]
    const Counters *cn = ws.counters;
    if ((cn->n_jobs[3] | cn->n_jobs[4] | cn->n_jobs[5] | cn->n_jobs[6] | cn->n_jobs[CLS_FALLBACK]) == 0) return;
    load_lut(s_lut, lut);
    const int wib = threadIdx.x >> 5;
    double *stg = s_stage + wib * STAGE_DOUBLES;
    heavy_list<64>(cf, b, s_lut, ws, 6, s_par[wib], s_hist[wib], stg);
    heavy_list<32>(cf, b, s_lut, ws, 5, s_par[wib], s_hist[wib], stg);
    heavy_list<16>(cf, b, s_lut, ws, 4, s_par[wib], s_hist[wib], stg);
    heavy_list<8>(cf, b, s_lut, ws, 3, s_par[wib], s_hist[wib], stg);
    heavy_list<8>(cf, b, s_lut, ws, CLS_FALLBACK, s_par[wib], s_hist[wib], stg);
}

// ------------------------------------------------------------------------------------------------
// k_heavy_xl: one CTA per column for 2048 < K <= 16384 (deep amplicons: depth >= ~4100 at AF 50 %).
// The same odds-form recurrence as dp_run, K cells over 256 threads x 64 registers; the cell that crosses
// a warp boundary goes through shared memory (double-buffered, one barrier per read).
// ------------------------------------------------------------------------------------------------
constexpr int XL_T = 256, XL_R = 64, XL_W = XL_T / 32;

struct XlShared {
    double2 par[XL_T];          // (o, 1/q) of the reads of the current stripe, compacted per warp
    int par_n[XL_W];
    double edge[2][XL_W];       // top cell of every warp after the previous read
    double red[XL_W];
    int redi[XL_W];
};

__device__ __forceinline__ double block_sum(double v, XlShared &sh)
{
    v = warp_sum(v);
    __syncthreads();
    if (lane_id() == 0) sh.red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < XL_W; ++w) t += sh.red[w];
    return t;
}

__device__ __forceinline__ int block_max_int(int v, XlShared &sh)
{
    v = __reduce_max_sync(FULL, v);
    __syncthreads();
    if (lane_id() == 0) sh.redi[threadIdx.x >> 5] = v;
    __syncthreads();
    int t = sh.redi[0];
#pragma unroll
    for (int w = 1; w < XL_W; ++w) t = max(t, sh.redi[w]);
    return t;
}

struct XlRow {
    double E[XL_R];
    double T;          // valid on the last thread
    int e2;
    double sum_lq, ln_s;
};

__device__ void xl_rescale(XlRow &row, XlShared &sh, bool force)
{
    int hi = 0;
#pragma unroll
    for (int r = 0; r < XL_R; ++r) hi = max(hi, __double2hiint(row.E[r]));
    if (threadIdx.x == XL_T - 1) hi = max(hi, __double2hiint(row.T));
    hi = block_max_int(hi, sh);
    const int ex = (hi >> 20) - 1023;
    if (force ? (ex != 0) : (ex > 200 || ex < -200)) {
        const double f = __hiloint2double((1023 - ex) << 20, 0);
#pragma unroll
        for (int r = 0; r < XL_R; ++r) row.E[r] *= f;
        row.T *= f;
        row.e2 += ex;
    }
}

template <class Src>
__device__ void xl_dp_run(const Src &src, int K, double ln_s, XlShared &sh, XlRow &row)
{
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const double s = (ln_s == 0.0) ? 1.0 : exp(ln_s);
    const int k0 = K - XL_T * XL_R + tid * XL_R;
#pragma unroll
    for (int r = 0; r < XL_R; ++r) row.E[r] = (k0 + r == 0) ? 1.0 : 0.0;
    row.T = 0.0;
    row.e2 = 0;
    row.ln_s = ln_s;
    double lq_acc = 0.0, q_prod = 1.0;
    if (lane == 31) sh.edge[0][w] = 0.0;     // before the first read every boundary cell is 0 (k0 + 63 == 0 cannot be a warp top unless K is tiny)
    if (lane == 31 && k0 + XL_R - 1 == 0) sh.edge[0][w] = 1.0;
    int parity = 0;
    const int n = src.size();
    const unsigned lt_mask = (1u << lane) - 1u;
    for (int n0 = 0; n0 < n; n0 += XL_T) {
        const int pos = n0 + tid;
        double jp = 0.0, o = 0.0, rq = 1.0;
        const bool ok = pos < n && src.get(pos, jp);
        if (ok) {
            double p, q;
            guard_pq(jp, p, q);
            rq = 1.0 / q;
            o = p * s * rq;
            q_prod *= q;
            if (q_prod < 1e-200) {
                lq_acc += log(q_prod);
                q_prod = 1.0;
            }
        }
        const unsigned m = __ballot_sync(FULL, ok);
        const bool slow = __syncthreads_or(ok && (o > 1048576.0 || rq > 1048576.0));
        if (ok) sh.par[w * 32 + __popc(m & lt_mask)] = make_double2(o, rq);
        if (lane == 0) sh.par_n[w] = __popc(m);
        __syncthreads();
        for (int ws = 0; ws < XL_W; ++ws) {
            const int cnt = sh.par_n[ws];
            for (int j = 0; j < cnt; ++j) {
                const double2 c = sh.par[ws * 32 + j];
                const double top = row.E[XL_R - 1];
                double in = __shfl_up_sync(FULL, top, 1);
                if (lane == 0) in = w ? sh.edge[parity][w - 1] : 0.0;
                row.T = fma(top, c.x, row.T * c.y);
#pragma unroll
                for (int r = XL_R - 1; r >= 1; --r) row.E[r] = fma(row.E[r - 1], c.x, row.E[r]);
                row.E[0] = fma(in, c.x, row.E[0]);
                parity ^= 1;
                if (slow) xl_rescale(row, sh, true);          // before the boundary copy: it must carry the new scale
                if (lane == 31) sh.edge[parity][w] = row.E[XL_R - 1];
                __syncthreads();
            }
            if (!slow) {
                // every 32 reads, like the warp kernel: (1 + 2^20)^32 stays inside fp64
                xl_rescale(row, sh, false);
                if (lane == 31) sh.edge[parity][w] = row.E[XL_R - 1];   // same cell, possibly rescaled
                __syncthreads();
            }
        }
    }
    lq_acc += log(q_prod);
    row.sum_lq = block_sum(lq_acc, sh);
}

template <class Src>
__device__ double xl_newton(const Src &src, int K, int N, double lam, XlShared &sh)
{
    const double kt = fmin((double)K, (double)N - 0.5);
    const double s0 = kt * fmax((double)N - lam, 1e-300) / (fmax(lam, 1e-300) * ((double)N - kt));
    double lo = 0.0, hi = 60.0;
    double ls = fmin(log(fmax(s0, 1.0)), hi);
    const int n = src.size();
    for (int it = 0; it < 40; ++it) {
        const double s = exp(ls);
        double g = 0.0, d = 0.0;
        for (int pos = threadIdx.x; pos < n; pos += XL_T) {
            double jp;
            if (!src.get(pos, jp)) continue;
            double p, q;
            guard_pq(jp, p, q);
            const double ps = p * s;
            const double wv = ps / (q + ps);
            g += wv;
            d += wv * (1.0 - wv);
        }
        g = block_sum(g, sh) - kt;
        d = block_sum(d, sh);
        if (g > 0.0) hi = fmin(hi, ls); else lo = fmax(lo, ls);
        double nl = d > 0.0 ? ls - g / d : 0.5 * (lo + hi);
        if (!(nl > lo && nl < hi)) nl = 0.5 * (lo + hi);
        const bool done = fabs(nl - ls) * sqrt(fmax(d, 1.0)) < 0.5;
        ls = nl;
        if (done) break;
    }
    return ls;
}

// ln P(X >= K) and ln P(X = K-1) for one column, whole CTA
template <class Src>
__device__ TailOut xl_tail(const Src &src, int K, int N, double lam, XlShared &sh, XlRow &row)
{
    TailOut out;
    out.flags = 0;
    const double cher = ((double)K > lam) ? ((double)K * log((double)K / lam) - (double)K + lam) : 0.0;
    double ln_s = 0.0;
    if (cher > 300.0) ln_s = xl_newton(src, K, N, lam, sh);
    for (int attempt = 0; attempt < 2; ++attempt) {
        xl_dp_run(src, K, ln_s, sh, row);
        int hi = 0;
#pragma unroll
        for (int r = 0; r < XL_R; ++r) hi = max(hi, __double2hiint(row.E[r]));
        hi = block_max_int(hi, sh);
        const int hiT = block_max_int(threadIdx.x == XL_T - 1 ? __double2hiint(row.T) : 0, sh);
        const int gap = ((max(hi, hiT)) >> 20) - (hiT >> 20);
        if (gap > 580 && ln_s == 0.0 && attempt == 0) {
            ln_s = xl_newton(src, K, N, lam, sh);
            continue;
        }
        if (gap > 900) out.flags |= CF_RANGE;
        break;
    }
    // T and the cell K-1 live on the last thread
    __syncthreads();
    if (threadIdx.x == XL_T - 1) {
        sh.red[0] = row.T;
        sh.red[1] = row.E[XL_R - 1];
    }
    __syncthreads();
    const double T = sh.red[0], top = sh.red[1];
    __syncthreads();
    const double base = (double)row.e2 * LN2 + row.sum_lq;
    out.lnT = log(T) + base - (double)K * row.ln_s;
    out.lnKm1 = log(top) + base - (double)(K - 1) * row.ln_s;
    return out;
}

// the snpcaller() arithmetic for one column / problem, whole CTA; false = clearly insignificant
template <class Src>
__device__ bool xl_problem(const Src &src, const int (&cnt)[3], long long bonf, double sig, XlShared &sh, double *s_small, Cand &cd)
{
    const int K = max(cnt[0], max(cnt[1], cnt[2]));
    int Nl = 0;
    double laml = 0.0;
    for (int pos = threadIdx.x; pos < src.size(); pos += XL_T) {
        double jp;
        if (!src.get(pos, jp)) continue;
        double p, q;
        guard_pq(jp, p, q);
        laml += p;
        ++Nl;
    }
    const int N = (int)(block_sum((double)Nl, sh) + 0.5);
    const double lam = block_sum(laml, sh);
    XlRow row;
    const TailOut main_t = xl_tail(src, K, N, lam, sh, row);
    cd.flags = main_t.flags;
    cd.lnp[0] = cd.lnp[1] = cd.lnp[2] = 0.0;
    cd.ln_floor = 0.0;
    if (main_t.lnT > -700.0 && exp(main_t.lnT) * (double)bonf > sig * (1.0 + 1e-9)) return false;
    cd.ln_floor = fmin(main_t.lnT, main_t.lnKm1);
#pragma unroll 1
    for (int i = 0; i < 3; ++i) {
        const int cc = cnt[i];
        if (cc == 0) continue;
        if (cc == K) { cd.lnp[i] = main_t.lnT; continue; }
        if (i == 2 && cc == cnt[1]) { cd.lnp[i] = cd.lnp[1]; continue; }
        if (i >= 1 && cc == cnt[0]) { cd.lnp[i] = cd.lnp[0]; continue; }
        if (cc <= KS) {
            if (threadIdx.x < 32) {
                double P[KS], T;
                src_small_dispatch(src, cc, P, T);
                if (threadIdx.x == 0) s_small[0] = log(T);
            }
            __syncthreads();
            cd.lnp[i] = s_small[0];
            __syncthreads();
        } else {
            const TailOut t2 = xl_tail(src, cc, N, lam, sh, row);
            cd.flags |= t2.flags;
            cd.lnp[i] = t2.lnT;
        }
    }
    return true;
}

__global__ void __launch_bounds__(XL_T, 1) k_heavy_xl(const __grid_constant__ DevConf cf, const __grid_constant__ DevBatch b,
                                                      const Lut *lut, const Workspace ws, int cls)
{
    __shared__ double s_lut[768];
    __shared__ XlShared sh;
    __shared__ int s_hist[256];
    __shared__ unsigned s_job;
    __shared__ double s_small[KS + 1];
    if (ws.counters->n_jobs[cls] == 0) return;                 // nothing listed: leave before the table is even loaded
    load_lut(s_lut, lut);
    const unsigned njobs = ws.counters->n_jobs[cls];
    const int *jobs = ws.jobs + (long long)cls * ws.cap_cols;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_job = atomicAdd(&ws.counters->next_job[cls], 1u);
        __syncthreads();
        const unsigned j = s_job;
        if (j >= njobs) break;
        const long long c = jobs[j];
        ByteSrc src;
        src.cf = &cf;
        src.b = &b;
        src.lut = s_lut;
        int cov;
        load_geom(b, c, src.g, cov);
        if (cf.alt_bq_mode) {
            // every warp computes the same override (the median needs a histogram: warp 0 does it, others read it)
            if (threadIdx.x < 32) setup_alt_bq(cf, b, s_lut, src.g, s_hist);
            if (threadIdx.x == 0) sh.red[0] = src.g.alt_bp;
            __syncthreads();
            src.g.alt_bp = sh.red[0];
            __syncthreads();
        }
        int cnt[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) cnt[i] = ws.cnt6[6 * c + i];
        const int K = max(cnt[0], max(cnt[1], cnt[2]));
        const long long bonf = bonf_of(cf, ws.counters->bonf_start_used, col_rank(ws, c));
        Cand cd;
        if (K > XL_T * XL_R) {
            // no kernel of this build takes an alt count this large: the column is reported as a site whose alleles carry
            // the status "unsupported" (never silently skipped), every other column of the batch is unaffected
            cd.flags = CF_UNSUPPORTED;
            cd.lnp[0] = cd.lnp[1] = cd.lnp[2] = 0.0;
            cd.ln_floor = 0.0;
        } else if (!xl_problem(src, cnt, bonf, cf.sig, sh, s_small, cd)) continue;
        if (threadIdx.x == 0) {
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

// stand-alone problems with 2048 < K <= 16384, one CTA each
__global__ void __launch_bounds__(XL_T, 1) k_prob_jobs_xl(const ProbBatch pb, Cand *out)
{
    __shared__ XlShared sh;
    __shared__ double s_small[KS + 1];
    for (long long i = blockIdx.x; i < pb.n; i += gridDim.x) {
        int cnt[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) cnt[a] = pb.counts[3 * i + a];
        const int K = max(cnt[0], max(cnt[1], cnt[2]));
        if (K <= MAXK_WARP) continue;
        ProbSrc src;
        src.ep = pb.err_probs + pb.ep_off[i];
        src.n = (int)(pb.ep_off[i + 1] - pb.ep_off[i]);
        Cand cd;
        cd.flags = 0;
        cd.ln_floor = 0.0;
        cd.lnp[0] = cd.lnp[1] = cd.lnp[2] = 0.0;
        bool site = false;
        const bool ok = K <= XL_T * XL_R && K <= src.n;
        __syncthreads();
        if (ok) site = xl_problem(src, cnt, pb.bonf[i], pb.sig, sh, s_small, cd);
        if (threadIdx.x == 0) {
            if (!site) cd.flags |= CF_INSIG;
            if (!ok) cd.flags |= CF_UNSUPPORTED;
            cd.col = i;
            cd.bonf = pb.bonf[i];
#pragma unroll
            for (int a = 0; a < 3; ++a) { cd.cnt[a] = cnt[a]; cd.raw[a] = cnt[a]; }
            cd.pad = 0;
            out[i] = cd;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// k_prob_jobs<R>: stand-alone snpcaller() problems on ready-made error probabilities
// ------------------------------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(128) k_prob_jobs(const ProbBatch pb, Cand *out, int cls)
{
    __shared__ double2 s_par[4][32];
    extern __shared__ double s_stage[];          // [4][STAGE_DOUBLES]
    const int lane = lane_id(), wib = threadIdx.x >> 5;
    const long long warp0 = (long long)blockIdx.x * 4 + wib, nwarps = (long long)gridDim.x * 4;
    for (long long i = warp0; i < pb.n; i += nwarps) {
        int cnt[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) cnt[a] = pb.counts[3 * i + a];
        const int K = max(cnt[0], max(cnt[1], cnt[2]));
        const int my_cls = K <= KS ? 0 : class_of(K);
        if (my_cls != cls) continue;
        Staged<ProbSrc> src;
        src.init(s_stage + wib * STAGE_DOUBLES);
        src.raw.ep = pb.err_probs + pb.ep_off[i];
        src.raw.n = (int)(pb.ep_off[i + 1] - pb.ep_off[i]);
        Cand cd;
        bool site = false;
        if (K > 0 && K <= src.raw.n && cls < CLS_XL) site = run_problem<R>(src, cnt, pb.bonf[i], pb.sig, s_par[wib], cd);
        else { cd.flags = 0; cd.ln_floor = 0.0; cd.lnp[0] = cd.lnp[1] = cd.lnp[2] = 0.0; }
        if (lane == 0) {
            if (!site) cd.flags |= CF_INSIG;
            if (K > MAXK_WARP || K > src.raw.n) cd.flags |= CF_UNSUPPORTED;
            cd.col = i;
            cd.bonf = pb.bonf[i];
#pragma unroll
            for (int a = 0; a < 3; ++a) { cd.cnt[a] = cnt[a]; cd.raw[a] = cnt[a]; }
            cd.pad = 0;
            out[i] = cd;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
// dynamic shared memory of the kernels that stage a tile of merged probabilities per warp (4 warps per CTA);
// together with their static arrays they pass the 48 KB default limit, hence the opt-in
constexpr int STAGE_BYTES = 4 * STAGE_DOUBLES * (int)sizeof(double);

template <int R>
static void stage_optin_one()
{
    cudaFuncSetAttribute(k_prob_jobs<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, STAGE_BYTES);
}

// Everything a launch needs besides its arguments lives in the context: created once per context, after
// cudaSetDevice(device) — function attributes are per device, streams and events belong to the device current at
// creation, and two contexts (two host threads, two GPUs) must never fork and join through the same events.
int launch_state_init(LaunchState &ls, int device)
{
    if (ls.ready) return 0;
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || n <= 0) n = 148;
    ls.sms = n;
    stage_optin_one<1>(); stage_optin_one<2>(); stage_optin_one<4>(); stage_optin_one<8>();
    stage_optin_one<16>(); stage_optin_one<32>(); stage_optin_one<64>();
    cudaFuncSetAttribute(k_heavy_all, cudaFuncAttributeMaxDynamicSharedMemorySize, STAGE_BYTES);
    if (dp_smem_optin() || xl_smem_optin()) return 1;
    if (cudaEventCreateWithFlags(&ls.ev_fork, cudaEventDisableTiming) != cudaSuccess) return 1;
    if (cudaEventCreateWithFlags(&ls.ev_fork2, cudaEventDisableTiming) != cudaSuccess) return 1;
    for (int i = 0; i < NSIDE; ++i) {
        if (cudaStreamCreateWithFlags(&ls.side[i], cudaStreamNonBlocking) != cudaSuccess) return 1;
        if (cudaEventCreateWithFlags(&ls.ev_join[i], cudaEventDisableTiming) != cudaSuccess) return 1;
    }
    ls.ready = true;
    return 0;
}

void launch_state_destroy(LaunchState &ls)
{
    if (ls.ev_fork) cudaEventDestroy(ls.ev_fork);
    if (ls.ev_fork2) cudaEventDestroy(ls.ev_fork2);
    for (int i = 0; i < NSIDE; ++i) {
        if (ls.side[i]) cudaStreamDestroy(ls.side[i]);
        if (ls.ev_join[i]) cudaEventDestroy(ls.ev_join[i]);
    }
    ls = LaunchState();
}

__global__ void k_rank_cands(const Workspace ws);

// the exact factor a shard's batch continues from (the shards before it have added to what k_front saw): every kernel of
// the test phase reads it from the counters
__global__ void k_set_start(const Workspace ws, const long long *bonf_start_dev) { ws.counters->bonf_start_used = *bonf_start_dev; }

void launch_test(const LaunchState &ls, const DevConf &cf, const DevBatch &b, const Lut *lut, const Workspace &ws, cudaStream_t st,
                 cudaEvent_t after_finalize, cudaEvent_t after_heavy, const long long *bonf_start_dev)
{
    if (b.n_cols <= 0) return;
    const int nb = (int)((b.n_cols + FIN_BLOCK - 1) / FIN_BLOCK);
    // Independent work side by side, so that the warps of the small kernels share the SMs with k_dp<0>: the second stage of
    // the prune with the exact factor (k_front, in the screen phase, did the first) and k_mid for what survives it (K <= 8);
    // k_dp<1>, k_dp<2> (256 < K <= 2048, their own register budgets), k_xl (K > 2048, one CTA per column); an empty list
    // costs an early exit.  Then the per-column fallbacks for the few columns they hand back.  k_dp<0>, the longest, goes
    // first on the launching stream: nothing it needs comes from the prune.
    // A kernel whose list was empty in the context's previous batch is simply queued behind k_dp<0> (side_mask): on data
    // without deep columns that saves the fork/join events — host time per batch is what limits 8 GPUs on one box.
    const unsigned m = ls.side_mask;
    cudaStream_t s_mid = (m & 1u) ? ls.side[0] : st, s_dp1 = (m & 2u) ? ls.side[1] : st, s_xl = (m & 4u) ? ls.side[2] : st,
                 s_dp2 = (m & 8u) ? ls.side[3] : st;
    const bool timed = after_finalize != nullptr;       // per-phase timing (lfb200_set_profiling): the prune alone, first
    if (bonf_start_dev) k_set_start<<<1, 1, 0, st>>>(ws, bonf_start_dev);
    if (timed) {
        k_prune2<<<ls.sms * 2, 128, 0, st>>>(cf, b, lut, ws, bonf_start_dev);
        cudaEventRecord(after_finalize, st);
    }
    if (m) {
        cudaEventRecord(ls.ev_fork, st);
        for (int i = 0; i < 4; ++i)
            if (m & (1u << i)) cudaStreamWaitEvent(ls.side[i], ls.ev_fork, 0);
    }
    if (m & 1u) {
        if (!timed) k_prune2<<<ls.sms * 2, 128, 0, s_mid>>>(cf, b, lut, ws, bonf_start_dev);
        k_mid<<<ls.sms * 4, 128, 0, s_mid>>>(cf, b, lut, ws);
    }
    if (m & 4u) launch_xl(ls, cf, b, lut, ws, s_xl);
    launch_dp(ls, cf, b, lut, ws, st, s_dp1, s_dp2);
    if (!(m & 1u)) {
        if (!timed) k_prune2<<<ls.sms * 2, 128, 0, st>>>(cf, b, lut, ws, bonf_start_dev);
        k_mid<<<ls.sms * 4, 128, 0, st>>>(cf, b, lut, ws);
    }
    if (!(m & 4u)) launch_xl(ls, cf, b, lut, ws, st);
    for (int i = 0; i < 4; ++i)
        if (m & (1u << i)) {
            cudaEventRecord(ls.ev_join[i], ls.side[i]);
            cudaStreamWaitEvent(st, ls.ev_join[i], 0);
        }
    k_heavy_xl<<<ls.sms, XL_T, 0, st>>>(cf, b, lut, ws, CLS_XLFB);
    k_heavy_all<<<ls.sms, 128, STAGE_BYTES, st>>>(cf, b, lut, ws);
    if (after_heavy) cudaEventRecord(after_heavy, st);
    // the sites in column order (see k_rank_cands)
    k_scan_blocks<<<1, 1024, 0, st>>>(ws.candtile, ws.candpre, nb, nullptr);
    k_rank_cands<<<ls.sms, 256, 0, st>>>(ws);
}

// Column order of the sites on the device: every kernel that emits a candidate marks its column (mark_cand), a prefix
// sum over the marks per tile of 256 columns gives each candidate its rank, and perm[rank] = index into ws.cand — the
// host finishes the sites straight into their final places instead of sorting them (a column yields one candidate).
__global__ void __launch_bounds__(256) k_rank_cands(const Workspace ws)
{
    const unsigned n_cand = ws.counters->n_cand;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n_cand; i += gridDim.x * blockDim.x) {
        const long long c = ws.cand[i].col;
        const long long t0 = c & ~255ll;
        int before = 0;
        for (long long a = t0; a < c; a += 16) {              // marks of the tile before column c, 16 at a time
            uint4 v = *reinterpret_cast<const uint4 *>(ws.is_cand + a);
            const int keep = (int)(c - a);                     // bytes of this chunk that lie before c
            if (keep < 16) {
                unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int kb = keep - 4 * q;
                    if (kb <= 0) w[q] = 0;
                    else if (kb < 4) w[q] &= (1u << (8 * kb)) - 1u;
                }
                v = make_uint4(w[0], w[1], w[2], w[3]);
            }
            before += __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w);
        }
        ws.cand_perm[ws.candpre[c >> 8] + before] = (int)i;
    }
}

// ------------------------------------------------------------------------------------------------
// k_emit_sites: the decision of lofreq_call.c:832 / snpcaller.c:1144-1196 and PROB_TO_PHREDQUAL (utils.h:45) on ln p.
// The reference decides on long double p-values, p = expl(ln p):  clamp to LDBL_MIN / LDBL_MAX when an exp()
// underflowed, `p * bonf > sig` -> nothing called, per allele `p * bonf < sig` -> called with QUAL = (int)(-10 log10l(p)).
// Every one of these is a comparison of ln p (or of a difference of two ln p) with a constant, so it is made here in
// double with a guard band: a site with any comparison inside its band is flagged SITE_NEEDS_HOST and the host repeats
// exactly the reference's long double sequence for it (finish_site, host_api.cpp) — a handful per million columns.
// Records go out in column order (cand_perm) as byte images of lfb200_site_t, straight into mapped pinned memory.
// ------------------------------------------------------------------------------------------------
constexpr double LN_LDBL_MIN = -11355.137111933024;        // expl() underflows (FE_UNDERFLOW) below this
constexpr double LN_DBL_EPS = -36.04365338911715;          // ln(DBL_EPSILON)
constexpr double LN_EXP_UNDER = 708.3964185322641;         // glibc exp() raises FE_UNDERFLOW for arguments below -this

__device__ __forceinline__ bool near(double x, double band) { return fabs(x) < band; }

__global__ void __launch_bounds__(128) k_emit_sites(const __grid_constant__ DevConf cf, const Workspace ws, SiteRec *out, unsigned cap)
{
    Counters *hdr = ws.counters;
    const unsigned n_cand = hdr->n_cand;
    if (blockIdx.x == 0 && threadIdx.x == 0) hdr->emit_overflow = n_cand > cap;
    const unsigned n = min(n_cand, cap);
    // a CTA takes 128 consecutive records: every thread decides one into shared memory, then the CTA writes the
    // 128 x 144 bytes out as consecutive 16-byte words (full lines over PCIe instead of nine scattered stores per site)
    __shared__ SiteRec s_rec[128];
    for (unsigned r0 = blockIdx.x * 128u; r0 < n; r0 += gridDim.x * 128u) {
      const unsigned r = r0 + threadIdx.x;
      if (r < n) {
        const Cand cd = ws.cand[ws.cand_perm[r]];
        SiteRec &s = s_rec[threadIdx.x];
        s.col = cd.col;
        s.bonf = cd.bonf;
        int K = 0, imax = 0;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            s.lnp[i] = cd.lnp[i];
            s.cnt[i] = cd.cnt[i];
            s.raw[i] = cd.raw[i];
            s.qual[i] = -1;
            s.status[i] = 1;                     // LFB200_ST_LDBLMAX
            s.called[i] = 0;
            s.pvalue_bits[2 * i] = s.pvalue_bits[2 * i + 1] = 0ull;
            if (cd.cnt[i] > K) { K = cd.cnt[i]; imax = i; }
        }
        s.flags = 0;
        s.pad = 0;
        s.ln_floor = cd.ln_floor;
        bool host = false;
        if (cd.flags & CF_UNSUPPORTED) {
            s.flags |= SITE_UNSUPPORTED;
#pragma unroll
            for (int i = 0; i < 3; ++i) if (cd.cnt[i] > 0) s.status[i] = 3;      // LFB200_ST_UNSUPPORTED
            atomicAdd(&hdr->n_unsupported, 1u);
        } else if (K > 0 && !(cd.flags & CF_INSIG)) {
            if (cd.flags & CF_RANGE) atomicOr(&ws.counters->err_flags, (unsigned)CF_RANGE);
            const double lb = log((double)cd.bonf);
            const double tK = cd.lnp[imax];
            // poissbin: pvalue = expl(row[K]), clamped to LDBL_MIN when expl underflows (snpcaller.c:1047-1059)
            host |= near(tK - LN_LDBL_MIN, 1e-6);
            const bool underK = tK < LN_LDBL_MIN;
            bool go = true;
            if (!underK) {                        // snpcaller.c:1155: pvalue * bonf > sig -> every allele stays LDBL_MAX
                const double d = tK + lb - cf.ln_sig;
                host |= near(d, 1e-9);
                go = !(d > 0.0);
            }
            if (go) {
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const int c = cd.cnt[i];
                    if (c == 0) continue;
                    const double t = cd.lnp[i];
                    bool flagged;
                    if (c == K) {
                        flagged = underK;         // the same expl() as poissbin's
                    } else {
                        // probvec_tailsum's exp() underflows when the running sum and the next term are more than
                        // 708.396 nats apart; the widest gap of the log-concave row is against ln_floor
                        const double g = t - cd.ln_floor - LN_EXP_UNDER;
                        host |= near(g, 1e-6) || near(t - LN_LDBL_MIN, 1e-6);
                        flagged = g > 0.0 || t < LN_LDBL_MIN;
                    }
                    int st = 0;                   // LFB200_ST_VALUE
                    if (flagged) {                // snpcaller.c:1174-1188: p < DBL_EPSILON ? LDBL_MIN : LDBL_MAX
                        host |= near(t - LN_DBL_EPS, 1e-9);
                        st = t < LN_DBL_EPS ? 2 : 1;
                    }
                    s.status[i] = (unsigned char)st;
                    if (st == 2) {                // LDBL_MIN * bonf < sig for every representable bonf
                        s.called[i] = 1;
                        s.qual[i] = cf.qual_ldblmin;
                    } else if (st == 0) {
                        const double d = t + lb - cf.ln_sig;              // lofreq_call.c:832
                        host |= near(d, 1e-9);
                        if (d < 0.0) {
                            s.called[i] = 1;
                            // PROB_TO_PHREDQUAL truncates -10 log10l(p): decided unless that lies within 1e-7 of an integer
                            const double q = t * -4.3429448190325182765;
                            const double f = q - floor(q);
                            host |= !(f > 1e-7 && f < 1.0 - 1e-7) || !(t < 0.0 && t > -11000.0);
                            s.qual[i] = (int)q;
                        }
                    }
                }
            }
        }
        if (host) {
            s.flags |= SITE_NEEDS_HOST;
            const unsigned k = atomicAdd(&hdr->n_fix, 1u);
            if (k < (unsigned)EMIT_FIX_MAX) hdr->fix[k] = r;
        }
      }
      __syncthreads();
      const unsigned nrec = min(128u, n - r0);
      const uint4 *src = reinterpret_cast<const uint4 *>(s_rec);
      uint4 *dst = reinterpret_cast<uint4 *>(out + r0);
      // words 3..5 of a record are the long double p-values: the host's business, never written from here
      for (unsigned w = threadIdx.x; w < nrec * 6u; w += 128u) {
          const unsigned rec = w / 6u, k = w - rec * 6u;
          const unsigned at = rec * 9u + (k < 3u ? k : k + 3u);
          dst[at] = src[at];
      }
      __syncthreads();
    }
}

void launch_emit_sites(const LaunchState &ls, const DevConf &cf, const Workspace &ws, SiteRec *out, unsigned cap, cudaStream_t st)
{
    k_emit_sites<<<ls.sms, 128, 0, st>>>(cf, ws, out, cap);
}

__global__ void k_bonf_start(const long long *counts, int rank, long long bonf_subst, long long *start)
{
    long long before = 0;
    for (int r = 0; r < rank; ++r) before += counts[r];
    // lofreq_call.c:794-800: after j tested columns the factor is 3j when it started at 1, else start + 3j
    *start = before > 0 ? (bonf_subst == 1 ? 0 : bonf_subst) + 3 * before : bonf_subst;
}

void launch_bonf_start(const long long *counts, int rank, long long bonf_subst, long long *start, cudaStream_t st)
{
    k_bonf_start<<<1, 1, 0, st>>>(counts, rank, bonf_subst, start);
}

__global__ void k_bonf_start_strided(const long long *counts, int stride, int rank, long long bonf_subst, long long *start)
{
    long long before = 0;
    for (int r = 0; r < rank; ++r) before += counts[(long long)r * stride];
    *start = before > 0 ? (bonf_subst == 1 ? 0 : bonf_subst) + 3 * before : bonf_subst;
}

void launch_bonf_start_strided(const long long *counts, int stride, int rank, long long bonf_subst, long long *start, cudaStream_t st)
{
    k_bonf_start_strided<<<1, 1, 0, st>>>(counts, stride, rank, bonf_subst, start);
}

__global__ void k_set_i64(long long *dst, long long v) { *dst = v; }

void launch_set_i64(long long *dst, long long v, cudaStream_t st) { k_set_i64<<<1, 1, 0, st>>>(dst, v); }

// DFMA throughput probe (the fp64-pipe roofline denominator; MEASURED_PEAKS.json has none for fp64)
__global__ void __launch_bounds__(256) k_dfma_probe(double *out, int iters, double a, double b)
{
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = 1.0 + threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = fma(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    if (s == 123.456) out[0] = s;
}

double measure_dfma_per_second(int sms, cudaStream_t st)
{
    double *out = nullptr;
    if (cudaMalloc(&out, 8) != cudaSuccess) return 0.0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int grid = sms * 8, iters = 20000;
    k_dfma_probe<<<grid, 256, 0, st>>>(out, 1000, 0.999999, 1e-7);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0, st);
        k_dfma_probe<<<grid, 256, 0, st>>>(out, iters, 0.999999, 1e-7);
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    return (double)grid * 256 * 8 * iters / (best * 1e-3);
}

void launch_prob_jobs(int sms, const ProbBatch &pb, Cand *out, cudaStream_t st)
{
    if (pb.n <= 0) return;
    const long long want = (pb.n + 3) / 4;
    const int g = (int)(want < (long long)sms * 4 ? want : (long long)sms * 4);
    k_prob_jobs<1><<<g, 128, STAGE_BYTES, st>>>(pb, out, 0);
    k_prob_jobs<2><<<g, 128, STAGE_BYTES, st>>>(pb, out, 1);
    k_prob_jobs<4><<<g, 128, STAGE_BYTES, st>>>(pb, out, 2);
    k_prob_jobs<8><<<g, 128, STAGE_BYTES, st>>>(pb, out, 3);
    k_prob_jobs<16><<<g, 128, STAGE_BYTES, st>>>(pb, out, 4);
    k_prob_jobs<32><<<g, 128, STAGE_BYTES, st>>>(pb, out, 5);
    k_prob_jobs<64><<<g, 128, STAGE_BYTES, st>>>(pb, out, 6);
    k_prob_jobs_xl<<<(int)(pb.n < sms ? pb.n : sms), XL_T, 0, st>>>(pb, out);
}

}  // namespace lfb
