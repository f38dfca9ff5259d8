// Device helpers shared by the CUDA translation units: one read of plp_to_errprobs (snpcaller.c:399-491),
// the 4-way merge (snpcaller.c:334), the DBL_EPSILON guards (snpcaller.c:872-881), column geometry.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <float.h>
#include "internal.h"

namespace lfb {

#define FULL 0xffffffffu
static constexpr double LN2 = 0.69314718055994530942;
static constexpr double DEPS = 2.220446049250313e-16;   // DBL_EPSILON

// ------------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(FULL, v, m);
    return v;
}

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// merge_srcq_mapq_baq_and_bq (snpcaller.c:334), same association, no contraction
__device__ __forceinline__ double merge4(double sp, double mp, double bap, double bp)
{
    const double om = __dsub_rn(1.0, mp);
    const double os = __dsub_rn(1.0, sp);
    const double oa = __dsub_rn(1.0, bap);
    double acc = __dadd_rn(mp, __dmul_rn(om, sp));
    const double oms = __dmul_rn(om, os);
    acc = __dadd_rn(acc, __dmul_rn(oms, bap));
    acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(oms, oa), bp));
    return acc;
}

// DBL_EPSILON guards of pruned_calc_prob_dist (snpcaller.c:872-881), in linear space
__device__ __forceinline__ void guard_pq(double jp, double &p, double &q)
{
    p = (fabs(jp) < DEPS) ? DEPS : jp;
    q = (fabs(jp - 1.0) < DEPS) ? (1.0 - jp + DEPS) : (1.0 - jp);
}

struct Geom {            // one column of a batch
    long long off;       // first read in the planes
    int n;               // reads in the column
    int b1, b2, b3;      // group boundaries: A [0,b1) C [b1,b2) G [b2,b3) T [b3,n)
    int ref_idx;         // 0..3, -1 = not A/C/G/T
    double alt_bp;       // base-quality probability forced on alt reads (alt_bq_mode != 0)
};

__device__ __forceinline__ int ref_index(char r)
{
    return r == 'A' ? 0 : r == 'C' ? 1 : r == 'G' ? 2 : r == 'T' ? 3 : -1;
}

// plp_to_errprobs for one read (snpcaller.c:399-491).  Returns false when the read is filtered out.
// lut points at the shared-memory copy of Lut (bq | mq | aq).
template <bool NEEDP>
__device__ __forceinline__ bool eval_read(const DevConf &cf, const double *lut, const Geom &g, int pos, int bq, int mq,
                                          int baq, int sq, bool &is_alt, int &slot, double &jp)
{
    const int grp = (pos >= g.b1) + (pos >= g.b2) + (pos >= g.b3);
    is_alt = grp != g.ref_idx;
    slot = grp - (grp > g.ref_idx);
    if (bq < cf.min_bq) return false;
    if (is_alt && bq < cf.min_alt_bq) return false;
    if (!NEEDP && !cf.jq_filters) return true;
    double bp = lut[bq];
    if (is_alt && cf.alt_bq_mode) bp = g.alt_bp;
    const double mp = cf.use_mq ? lut[256 + mq] : 0.0;
    const double bap = cf.use_baq ? lut[512 + baq] : 0.0;
    const double sp = cf.use_sq ? lut[512 + sq] : 0.0;
    jp = merge4(sp, mp, bap, bp);
    if (cf.jq_filters) {
        if (jp >= cf.skip_jp) return false;
        if (is_alt && jp >= cf.skip_alt_jp) return false;
    }
    if (is_alt && cf.def_alt_jq_on) jp = cf.def_alt_jq_prob;
    return true;
}

// ------------------------------------------------------------------------------------------------
// one read of plp_to_errprobs (snpcaller.c:399-491)
// ------------------------------------------------------------------------------------------------
// UNIFORM: the configuration treats reference and alt reads alike (the defaults: min_alt_bq <= min_bq, no def_alt_bq /
// def_alt_jq, no jq filters) — no position bookkeeping; plain_merge: only bq and mq are merged (sp = bap = 0: the dropped
// terms of merge_srcq_mapq_baq_and_bq are exact zeros and ones)
struct EvalMode {
    bool uniform, plain_merge;
};

__device__ __forceinline__ EvalMode eval_mode(const DevConf &cf)
{
    EvalMode em;
    em.uniform = cf.min_bq >= 0 && cf.min_alt_bq <= cf.min_bq && cf.alt_bq_mode == 0 && !cf.def_alt_jq_on && !cf.jq_filters;
    em.plain_merge = !(cf.use_baq | cf.use_sq);
    return em;
}

// the general read (alt-specific filters and overrides); k_dp keeps it out of line (LFB_EVAL_INLINE, dp_fused.cu)
#ifndef LFB_EVAL_INLINE
#define LFB_EVAL_INLINE __forceinline__
#endif
static __device__ LFB_EVAL_INLINE bool eval_read_general(const DevConf &cf, const double *s_lut, const Geom &g, int pos, int bq, int mq, int baq, int sq,
                                                      double &jp)
{
    bool is_alt;
    int slot;
    return eval_read<true>(cf, s_lut, g, pos, bq, mq, baq, sq, is_alt, slot, jp);
}

__device__ __forceinline__ bool dp_eval(const DevConf &cf, const EvalMode &em, const double *s_lut, const Geom &g, int pos, int bq, int mq,
                                        int baq, int sq, double &jp)
{
    if (em.uniform) {
        if (bq < cf.min_bq) return false;
        const double bp = s_lut[bq];
        if (em.plain_merge) {
            if (!cf.use_mq) { jp = bp; return true; }
            const double mp = s_lut[256 + mq];
            jp = __dadd_rn(mp, __dmul_rn(__dsub_rn(1.0, mp), bp));
            return true;
        }
        jp = merge4(cf.use_sq ? s_lut[512 + sq] : 0.0, cf.use_mq ? s_lut[256 + mq] : 0.0, cf.use_baq ? s_lut[512 + baq] : 0.0, bp);
        return true;
    }
    return eval_read_general(cf, s_lut, g, pos, bq, mq, baq, sq, jp);
}

__device__ __forceinline__ void load_lut(double *s_lut, const Lut *lut)
{
    const double *src = reinterpret_cast<const double *>(lut);
    for (int i = threadIdx.x; i < 768; i += blockDim.x) s_lut[i] = src[i];
    __syncthreads();
}

// median of the reference-base qualities (int_median, utils.c:435-458) for def_alt_bq == -1.
// hist: 256 ints of shared memory owned by this warp.
static __device__ int warp_ref_median(const unsigned char *bqp, long long off, int lo, int hi, int *hist)
{
    const int lane = lane_id();
    for (int i = lane; i < 256; i += 32) hist[i] = 0;
    __syncwarp();
    for (int i = lo + lane; i < hi; i += 32) atomicAdd(&hist[bqp[off + i]], 1);
    __syncwarp();
    int med = -1;
    const int n = hi - lo;
    if (n > 0 && lane == 0) {
        // order statistics n/2 and n/2-1 of the sorted values
        const int r_hi = n / 2, r_lo = n / 2 - 1;
        int acc = 0, v_hi = -1, v_lo = -1;
        for (int v = 0; v < 256; ++v) {
            const int nxt = acc + hist[v];
            if (v_lo < 0 && r_lo >= 0 && r_lo < nxt) v_lo = v;
            if (v_hi < 0 && r_hi < nxt) v_hi = v;
            acc = nxt;
        }
        med = (n & 1) ? v_hi : (int)((v_hi + v_lo) / 2.0);
    }
    med = __shfl_sync(FULL, med, 0);
    __syncwarp();
    return med;
}

__device__ __forceinline__ bool load_geom(const DevBatch &b, long long c, Geom &g, int &cov)
{
    const int4 cnt = reinterpret_cast<const int4 *>(b.nt_cnt)[c];
    g.off = b.col_off[c];
    g.b1 = cnt.x;
    g.b2 = g.b1 + cnt.y;
    g.b3 = g.b2 + cnt.z;
    g.n = g.b3 + cnt.w;
    g.ref_idx = ref_index(b.ref_base[c]);
    g.alt_bp = 0.0;
    cov = b.coverage ? b.coverage[c] : g.n;
    return true;
}

__device__ __forceinline__ void setup_alt_bq(const DevConf &cf, const DevBatch &b, const double *s_lut, Geom &g, int *hist)
{
    if (cf.alt_bq_mode == 1) {
        g.alt_bp = cf.alt_bq_prob;
    } else if (cf.alt_bq_mode == 2) {
        const int lo = g.ref_idx == 0 ? 0 : g.ref_idx == 1 ? g.b1 : g.ref_idx == 2 ? g.b2 : g.b3;
        const int hi = g.ref_idx == 0 ? g.b1 : g.ref_idx == 1 ? g.b2 : g.ref_idx == 2 ? g.b3 : g.n;
        const int med = warp_ref_median(b.bq, g.off, lo, hi, hist);
        g.alt_bp = med < 0 ? 0.0 : s_lut[med];      // empty ref group: bq = -1 -> probability 0 (snpcaller.c:435, 331)
    }
}

// 16 consecutive bytes of each plane
__device__ __forceinline__ int byte_of(const uint4 &v, int j)
{
    const unsigned w = j < 4 ? v.x : j < 8 ? v.y : j < 12 ? v.z : v.w;
    return (w >> (8 * (j & 3))) & 0xff;
}

__device__ __forceinline__ uint4 ldg16(const unsigned char *p)
{
    return __ldg(reinterpret_cast<const uint4 *>(p));
}

struct Chunk16 {        // 16 consecutive bytes of each plane, one lane's share of a 512-read stripe
    uint4 bq, mq, baq, sq;
};

__device__ __forceinline__ void load_chunk(const DevConf &cf, const DevBatch &b, long long a, Chunk16 &ch)
{
    const uint4 zero = make_uint4(0, 0, 0, 0);
    ch.bq = ldg16(b.bq + a);
    ch.mq = cf.use_mq ? ldg16(b.mq + a) : zero;
    ch.baq = cf.use_baq ? ldg16(b.baq + a) : zero;
    ch.sq = cf.use_sq ? ldg16(b.sq + a) : zero;
}

// the same 16 bytes from an address that is only 8-byte aligned
// (the upper half only when the column reaches into it: the planes are readable up to the 16-byte boundary past their
// last read, not further)
__device__ __forceinline__ uint4 ldg8x2(const unsigned char *p, bool hi_too)
{
    const uint2 lo = __ldg(reinterpret_cast<const uint2 *>(p));
    const uint2 hi = hi_too ? __ldg(reinterpret_cast<const uint2 *>(p) + 1) : make_uint2(0u, 0u);
    return make_uint4(lo.x, lo.y, hi.x, hi.y);
}

__device__ __forceinline__ void load_chunk8(const DevConf &cf, const DevBatch &b, long long a, bool hi_too, Chunk16 &ch)
{
    const uint4 zero = make_uint4(0, 0, 0, 0);
    ch.bq = ldg8x2(b.bq + a, hi_too);
    ch.mq = cf.use_mq ? ldg8x2(b.mq + a, hi_too) : zero;
    ch.baq = cf.use_baq ? ldg8x2(b.baq + a, hi_too) : zero;
    ch.sq = cf.use_sq ? ldg8x2(b.sq + a, hi_too) : zero;
}

__device__ __forceinline__ int class_of(int K)
{
    const int need = (K + 31) >> 5;          // cells per lane
    if (need > 64) return CLS_XL;
    int cls = 0;
    while ((1 << cls) < need) ++cls;
    return cls;
}


// a candidate was emitted for column c: one mark per column, counted per tile of 256 columns (k_rank_cands)
__device__ __forceinline__ void mark_cand(const Workspace &ws, long long c)
{
    ws.is_cand[c] = 1;
    atomicAdd(&ws.candtile[c >> 8], 1u);
}

// k_dp job list of a column with largest alt count K (KS < K <= DP_MAXK) and n reads: class * DP_NBIN1 + depth bin
__device__ __forceinline__ int dp_class(int K)
{
    return K <= 32 ? 0 : K <= 64 ? 1 : K <= 128 ? 2 : K <= 256 ? 3 : K <= 512 ? 4 : K <= 1024 ? 5 : 6;
}

__device__ __forceinline__ int dp_list(int K, int n)
{
    int bin = 0;
    while (bin < DP_NBIN - 1 && n > (128 << bin)) ++bin;
    return dp_class(K) * DP_NBIN1 + bin;
}
}  // namespace lfb
