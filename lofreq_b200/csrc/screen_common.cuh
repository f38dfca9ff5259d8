// Device helpers of the stream side (k_front, k_prune2): column metadata, alt-read counting, the lane-per-column prune.
#pragma once
#include "dev_common.cuh"

namespace lfb {

constexpr int FIN_BLOCK = 256;      // columns per tile of the prefix sums = per CTA of k_front
constexpr int PRUNE_CAP1 = 8;       // reads k_front itself looks at (K <= 3 is decided by then); the rest of the prune is k_prune2's
constexpr int PRUNE_CAP = 32;       // reads the lane-per-column prune looks at before it hands the column to k_mid (4x / 8x more when the exit is within reach)

// running Bonferroni factor of the tested column with 1-based rank `rank` in a batch that starts from `start`
// (lofreq_call.c:794-800: the first tested column sets 3 when bonf_subst was 1, else += 3)
__device__ __forceinline__ long long bonf_of(const DevConf &cf, long long start, long long rank)
{
    return cf.bonf_dynamic ? ((start == 1 ? 0 : start) + 3 * rank) : start;
}

// 1-based rank of a tested column among the tested columns of the batch (0 = untested): prefix of its tile of 256
// columns + tested columns of the warps before its own in the tile + rank among its warp's 32 columns
__device__ __forceinline__ long long col_rank(const Workspace &ws, long long c)
{
    const int r = ws.rank[c];
    if (!r) return 0;
    const uint2 x8 = __ldg(reinterpret_cast<const uint2 *>(ws.wcount) + (c >> 8));
    const int wi = (int)(c >> 5) & 7;
    const unsigned lo = wi >= 4 ? x8.x : (x8.x & ((1u << (8 * wi)) - 1u));
    const unsigned hi = wi > 4 ? (x8.y & ((1u << (8 * (wi - 4))) - 1u)) : 0u;
    return ws.blocksum[c >> 8] + (long long)(__vsadu4(lo, 0u) + __vsadu4(hi, 0u)) + r;
}

// [lo, hi) of the reads showing the reference base
__device__ __forceinline__ void ref_range(const Geom &g, int &lo, int &hi)
{
    lo = g.ref_idx == 0 ? 0 : g.ref_idx == 1 ? g.b1 : g.ref_idx == 2 ? g.b2 : g.b3;
    hi = g.ref_idx == 0 ? g.b1 : g.ref_idx == 1 ? g.b2 : g.ref_idx == 2 ? g.b3 : g.n;
}

struct RawGeom {        // the per-column metadata as loaded, one column ahead of its use
    long long off;
    int4 cnt;
    int cov, nb;
    char ref;
};

__device__ __forceinline__ void load_raw(const DevBatch &b, long long c, RawGeom &r)
{
    r.cnt = __ldg(reinterpret_cast<const int4 *>(b.nt_cnt) + c);
    r.off = __ldg(b.col_off + c);
    r.ref = __ldg(b.ref_base + c);
    r.cov = b.coverage ? __ldg(b.coverage + c) : -1;
    r.nb = b.num_bases ? __ldg(b.num_bases + c) : -1;
}

// gates and alt counts.  Only reads that show a non-reference base decide whether a column is
// tested and what K is (snpcaller.c:418-420,489), so only those bytes are touched.  A warp takes 32
// consecutive columns: metadata, gates and columns with at most 8 non-reference reads run lane-per-column;
// the rare columns with more (variant sites) are then counted by the whole warp.
__device__ __forceinline__ void count_alt_read(const DevConf &cf, const DevBatch &b, const double *s_lut, const Geom &g,
                                               int ref_lo, int ref_hi, int i, int (&cnt)[3], int (&raw)[3], int bq_ready = -1)
{
    const int pos = i < ref_lo ? i : i - ref_lo + ref_hi;
    const long long a = g.off + pos;
    const int bq = bq_ready >= 0 ? bq_ready : b.bq[a];
    int mq = 0, baq = 0, sq = 0;
    if (cf.jq_filters) {
        if (cf.use_mq) mq = b.mq[a];
        if (cf.use_baq) baq = b.baq[a];
        if (cf.use_sq) sq = b.sq[a];
    }
    bool is_alt;
    int slot;
    double jp;
    const bool ok = eval_read<false>(cf, s_lut, g, pos, bq, mq, baq, sq, is_alt, slot, jp);
    raw[0] += slot == 0;              // raw counts precede every filter (snpcaller.c:418-420)
    raw[1] += slot == 1;
    raw[2] += slot == 2;
    if (ok) {
        cnt[0] += slot == 0;
        cnt[1] += slot == 1;
        cnt[2] += slot == 2;
    }
}

// The reference's early exit (snpcaller.c:916-958), one lane per column: walk the first `cap` reads until
// P(X >= K among the reads seen) > limit = sig / bonf.  Returns true when the column is still alive after `cap` reads.
// Cells are kept top-aligned (register KP-1 = cell K-1, padding below cell 0 stays 0), so one code path serves every
// K <= KP.  Lanes with live == false only take part in the votes.
// ext: a deep column still alive after cap_reads reads goes on when the tail it has reached says the early exit is
// within reach — P(X >= K among n reads) grows about like n^K, so 8 times more reads close a gap of 8^K.
template <int KP>
__device__ __forceinline__ bool lane_prune(const DevConf &cf, const DevBatch &b, const double *s_lut, const Geom &mg, int K,
                                           double limit, int cap_reads, bool live, const Chunk16 *first = nullptr, bool ext = false,
                                           int first_align = 16, bool bq_only = false)
{
    double R[KP], T = 0.0;
#pragma unroll
    for (int j = 0; j < KP; ++j) R[j] = (j == KP - K) ? 1.0 : 0.0;
    const EvalMode em = eval_mode(cf);
    int cap = min(mg.n, cap_reads);
    // only deep columns go on: below ~2000 reads evaluating everything in k_mid (beside k_dp) is cheaper than a longer
    // lane-serial walk on the critical path
    const int fshift = 3;
    const int cap2 = (ext && mg.n >= 2048) ? min(mg.n, cap_reads << fshift) : 0;
    // the reads come in aligned 16-byte chunks per plane: one load per plane covers what most columns need; the chunk
    // after the current one is requested while the current one is walked
    // (first_align == 8: the caller's first chunk starts at the 8-byte boundary below the column and cap_reads <= 9, so
    // the walk never leaves it)
    const long long ca = mg.off & ~(long long)(first_align - 1);
    const int lead = (int)(mg.off - ca);
    Chunk16 ch, nx;
    ch.bq = ch.mq = ch.baq = ch.sq = make_uint4(0, 0, 0, 0);
    nx = ch;
    if (first) ch = *first;                     // the caller requested the first chunk ahead of time
    else if (live && cap > 0) load_chunk(cf, b, ca, ch);
    if (ext && live && lead + cap > 16) load_chunk(cf, b, ca + 16, nx);
#pragma unroll 1
    for (int i = 0; __any_sync(FULL, live && i < cap); ++i) {
        if (!(live && i < cap)) continue;
        const int idx = lead + i, j = idx & 15;
        if (j == 0 && i > 0) {
            if (ext) {
                ch = nx;
                if (idx + 16 < lead + max(cap, cap2)) load_chunk(cf, b, ca + idx + 16, nx);
            } else {
                load_chunk(cf, b, ca + idx, ch);
            }
        }
        double jp;
        // bq_only (k_front, uniform configurations): the base quality alone.  Every other quality only raises a read's error
        // probability, and the tail grows with every probability — a column ruled out on base qualities alone is ruled out.
        if (bq_only ? !dp_eval(cf, em, s_lut, mg, i, byte_of(ch.bq, j), 255, 255, 255, jp)
                    : !dp_eval(cf, em, s_lut, mg, i, byte_of(ch.bq, j), byte_of(ch.mq, j), byte_of(ch.baq, j), byte_of(ch.sq, j), jp))
            continue;
        double p, q;
        guard_pq(jp, p, q);
        T = fma(R[KP - 1], p, T);
#pragma unroll
        for (int j2 = KP - 1; j2 >= 1; --j2) R[j2] = fma(R[j2 - 1], p, R[j2] * q);
        R[0] = R[0] * q;
        if (T > limit) live = false;          // clearly insignificant: snpcaller() leaves LDBL_MAX everywhere (snpcaller.c:1155)
        if (live && i + 1 == cap && cap < cap2 && T * __hiloint2double((1023 + fshift * K) << 20, 0) >= limit) cap = cap2;
    }
    return live;
}

}  // namespace lfb
