// k_front: the stream side of the test in one pass over the batch.
//
// Per column, what call_vars()/call_snvs() do before snpcaller() (lofreq_call.c:886-935, 734-801): the gates, the
// filtered and raw alt counts of plp_to_errprobs (only the reads showing a non-reference base decide them,
// snpcaller.c:418-420,489), "is this column a test", its place in the running Bonferroni count
// (lofreq_call.c:794-800) — and then at once the reference's early exit (snpcaller.c:916-958), one lane per column,
// on the first reads of the column: with a real Bonferroni factor nearly every tested column is ruled out after a
// handful of reads, without a warp ever being dedicated to it.
//
// Persistent CTAs, thread per column, tile = 256 consecutive columns; CTA b takes tiles b, b + G, b + 2G, ... (round
// k = tiles [kG, (k+1)G)).  No CTA ever waits for another one, and no warp for another warp (the kernel has no barrier:
// a warp whose 32 columns are dull runs ahead into the next round while its neighbours still walk theirs):
//   * the exact place of a column in the running count needs the tested columns of ALL tiles before it.  The prune
//     does not: it only needs a factor that is not larger than the true one (a column ruled out under a smaller
//     factor is ruled out under the larger one).  The tested columns of a tile are added to a counter of its round
//     (the last of the tile's warps to get there adds them); a column of round k uses start + 3 * (whatever the
//     counters of rounds < k hold at that moment — all of it belongs to tiles before this one — + its rank among the
//     32 columns of its warp).  After the first round that is within a fraction of the exact factor.
//   * the exact ranks are tile prefix + tested columns of the warps before it in the tile + rank inside the warp:
//     k_front leaves the per-warp counts (a byte each, eight per tile) and the ranks inside the warp (a byte each),
//     k_scan_tiles (one CTA) turns the counts into exclusive tile prefixes.  Everything that survives the prune is
//     decided with the exact factor afterwards (k_prune2 and later: col_rank()).
// Per round and tile: metadata (requested a round earlier), gates; the alt reads and the first 16 reads of every column
// that has alt reads are requested together; alt counts — lane per column for up to 8 non-reference reads, the whole
// warp for the rare columns with more (variant sites); routing (K > 8 -> job list of k_dp / k_xl) and the prune over
// the first PRUNE_CAP1 reads, survivors to k_prune2's list.
// When region shards on several GPUs continue each other's count (lfb200_comm_exchange), the shards before this one add
// to the start AFTER this pass — one more reason why the factor used here is a lower bound.
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <float.h>
#include <stdlib.h>
#include "internal.h"
#include "dev_common.cuh"
#include "screen_common.cuh"

namespace lfb {

__device__ __forceinline__ unsigned int ld_relaxed_u32(const unsigned int *p)
{
    unsigned int v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

#ifndef LFB_FRONT_T
#define LFB_FRONT_T 256
#endif
constexpr int FRONT_T = LFB_FRONT_T;      // threads per CTA = columns per unit of work (a half or a whole tile of FIN_BLOCK columns)
#ifndef LFB_FRONT_CTAS
#define LFB_FRONT_CTAS (768 / LFB_FRONT_T)
#endif
constexpr int FRONT_CTAS_PER_SM = LFB_FRONT_CTAS;
#ifndef LFB_KS1
#define LFB_KS1 4
#endif
constexpr int KS1 = LFB_KS1;          // largest alt count the first stage of the prune walks itself

__global__ void __launch_bounds__(FRONT_T, FRONT_CTAS_PER_SM) k_front(const __grid_constant__ DevConf cf, const __grid_constant__ DevBatch b,
                                                                        const Lut *lut, const Workspace ws)
{
    __shared__ double s_lut[768];
    __shared__ int s_hist[FRONT_T / 32][256];
    __shared__ unsigned int s_round[FRONT_MAXROUNDS];     // per round of this CTA: (warps arrived << 16) + tested columns
    for (int i = threadIdx.x; i < FRONT_MAXROUNDS; i += blockDim.x) s_round[i] = 0;
    load_lut(s_lut, lut);                                 // ends with the kernel's only barrier
    const long long n = b.n_cols;
    // units of FRONT_T columns; both halves of the last tile are walked (the per-warp counts of a tile are read as one word)
    const long long ntiles = (n + FIN_BLOCK - 1) / FIN_BLOCK * (FIN_BLOCK / FRONT_T);
    const int lane = lane_id(), w = threadIdx.x >> 5;
    if (blockIdx.x == 0 && threadIdx.x == 0) ws.counters->bonf_start_used = cf.bonf_start;
    // the median override (def_alt_bq == -1) needs a warp-wide histogram: no lane-per-column path then
    const int serial_max = cf.alt_bq_mode == 2 ? 0 : 8;
    unsigned int *round_cnt = ws.counters->front_round;
    const bool bq_only = eval_mode(cf).uniform;

    RawGeom nxt;
    nxt.off = 0; nxt.cnt = make_int4(0, 0, 0, 0); nxt.cov = -1; nxt.nb = -1; nxt.ref = 'N';
    {
        const long long c0 = (long long)blockIdx.x * FRONT_T + threadIdx.x;
        if (c0 < n) load_raw(b, c0, nxt);
    }
    int round = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++round) {
        const long long c = tile * FRONT_T + threadIdx.x;
        const RawGeom raw_g = nxt;
        {
            const long long cn = c + (long long)gridDim.x * FRONT_T;
            nxt.off = 0; nxt.cnt = make_int4(0, 0, 0, 0); nxt.cov = -1; nxt.nb = -1; nxt.ref = 'N';
            if (cn < n) load_raw(b, cn, nxt);                  // next round's metadata in flight
        }
        // What the rounds before this one have counted so far — every tile of an earlier round lies before this tile, so any
        // snapshot of their counters is a lower bound of the tested columns before it.  Requested now, used after the counts.
        // (lane j takes the counter of round j; summed where the factor is needed)
        unsigned int before_l = 0;
        if (lane < min(round, FRONT_MAXROUNDS)) before_l = ld_relaxed_u32(&round_cnt[lane]);
        // ---- metadata, gates ----
        Geom mg;
        mg.off = raw_g.off;
        mg.b1 = raw_g.cnt.x;
        mg.b2 = mg.b1 + raw_g.cnt.y;
        mg.b3 = mg.b2 + raw_g.cnt.z;
        mg.n = mg.b3 + raw_g.cnt.w;
        mg.ref_idx = ref_index(raw_g.ref);
        mg.alt_bp = cf.alt_bq_prob;
        const int m_cov = raw_g.cov < 0 ? mg.n : raw_g.cov;
        const int m_nb = raw_g.nb < 0 ? mg.n : raw_g.nb;           // plp_col_t.num_bases
        const bool m_gate = c < n && mg.ref_idx >= 0 && !(m_nb * 2 < m_cov) && !(m_nb < cf.min_cov);   // lofreq_call.c:892,931,747,754
        int m_lo, m_hi;
        ref_range(mg, m_lo, m_hi);
        const int m_alt = m_gate ? mg.n - (m_hi - m_lo) : 0;
        // the first PRUNE_CAP1 reads of a column the prune may walk (16 bytes per plane from the 8-byte boundary below the
        // column hold them wherever the column starts): requested together with its alt reads
        Chunk16 first;
        first.bq = first.mq = first.baq = first.sq = make_uint4(0, 0, 0, 0);
        if (m_alt > 0 && cf.alt_bq_mode != 2) {
            const long long a8 = mg.off & ~7ll;
            const bool hi_too = a8 + 8 < mg.off + mg.n;
            // (uniform configurations: the first stage prunes on the base qualities alone, see lane_prune)
            if (bq_only) first.bq = ldg8x2(b.bq + a8, hi_too);
            else load_chunk8(cf, b, a8, hi_too, first);
        }
        // ---- alt counts ----
        int cnt[3] = {0, 0, 0}, raw[3] = {0, 0, 0};
        if (m_alt > 0 && m_alt <= serial_max) {
            // the base qualities of the first four non-reference reads are requested together: one memory latency per
            // round instead of one per read
            int bq4[4] = {-1, -1, -1, -1};
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (i < m_alt) bq4[i] = b.bq[mg.off + (i < m_lo ? i : i - m_lo + m_hi)];
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (i < m_alt) count_alt_read(cf, b, s_lut, mg, m_lo, m_hi, i, cnt, raw, bq4[i]);
            for (int i = 4; i < m_alt; ++i) count_alt_read(cf, b, s_lut, mg, m_lo, m_hi, i, cnt, raw);
        }
        // whole warp per column with many non-reference reads
        unsigned todo = __ballot_sync(FULL, m_alt > serial_max);
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            Geom g;
            g.off = __shfl_sync(FULL, mg.off, src);
            g.b1 = __shfl_sync(FULL, mg.b1, src);
            g.b2 = __shfl_sync(FULL, mg.b2, src);
            g.b3 = __shfl_sync(FULL, mg.b3, src);
            g.n = __shfl_sync(FULL, mg.n, src);
            g.ref_idx = __shfl_sync(FULL, mg.ref_idx, src);
            g.alt_bp = 0.0;
            int ref_lo, ref_hi;
            ref_range(g, ref_lo, ref_hi);
            const int n_alt = g.n - (ref_hi - ref_lo);
            setup_alt_bq(cf, b, s_lut, g, s_hist[w]);
            int wc[3] = {0, 0, 0}, wr[3] = {0, 0, 0};
            for (int i = lane; i < n_alt; i += 32) count_alt_read(cf, b, s_lut, g, ref_lo, ref_hi, i, wc, wr);
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                wc[i] = __reduce_add_sync(FULL, wc[i]);
                wr[i] = __reduce_add_sync(FULL, wr[i]);
                if (lane == src) { cnt[i] = wc[i]; raw[i] = wr[i]; }
            }
        }
        // no alt read left after filtering -> not a test (lofreq_call.c:768-780)
        const int t = (c < n && (cnt[0] | cnt[1] | cnt[2]) != 0) ? 1 : 0;
        if (c < n) {
            int2 *o = reinterpret_cast<int2 *>(ws.cnt6 + 6 * c);
            o[0] = make_int2(cnt[0], cnt[1]);
            o[1] = make_int2(cnt[2], raw[0]);
            o[2] = make_int2(raw[1], raw[2]);
            ws.tested[c] = (unsigned char)t;
        }
        // ---- rank inside the warp; the warp's count for the exact prefix (k_scan_tiles, col_rank) and, through the CTA's
        // counter of the round, for the rounds after this one
        const unsigned bal = __ballot_sync(FULL, t);
        const int lrank = t ? __popc(bal & ((2u << lane) - 1u)) : 0;      // 1-based among the tested columns of the warp
        if (c < n) ws.rank[c] = (unsigned char)lrank;
        if (lane == 0) {
            ws.wcount[tile * (FRONT_T / 32) + w] = (unsigned char)__popc(bal);
            if (round < FRONT_MAXROUNDS) {
                // the warps of a CTA drift apart by whole rounds; the last one of this round hands the tile's count on
                const unsigned v = atomicAdd(&s_round[round], (1u << 16) + (unsigned)__popc(bal)) + (1u << 16) + (unsigned)__popc(bal);
                if ((v >> 16) == FRONT_T / 32 && (v & 0xffffu)) atomicAdd(&round_cnt[round], v & 0xffffu);
            }
        }
        const long long before = (long long)__reduce_add_sync(FULL, before_l);
        // factor for the prune: never larger than the true one
        const long long bonf = t ? bonf_of(cf, cf.bonf_start, before + lrank) : 0;

        // ---- routing and the first stage of the prune ----
        const int K = max(cnt[0], max(cnt[1], cnt[2]));
        // K > KS: a slot in a job list of k_dp / k_xl.  The slot is requested here and used after the prune walk below, which
        // hides the round trip of the atomic.
        unsigned slot_a = 0;
        int li = -1;
        if (t && K > KS) {
            if (K <= DP_MAXK) {
                // 8 < K <= 2048: k_dp, several columns per warp; job list by (class, depth bin)
                li = dp_list(K, mg.n);
                slot_a = atomicAdd(&ws.counters->n_pjobs[li], 1u);
            } else {
                slot_a = atomicAdd(&ws.counters->n_jobs[CLS_XL], 1u);      // one CTA per column
            }
        }
        // K <= KS: with the Bonferroni factors of a real run the early exit fires after a handful of reads (K = 1: one;
        // K = 3: ~6; K = 4: ~27 at depth-500 Q30), so almost every column ends here.  A warp runs as long as its slowest
        // lane, and the few columns with K >= 4 would keep 31 finished lanes waiting: this kernel stops after PRUNE_CAP1
        // reads and lists what is still alive for k_prune2, whose warps are full of such columns.
        bool small = t && K <= KS;
        const double limit = small ? cf.sig * (1.0 + 1e-9) / (double)bonf : 0.0;   // margin: borderline columns go to the host
        if (cf.alt_bq_mode != 2) {          // the median override needs a warp-wide histogram: no lane-serial prune
            // (K > KS1 cannot be ruled out within PRUNE_CAP1 reads under any real factor: listed without a walk)
            const bool alive = lane_prune<KS1>(cf, b, s_lut, mg, K, limit, PRUNE_CAP1, small && K <= KS1, &first, false, 8, bq_only);
            small = small && (K > KS1 || alive);
            // one atomic per warp for the survivors' slots
            const unsigned sm_ = __ballot_sync(FULL, small);
            if (sm_) {
                unsigned base = 0;
                if (lane == __ffs(sm_) - 1) base = atomicAdd(&ws.counters->n_jobs[CLS_PRUNE2], (unsigned)__popc(sm_));
                base = __shfl_sync(FULL, base, __ffs(sm_) - 1);
                if (small) ws.jobs[(long long)CLS_PRUNE2 * ws.cap_cols + base + __popc(sm_ & ((1u << lane) - 1u))] = (int)c;
            }
        } else if (small) {
            // every small column joins k_mid's job list: full evaluation, whole warp
            const unsigned slot = atomicAdd(&ws.counters->n_jobs[0], 1u);
            ws.jobs[slot] = (int)c;
        }
        if (t && K > KS) {
            if (li >= 0) {
                // the class's unbinned list when the binned one is full
                const int cls = li / DP_NBIN1;
                if (slot_a < (unsigned)ws.pcap) {
                    ws.pjobs[((long long)cls * DP_NBIN + (li % DP_NBIN1)) * ws.pcap + slot_a] = (int)c;
                } else {
                    const unsigned s2 = atomicAdd(&ws.counters->n_pjobs[cls * DP_NBIN1 + DP_NBIN], 1u);
                    ws.ujobs[(long long)cls * ws.cap_cols + s2] = (int)c;
                }
            } else {
                ws.jobs[(long long)CLS_XL * ws.cap_cols + slot_a] = (int)c;
            }
        }
    }
}

// exclusive prefix over the tiles' tested columns (the eight per-warp bytes k_front left per tile), one CTA;
// total -> counters->n_tested
__global__ void __launch_bounds__(1024) k_scan_tiles(const unsigned char *wcount, long long *blocksum, int nb, unsigned long long *total)
{
    __shared__ long long s_w[32];
    __shared__ long long s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = lane_id(), w = threadIdx.x >> 5;
    for (int base = 0; base < nb; base += 1024) {
        const int i = base + threadIdx.x;
        long long v = 0;
        if (i < nb) {
            const uint2 x8 = reinterpret_cast<const uint2 *>(wcount)[i];
            v = __vsadu4(x8.x, 0u) + __vsadu4(x8.y, 0u);
        }
        long long x = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const long long y = __shfl_up_sync(FULL, x, d);
            if (lane >= d) x += y;
        }
        if (lane == 31) s_w[w] = x;
        __syncthreads();
        if (w == 0) {
            long long z = s_w[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const long long y = __shfl_up_sync(FULL, z, d);
                if (lane >= d) z += y;
            }
            s_w[lane] = z;
        }
        __syncthreads();
        const long long incl = x + (w ? s_w[w - 1] : 0) + s_carry;
        if (i < nb) blocksum[i] = incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = (unsigned long long)s_carry;
}

// bonf_used[] for callers that ask for the dense per-column output (lfb200_device_results, lfb200_dense_out_t)
__global__ void k_bonf_used(const __grid_constant__ DevConf cf, const Workspace ws, long long n)
{
    const long long start = ws.counters->bonf_start_used;
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (long long)gridDim.x * blockDim.x) {
        const long long r = col_rank(ws, c);
        ws.bonf_used[c] = r ? bonf_of(cf, start, r) : 0;
    }
}

void launch_front(const LaunchState &ls, const DevConf &cf, const DevBatch &b, const Lut *lut, const Workspace &ws, cudaStream_t st)
{
    if (b.n_cols <= 0) return;
    const int nb = (int)((b.n_cols + FIN_BLOCK - 1) / FIN_BLOCK);
    // per-batch state: counters (job lists, round counts), candidate marks
    cudaMemsetAsync(ws.counters, 0, sizeof(Counters), st);
    cudaMemsetAsync(ws.is_cand, 0, (size_t)nb * FIN_BLOCK, st);
    static const int front_ctas = getenv("LFB200_FRONT_CTAS") ? atoi(getenv("LFB200_FRONT_CTAS")) : FRONT_CTAS_PER_SM;
    const int nu = nb * (FIN_BLOCK / FRONT_T);
    const int grid = nu < ls.sms * front_ctas ? nu : ls.sms * front_ctas;
    k_front<<<grid, FRONT_T, 0, st>>>(cf, b, lut, ws);
    k_scan_tiles<<<1, 1024, 0, st>>>(ws.wcount, ws.blocksum, nb, &ws.counters->n_tested);
}

void launch_bonf_used(const LaunchState &ls, const DevConf &cf, const Workspace &ws, long long n, cudaStream_t st)
{
    if (n <= 0) return;
    k_bonf_used<<<ls.sms * 4, 256, 0, st>>>(cf, ws, n);
}

}  // namespace lfb
