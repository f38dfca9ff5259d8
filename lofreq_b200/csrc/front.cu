// k_front: the stream side of the test in one pass over the batch.
//
// Per column, what call_vars()/call_snvs() do before snpcaller() (lofreq_call.c:886-935, 734-801): the gates, the
// filtered and raw alt counts of plp_to_errprobs (only the reads showing a non-reference base decide them,
// snpcaller.c:418-420,489), "is this column a test", its place in the running Bonferroni count
// (lofreq_call.c:794-800) — and then at once the reference's early exit (snpcaller.c:916-958), one lane per column,
// on the first reads of the column: with a real Bonferroni factor nearly every tested column is ruled out after a
// handful of reads, without a warp ever being dedicated to it.
//
// Persistent CTAs (one resident wave), thread per column, tile = 256 consecutive columns, CTA b takes tiles
// b, b + G, b + 2G, ...  Every round is software-pipelined over two tiles so that nothing waits for memory it has just
// asked for:
//   A(k)   metadata of tile k (requested a round earlier), gates, alt counts — lane per column for up to 8
//          non-reference reads, the whole warp for the rare columns with more (variant sites); tested columns of the
//          tile -> its aggregate is PUBLISHED (single-pass prefix sum with decoupled look-back); the first 16 reads of
//          the columns the prune will walk and the metadata of tile k+1 are requested;
//   C(k-1) the look-back of tile k-1 — its predecessors published their aggregates a round ago, so it finds them
//          without spinning — gives every tested column its 1-based rank among the tested columns of the batch;
//          factor = start + 3 * rank; K > 8 -> job list of k_dp / k_xl; K <= 8 -> prune over the first PRUNE_CAP1
//          reads (already in registers), survivors to k_prune2's list.
// The factor a batch starts from is the caller's conf->bonf_subst.  When region shards on several GPUs continue each
// other's count (lfb200_comm_exchange), the shards before this one add to it AFTER this pass: the factor used here is
// then a lower bound of the true one, which makes the prune conservative (a column ruled out under a smaller factor is
// ruled out under the larger one), and everything that survives is decided with the exact factor in the second phase
// (k_prune2 and later read rank[] and the exchanged start).
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <float.h>
#include "internal.h"
#include "dev_common.cuh"
#include "screen_common.cuh"

namespace lfb {

// tile_state: bits 62..63 = status, low 62 bits = value
constexpr unsigned long long TS_AGG = 1ull << 62, TS_PREFIX = 2ull << 62, TS_MASK = (1ull << 62) - 1;

__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void st_release_gpu(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

constexpr int FRONT_CTAS_PER_SM = 3;

// what phase C needs of a column counted in phase A one round earlier
struct FrontCol {
    Geom mg;
    int cnt[3];
    int t;               // tested
    Chunk16 first;       // its first 16 reads (requested in phase A)
};

__global__ void __launch_bounds__(FIN_BLOCK, FRONT_CTAS_PER_SM) k_front(const __grid_constant__ DevConf cf, const __grid_constant__ DevBatch b,
                                                                        const Lut *lut, const Workspace ws)
{
    __shared__ double s_lut[768];
    __shared__ int s_hist[FIN_BLOCK / 32][256];
    __shared__ int s_warp[2][FIN_BLOCK / 32];       // inclusive per-warp counts of the tile counted in this round / the one before
    __shared__ long long s_excl;
    load_lut(s_lut, lut);
    const long long n = b.n_cols;
    const long long ntiles = (n + FIN_BLOCK - 1) / FIN_BLOCK;
    const int lane = lane_id(), w = threadIdx.x >> 5;
    unsigned long long *state = reinterpret_cast<unsigned long long *>(ws.blocksum);
    if (blockIdx.x == 0 && threadIdx.x == 0) ws.counters->bonf_start_used = cf.bonf_start;
    // the median override (def_alt_bq == -1) needs a warp-wide histogram: no lane-per-column path then
    const int serial_max = cf.alt_bq_mode == 2 ? 0 : 8;

    RawGeom nxt;
    nxt.off = 0; nxt.cnt = make_int4(0, 0, 0, 0); nxt.cov = -1; nxt.nb = -1; nxt.ref = 'N';
    {
        const long long c0 = (long long)blockIdx.x * FIN_BLOCK + threadIdx.x;
        if (c0 < n) load_raw(b, c0, nxt);
    }
    FrontCol prev;                                   // tile of the round before
    prev.t = 0;
    unsigned prev_bal = 0;
    long long prev_tile = -1, prev_total = 0;
    int par = 0;
    for (long long tile = blockIdx.x;; tile += gridDim.x, par ^= 1) {
        const bool have_tile = tile < ntiles;
        if (!have_tile && prev_tile < 0) break;
        FrontCol cur;
        cur.t = 0;
        unsigned bal = 0;
        long long total = 0;
        const long long c = tile * FIN_BLOCK + threadIdx.x;
        if (have_tile) {
            // ---- A. metadata, gates, alt counts of tile `tile` ----
            const RawGeom raw_g = nxt;
            {
                const long long cn = c + (long long)gridDim.x * FIN_BLOCK;
                nxt.off = 0; nxt.cnt = make_int4(0, 0, 0, 0); nxt.cov = -1; nxt.nb = -1; nxt.ref = 'N';
                if (cn < n) load_raw(b, cn, nxt);                  // next round's metadata in flight
            }
            Geom &mg = cur.mg;
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
            int cnt[3] = {0, 0, 0}, raw[3] = {0, 0, 0};
            if (m_alt > 0 && m_alt <= serial_max)
                for (int i = 0; i < m_alt; ++i) count_alt_read(cf, b, s_lut, mg, m_lo, m_hi, i, cnt, raw);
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
            cur.t = (c < n && (cnt[0] | cnt[1] | cnt[2]) != 0) ? 1 : 0;
#pragma unroll
            for (int i = 0; i < 3; ++i) cur.cnt[i] = cnt[i];
            if (c < n) {
                int2 *o = reinterpret_cast<int2 *>(ws.cnt6 + 6 * c);
                o[0] = make_int2(cnt[0], cnt[1]);
                o[1] = make_int2(cnt[2], raw[0]);
                o[2] = make_int2(raw[1], raw[2]);
                ws.tested[c] = (unsigned char)cur.t;
            }
            // the first 16 reads of a column the prune will walk: requested now, used a round later
            const int K = max(cnt[0], max(cnt[1], cnt[2]));
            cur.first.bq = cur.first.mq = cur.first.baq = cur.first.sq = make_uint4(0, 0, 0, 0);
            if (cur.t && K <= KS && cf.alt_bq_mode != 2 && mg.n > 0) load_chunk(cf, b, mg.off & ~15ll, cur.first);
            bal = __ballot_sync(FULL, cur.t);
            if (lane == 0) s_warp[par][w] = __popc(bal);
        }
        __syncthreads();                                     // (1) per-warp counts of this round's tile are in shared memory
        if (w == 0) {
            if (have_tile) {
                // inclusive scan over the warps of the tile, aggregate published at once
                int z = lane < FIN_BLOCK / 32 ? s_warp[par][lane] : 0;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int y = __shfl_up_sync(FULL, z, d);
                    if (lane >= d) z += y;
                }
                total = __shfl_sync(FULL, z, FIN_BLOCK / 32 - 1);
                if (lane < FIN_BLOCK / 32) s_warp[par][lane] = z;
                if (lane == 0) st_release_gpu(&state[tile], (tile == 0 ? TS_PREFIX : TS_AGG) | (unsigned long long)total);
            }
            if (prev_tile >= 0) {
                // ---- B. look back from the tile of the round before: its predecessors have long published ----
                long long excl = 0;
                if (prev_tile > 0) {
                    long long base = prev_tile - 1;
                    for (;;) {
                        unsigned long long v[8];
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const long long idx = base - (lane * 8 + k);
                            v[k] = idx >= 0 ? ld_acquire_gpu(&state[idx]) : TS_PREFIX;       // before the first tile: prefix 0
                        }
                        long long add = 0;
                        bool found = false;
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const long long idx = base - (lane * 8 + k);
                            while ((v[k] >> 62) == 0) v[k] = ld_acquire_gpu(&state[idx]);      // not counted yet: wait for its aggregate
                            if (!found) add += (long long)(v[k] & TS_MASK);
                            found = found || (v[k] >> 62) == 2;                                  // nearest tile whose inclusive prefix is known
                        }
                        const unsigned pm = __ballot_sync(FULL, found);
                        const int stop = pm ? __ffs(pm) - 1 : 32;
                        if (lane > stop) add = 0;
#pragma unroll
                        for (int m = 16; m >= 1; m >>= 1) add += __shfl_xor_sync(FULL, add, m);
                        excl += add;
                        if (pm) break;
                        base -= 256;
                    }
                    if (lane == 0) st_release_gpu(&state[prev_tile], TS_PREFIX | (unsigned long long)(excl + prev_total));
                }
                if (lane == 0) {
                    s_excl = excl;
                    if (prev_tile == ntiles - 1) ws.counters->n_tested = (unsigned long long)(excl + prev_total);
                }
            }
        }
        total = have_tile ? 0 : 0;                           // (only warp 0 knows it; carried below through shared memory)
        __syncthreads();                                     // (2) s_excl of the previous tile, scanned counts of this one
        if (prev_tile >= 0) {
            // ---- C. rank, routing and the first stage of the prune for the tile of the round before ----
            const long long pc = prev_tile * FIN_BLOCK + threadIdx.x;
            const int t = prev.t;
            int rank = 0;
            long long bonf = 0;
            if (t) {
                // 1-based rank of this column among the tested columns of the batch
                rank = (int)(s_excl + (w ? s_warp[par ^ 1][w - 1] : 0) + __popc(prev_bal & ((2u << lane) - 1u)));
                bonf = bonf_of(cf, cf.bonf_start, rank);
            }
            if (pc < n) ws.rank[pc] = rank;
            const int K = max(prev.cnt[0], max(prev.cnt[1], prev.cnt[2]));
            if (t && K > KS) {
                if (K <= DP_MAXK) {
                    // 8 < K <= 2048: k_dp, several columns per warp; job list by (class, depth bin), the class's unbinned
                    // list when the binned one is full
                    const int li = dp_list(K, prev.mg.n);
                    const int cls = li / DP_NBIN1;
                    const unsigned slot = atomicAdd(&ws.counters->n_pjobs[li], 1u);
                    if (slot < (unsigned)ws.pcap) {
                        ws.pjobs[((long long)cls * DP_NBIN + (li % DP_NBIN1)) * ws.pcap + slot] = (int)pc;
                    } else {
                        const unsigned s2 = atomicAdd(&ws.counters->n_pjobs[cls * DP_NBIN1 + DP_NBIN], 1u);
                        ws.ujobs[(long long)cls * ws.cap_cols + s2] = (int)pc;
                    }
                } else {
                    const unsigned slot = atomicAdd(&ws.counters->n_jobs[CLS_XL], 1u);      // one CTA per column
                    ws.jobs[(long long)CLS_XL * ws.cap_cols + slot] = (int)pc;
                }
            }
            // K <= KS: with the Bonferroni factors of a real run the early exit fires after a handful of reads (K = 1: one;
            // K = 3: ~6; K = 4: ~27 at depth-500 Q30), so almost every column ends here.  A warp runs as long as its
            // slowest lane, and the few columns with K >= 4 would keep 31 finished lanes waiting: this kernel stops after
            // PRUNE_CAP1 reads and lists what is still alive for k_prune2, whose warps are full of such columns.
            bool small = t && K <= KS;
            const double limit = small ? cf.sig * (1.0 + 1e-9) / (double)bonf : 0.0;   // margin: borderline columns go to the host
            if (cf.alt_bq_mode != 2) {          // the median override needs a warp-wide histogram: no lane-serial prune
                small = lane_prune(cf, b, s_lut, prev.mg, K, limit, PRUNE_CAP1, small, &prev.first);
                if (small) {
                    const unsigned slot = atomicAdd(&ws.counters->n_jobs[CLS_PRUNE2], 1u);
                    ws.jobs[(long long)CLS_PRUNE2 * ws.cap_cols + slot] = (int)pc;
                }
            } else if (small) {
                // every small column joins k_mid's job list: full evaluation, whole warp
                const unsigned slot = atomicAdd(&ws.counters->n_jobs[0], 1u);
                ws.jobs[slot] = (int)pc;
            }
        }
        // this round's tile becomes the previous one; its total travels through shared memory (last entry of the scan)
        prev = cur;
        prev_bal = bal;
        prev_tile = have_tile ? tile : -1;
        prev_total = have_tile ? (long long)s_warp[par][FIN_BLOCK / 32 - 1] : 0;
        __syncthreads();                                     // (3) s_excl and s_warp[par ^ 1] may be overwritten now
    }
}

// bonf_used[] for callers that ask for the dense per-column output (lfb200_device_results, lfb200_dense_out_t)
__global__ void k_bonf_used(const __grid_constant__ DevConf cf, const Workspace ws, long long n)
{
    const long long start = ws.counters->bonf_start_used;
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (long long)gridDim.x * blockDim.x) {
        const int r = ws.rank[c];
        ws.bonf_used[c] = r ? bonf_of(cf, start, r) : 0;
    }
}

void launch_front(const LaunchState &ls, const DevConf &cf, const DevBatch &b, const Lut *lut, const Workspace &ws, cudaStream_t st)
{
    if (b.n_cols <= 0) return;
    const int nb = (int)((b.n_cols + FIN_BLOCK - 1) / FIN_BLOCK);
    // per-batch state: counters (job lists, tickets), tile states of the look-back, candidate marks
    cudaMemsetAsync(ws.counters, 0, sizeof(Counters), st);
    cudaMemsetAsync(ws.blocksum, 0, (size_t)nb * sizeof(long long), st);
    cudaMemsetAsync(ws.is_cand, 0, (size_t)nb * FIN_BLOCK, st);
    const int grid = nb < ls.sms * FRONT_CTAS_PER_SM ? nb : ls.sms * FRONT_CTAS_PER_SM;     // one resident wave: the look-back needs it
    k_front<<<grid, FIN_BLOCK, 0, st>>>(cf, b, lut, ws);
}

void launch_bonf_used(const LaunchState &ls, const DevConf &cf, const Workspace &ws, long long n, cudaStream_t st)
{
    if (n <= 0) return;
    k_bonf_used<<<ls.sms * 4, 256, 0, st>>>(cf, ws, n);
}

}  // namespace lfb
