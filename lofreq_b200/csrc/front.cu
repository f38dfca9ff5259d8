// k_front: the stream side of the test in one pass over the batch.
//
// Per column, what call_vars()/call_snvs() do before snpcaller() (lofreq_call.c:886-935, 734-801): the gates, the
// filtered and raw alt counts of plp_to_errprobs (only the reads showing a non-reference base decide them,
// snpcaller.c:418-420,489), "is this column a test", its place in the running Bonferroni count
// (lofreq_call.c:794-800) — and then at once the reference's early exit (snpcaller.c:916-958), one lane per column,
// on the first reads of the column: with a real Bonferroni factor nearly every tested column is ruled out after a
// handful of reads, without a warp ever being dedicated to it.
//
// One CTA per tile of 256 consecutive columns, tiles taken in order (a ticket), thread per column:
//   A. metadata, gates, alt counts — lane per column for up to 8 non-reference reads, the whole warp for the rare
//      columns with more (variant sites);
//   B. the running count: tested columns of the tile -> decoupled look-back over the tiles before it (single-pass
//      prefix sum: a tile publishes its aggregate as soon as it has counted, then its inclusive prefix), so every
//      tested column knows its 1-based rank among the tested columns of the batch;
//   C. factor = start + 3 * rank; K > 8 -> job list of k_dp / k_heavy_xl; K <= 8 -> prune over the first PRUNE_CAP1
//      reads, survivors to k_prune2's list.
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

__global__ void __launch_bounds__(FIN_BLOCK, 4) k_front(const __grid_constant__ DevConf cf, const __grid_constant__ DevBatch b,
                                                        const Lut *lut, const Workspace ws)
{
    __shared__ double s_lut[768];
    __shared__ int s_hist[FIN_BLOCK / 32][256];
    __shared__ int s_warp[32];
    __shared__ long long s_excl;
    __shared__ unsigned s_tile;
    if (threadIdx.x == 0) s_tile = atomicAdd(&ws.counters->front_ticket, 1u);
    load_lut(s_lut, lut);            // (ends with the barrier that also publishes the ticket)
    const long long tile = s_tile;
    const long long n = b.n_cols;
    const long long c = tile * FIN_BLOCK + threadIdx.x;
    const int lane = lane_id(), w = threadIdx.x >> 5;
    if (tile == 0 && threadIdx.x == 0) ws.counters->bonf_start_used = cf.bonf_start;

    // ---- A. metadata, gates, alt counts ----
    // the median override (def_alt_bq == -1) needs a warp-wide histogram: no lane-per-column path then
    const int serial_max = cf.alt_bq_mode == 2 ? 0 : 8;
    RawGeom cur;
    cur.off = 0; cur.cnt = make_int4(0, 0, 0, 0); cur.cov = -1; cur.nb = -1; cur.ref = 'N';
    if (c < n) load_raw(b, c, cur);
    Geom mg;
    mg.off = cur.off;
    mg.b1 = cur.cnt.x;
    mg.b2 = mg.b1 + cur.cnt.y;
    mg.b3 = mg.b2 + cur.cnt.z;
    mg.n = mg.b3 + cur.cnt.w;
    mg.ref_idx = ref_index(cur.ref);
    mg.alt_bp = cf.alt_bq_prob;
    const int m_cov = cur.cov < 0 ? mg.n : cur.cov;
    const int m_nb = cur.nb < 0 ? mg.n : cur.nb;           // plp_col_t.num_bases
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
    const int t = (c < n && (cnt[0] | cnt[1] | cnt[2]) != 0) ? 1 : 0;
    if (c < n) {
        int2 *o = reinterpret_cast<int2 *>(ws.cnt6 + 6 * c);
        o[0] = make_int2(cnt[0], cnt[1]);
        o[1] = make_int2(cnt[2], raw[0]);
        o[2] = make_int2(raw[1], raw[2]);
        ws.tested[c] = (unsigned char)t;
    }

    // the first 16 reads of a column the prune will walk: requested now, used after the look-back
    const int K = max(cnt[0], max(cnt[1], cnt[2]));
    Chunk16 first;
    first.bq = first.mq = first.baq = first.sq = make_uint4(0, 0, 0, 0);
    if (t && K <= KS && cf.alt_bq_mode != 2 && mg.n > 0) load_chunk(cf, b, mg.off & ~15ll, first);

    // ---- B. rank among the tested columns of the batch ----
    const unsigned bal = __ballot_sync(FULL, t);
    if (lane == 0) s_warp[w] = __popc(bal);
    __syncthreads();
    if (w == 0) {
        int z = lane < FIN_BLOCK / 32 ? s_warp[lane] : 0;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int y = __shfl_up_sync(FULL, z, d);
            if (lane >= d) z += y;
        }
        const long long total = __shfl_sync(FULL, z, 31);
        unsigned long long *state = reinterpret_cast<unsigned long long *>(ws.blocksum);
        long long excl = 0;
        if (tile > 0) {
            if (lane == 0) st_release_gpu(&state[tile], TS_AGG | (unsigned long long)total);
            // Look back over the tiles before this one, 256 at a time, nearest first: every lane takes 8 consecutive
            // tiles (8 loads in flight), so a whole resident wave of tiles is covered in two or three L2 round trips.
            long long base = tile - 1;
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
        }
        if (lane == 0) {
            st_release_gpu(&state[tile], TS_PREFIX | (unsigned long long)(excl + total));
            s_excl = excl;
            if ((tile + 1) * FIN_BLOCK >= n) ws.counters->n_tested = (unsigned long long)(excl + total);
        }
        s_warp[lane] = z;
    }
    __syncthreads();
    int rank = 0;
    long long bonf = 0;
    if (t) {
        // 1-based rank of this column among the tested columns of the batch
        rank = (int)(s_excl + (w ? s_warp[w - 1] : 0) + __popc(bal & ((2u << lane) - 1u)));
        bonf = bonf_of(cf, cf.bonf_start, rank);
    }
    if (c < n) ws.rank[c] = rank;

    // ---- C. routing and the first stage of the prune ----
    if (t && K > KS) {
        if (K <= DP_MAXK) {
            // 8 < K <= 2048: k_dp, several columns per warp; job list by (class, depth bin), the class's unbinned list
            // when the binned one is full
            const int li = dp_list(K, mg.n);
            const int cls = li / DP_NBIN1;
            const unsigned slot = atomicAdd(&ws.counters->n_pjobs[li], 1u);
            if (slot < (unsigned)ws.pcap) {
                ws.pjobs[((long long)cls * DP_NBIN + (li % DP_NBIN1)) * ws.pcap + slot] = (int)c;
            } else {
                const unsigned s2 = atomicAdd(&ws.counters->n_pjobs[cls * DP_NBIN1 + DP_NBIN], 1u);
                ws.ujobs[(long long)cls * ws.cap_cols + s2] = (int)c;
            }
        } else {
            const unsigned slot = atomicAdd(&ws.counters->n_jobs[CLS_XL], 1u);      // one CTA per column
            ws.jobs[(long long)CLS_XL * ws.cap_cols + slot] = (int)c;
        }
    }
    // K <= KS: with the Bonferroni factors of a real run the early exit fires after a handful of reads (K = 1: one;
    // K = 3: ~6; K = 4: ~27 at depth-500 Q30), so almost every column ends here.  A warp runs as long as its slowest
    // lane, and the few columns with K >= 4 would keep 31 finished lanes waiting: this kernel stops after PRUNE_CAP1
    // reads and lists what is still alive for k_prune2, whose warps are full of such columns.
    bool small = t && K <= KS;
    const double limit = small ? cf.sig * (1.0 + 1e-9) / (double)bonf : 0.0;   // margin: borderline columns go to the host
    if (cf.alt_bq_mode != 2) {          // the median override needs a warp-wide histogram: no lane-serial prune
        small = lane_prune(cf, b, s_lut, mg, K, limit, PRUNE_CAP1, small, &first);
        if (small) {
            const unsigned slot = atomicAdd(&ws.counters->n_jobs[CLS_PRUNE2], 1u);
            ws.jobs[(long long)CLS_PRUNE2 * ws.cap_cols + slot] = (int)c;
        }
    } else if (small) {
        // every small column joins k_mid's job list: full evaluation, whole warp
        const unsigned slot = atomicAdd(&ws.counters->n_jobs[0], 1u);
        ws.jobs[slot] = (int)c;
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
    k_front<<<nb, FIN_BLOCK, 0, st>>>(cf, b, lut, ws);
}

void launch_bonf_used(const LaunchState &ls, const DevConf &cf, const Workspace &ws, long long n, cudaStream_t st)
{
    if (n <= 0) return;
    k_bonf_used<<<ls.sms * 4, 256, 0, st>>>(cf, ws, n);
}

}  // namespace lfb
