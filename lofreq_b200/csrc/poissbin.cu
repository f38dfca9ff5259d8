// The two low-volume entry points of snpcaller.h that hand arrays back to the caller:
//   poissbin()        (snpcaller.c:1019-1062): the row of natural-log probabilities pruned_calc_prob_dist
//                     (snpcaller.c:830-971) ends with — possibly a "partial" row, when the Bonferroni-aware early
//                     exit fired.  source_qual (plp.c:554-559) reads row[num_failures-1] of it.
//   plp_to_errprobs() (snpcaller.c:345-498): the merged error probabilities of one column, in pileup order.
// Unlike the batched path (which only needs tails and may evaluate in any order and in linear space), a caller of
// poissbin() sees every cell of the row, including cells hundreds of decades below the largest one, and the read at
// which the early exit fired.  So this kernel walks the reads in the caller's order with the reference's log-space
// recurrence, one CTA per problem, cells strided over the threads.
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "internal.h"
#include "dev_common.cuh"

namespace lfb {

static constexpr double LOGZERO = -1e100;              // snpcaller.c:66

__device__ __forceinline__ double log_sum_d(double a, double b)      // snpcaller.c:692-700
{
    return a > b ? a + log1p(exp(b - a)) : b + log1p(exp(a - b));
}

constexpr int PB_T = 128;

// buf: 2*(K+1) doubles per problem at 2*row_off[i]; rows: K+1 doubles at row_off[i]
__global__ void __launch_bounds__(PB_T) k_poissbin_rows(const ProbBatch pb, const int *num_failures, const long long *row_off,
                                                        double *buf, double *rows, int *n_end)
{
    __shared__ int s_pruned;
    for (long long i = blockIdx.x; i < pb.n; i += gridDim.x) {
        const double *ep = pb.err_probs + pb.ep_off[i];
        const int N = (int)(pb.ep_off[i + 1] - pb.ep_off[i]);
        const int K = num_failures[i];
        const double bonf = (double)pb.bonf[i];
        double *prev = buf + 2 * row_off[i], *cur = prev + (K + 1);
        for (int k = threadIdx.x; k <= K; k += PB_T) {
            prev[k] = k == 0 ? 0.0 : LOGZERO;          // the reference sets prev[n] = LOGZERO at step n < K (snpcaller.c:888-890)
            cur[k] = LOGZERO;
        }
        if (threadIdx.x == 0) s_pruned = 0;
        __syncthreads();
        int stop = N;
        for (int n = 1; n <= N; ++n) {
            const double pn = ep[n - 1];
            const double lp = fabs(pn) < DEPS ? log(DEPS) : log(pn);                          // snpcaller.c:872-881
            const double lq = fabs(pn - 1.0) < DEPS ? log1p(-pn + DEPS) : log1p(-pn);
            const int kmax = min(n, K - 1);
            for (int k = 1 + threadIdx.x; k <= kmax; k += PB_T) cur[k] = log_sum_d(prev[k] + lq, prev[k - 1] + lp);
            if (threadIdx.x == 0) {
                cur[0] = prev[0] + lq;
                if (n == K) {
                    cur[K] = prev[K - 1] + lp;
                } else if (n > K) {
                    const double t = log_sum_d(prev[K], prev[K - 1] + lp);
                    cur[K] = t;
                    if (exp(t) * bonf > pb.sig) s_pruned = 1;                                   // snpcaller.c:916-958
                }
            }
            __syncthreads();
            double *sw = prev; prev = cur; cur = sw;                                           // prev = the row of step n
            if (s_pruned) { stop = n; break; }
        }
        for (int k = threadIdx.x; k <= K; k += PB_T) rows[row_off[i] + k] = prev[k];
        if (threadIdx.x == 0) n_end[i] = stop;
        __syncthreads();
    }
}

void launch_poissbin_rows(const ProbBatch &pb, const int *num_failures, const long long *row_off, double *buf, double *rows,
                          int *n_end, cudaStream_t st)
{
    if (pb.n <= 0) return;
    const int grid = (int)(pb.n < 148 * 8 ? pb.n : 148 * 8);
    k_poissbin_rows<<<grid, PB_T, 0, st>>>(pb, num_failures, row_off, buf, rows, n_end);
}

// plp_to_errprobs for the columns of a batch, one warp per column, reads kept in pileup order (ballot + prefix count).
// ep_out: the kept probabilities of column c start at ep_out[col_off[c]]; counts9: [n][9] = alt_bases, alt_counts, alt_raw_counts
__global__ void __launch_bounds__(128) k_errprobs(const __grid_constant__ DevConf cf, const __grid_constant__ DevBatch b, const Lut *lut,
                                                  double *ep_out, int *n_out, int *counts9)
{
    __shared__ double s_lut[768];
    __shared__ int s_hist[4][256];
    load_lut(s_lut, lut);
    const int lane = lane_id(), wib = threadIdx.x >> 5;
    for (long long c = (long long)blockIdx.x * 4 + wib; c < b.n_cols; c += (long long)gridDim.x * 4) {
        Geom g;
        int cov;
        load_geom(b, c, g, cov);
        setup_alt_bq(cf, b, s_lut, g, s_hist[wib]);
        int kept = 0, ac0 = 0, ac1 = 0, ac2 = 0;
        for (int p0 = 0; p0 < g.n; p0 += 32) {
            const int pos = p0 + lane;
            bool ok = false, is_alt = false;
            int slot = 0;
            double jp = 0.0;
            if (pos < g.n) {
                const long long a = g.off + pos;
                ok = eval_read<true>(cf, s_lut, g, pos, b.bq[a], cf.use_mq ? b.mq[a] : 0, cf.use_baq ? b.baq[a] : 0,
                                     cf.use_sq ? b.sq[a] : 0, is_alt, slot, jp);
            }
            const unsigned m = __ballot_sync(FULL, ok);
            if (ok) ep_out[g.off + kept + __popc(m & ((1u << lane) - 1u))] = jp;
            kept += __popc(m);
            ac0 += __popc(__ballot_sync(FULL, ok && is_alt && slot == 0));
            ac1 += __popc(__ballot_sync(FULL, ok && is_alt && slot == 1));
            ac2 += __popc(__ballot_sync(FULL, ok && is_alt && slot == 2));
        }
        if (lane == 0) {
            n_out[c] = kept;
            int *o = counts9 + 9 * c;
            const int sizes[4] = {g.b1, g.b2 - g.b1, g.b3 - g.b2, g.n - g.b3};
            int k = 0;
            for (int nt = 0; nt < 4; ++nt) {           // alt slots are A,C,G,T minus the reference base (snpcaller.c:383-397)
                if (nt == g.ref_idx || k >= 3) continue;
                o[k] = "ACGT"[nt];
                o[6 + k] = sizes[nt];                  // raw count: every read showing the base, before any filter (snpcaller.c:418-420)
                ++k;
            }
            o[3] = ac0; o[4] = ac1; o[5] = ac2;
        }
    }
}

void launch_errprobs(const DevConf &cf, const DevBatch &b, const Lut *lut, double *ep_out, int *n_out, int *counts9, cudaStream_t st)
{
    if (b.n_cols <= 0) return;
    const long long want = (b.n_cols + 3) / 4;
    k_errprobs<<<(int)(want < 148 * 8 ? want : 148 * 8), 128, 0, st>>>(cf, b, lut, ep_out, n_out, counts9);
}

}  // namespace lfb
