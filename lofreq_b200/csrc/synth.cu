// Synthetic pileup columns generated on the device (benchmark input).
// Integer-only, stateless in (seed, column, read), so oracle/synth_np.py reproduces every byte on the
// host (tests/test_parity_gpu.py::test_synth_matches_numpy).  One warp per column: classify each read,
// then place it stably into its A/C/G/T group (the layout of plp_col_t.base_quals[], plp.h:89-92).
#include <cuda_runtime.h>
#include <stdint.h>
#include "internal.h"
#include "synth_tables.h"

namespace lfb {

#define FULL 0xffffffffu

__constant__ unsigned int c_err_thr[256];
__constant__ unsigned int c_af_q20[1024];
__constant__ unsigned int c_depth_tab[1024];
static bool g_tables_up = false;

static void upload_tables()
{
    if (g_tables_up) return;
    cudaMemcpyToSymbol(c_err_thr, k_err_thr, sizeof(k_err_thr));
    cudaMemcpyToSymbol(c_af_q20, k_af_q20, sizeof(k_af_q20));
    cudaMemcpyToSymbol(c_depth_tab, k_depth_tab, sizeof(k_depth_tab));
    g_tables_up = true;
}

__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x)
{
    unsigned long long z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__device__ __forceinline__ unsigned long long col_hash(unsigned long long c, unsigned long long salt)
{
    return splitmix64(LFB_SYNTH_SEED ^ (c << 20) ^ (salt << 56));
}

__device__ __forceinline__ unsigned long long read_hash(unsigned long long c, unsigned long long r, unsigned long long salt)
{
    return splitmix64(LFB_SYNTH_SEED ^ (c << 20) ^ r ^ (salt << 56));
}

// workload 2..5: fixed depth 500 / 2000 / 300, or log-uniform 50..10000
__device__ __forceinline__ int depth_of(int workload, unsigned long long c)
{
    if (workload == 2) return 500;
    if (workload == 3) return 2000;
    if (workload == 4) return 300;
    return (int)c_depth_tab[col_hash(c, 3) >> 54];
}

__global__ void k_synth_depths(int workload, long long c0, long long n, int *depth)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) depth[i] = depth_of(workload, (unsigned long long)(c0 + i));
}

struct SynRead {
    int nt, bq, baq;
};

__device__ __forceinline__ SynRead make_read(int workload, unsigned long long c, int r, int n_alt)
{
    SynRead s;
    const unsigned long long hq = read_hash(c, (unsigned long long)r, 0);
    const unsigned long long he = read_hash(c, (unsigned long long)r, 1);
    const int u = 20 + (int)((hq & 0xFFFFull) % 21ull);
    if (workload == 2) s.bq = 30;
    else if (workload == 4) s.bq = (((hq >> 16) & 0xFFFFull) < 45875ull) ? 30 : u;
    else s.bq = u;
    s.baq = 30 + (int)(((hq >> 32) & 0xFFFFull) % 31ull);
    const int ref_nt = (int)(c & 3ull);
    const bool err = (unsigned int)(he >> 32) < c_err_thr[s.bq];
    const int err_nt = (ref_nt + 1 + (int)((he & 0xFFFFull) % 3ull)) & 3;
    s.nt = err ? err_nt : ref_nt;
    if (r < n_alt) s.nt = (int)((c + 1ull) & 3ull);
    return s;
}

__global__ void __launch_bounds__(256) k_synth_columns(int workload, long long c0, long long n, const long long *col_off,
                                                       int *nt_cnt, char *ref, unsigned char *bq, unsigned char *mq,
                                                       unsigned char *baq)
{
    const int lane = threadIdx.x & 31;
    const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    const unsigned lt_mask = (1u << lane) - 1u;
    for (long long i = warp0; i < n; i += nwarps) {
        const unsigned long long c = (unsigned long long)(c0 + i);
        const int d = depth_of(workload, c);
        const unsigned long long hv = col_hash(c, 2);
        const bool is_var = (hv % 100ull) == 0ull;
        const long long af = (long long)c_af_q20[(hv >> 40) & 1023ull];
        const int n_alt = is_var ? (int)((af * (long long)d + (1ll << 19)) >> 20) : 0;
        // pass 1: group sizes
        int cnt[4] = {0, 0, 0, 0};
        for (int r0 = 0; r0 < d; r0 += 32) {
            const int r = r0 + lane;
            const int nt = r < d ? make_read(workload, c, r, n_alt).nt : -1;
#pragma unroll
            for (int g = 0; g < 4; ++g) cnt[g] += __popc(__ballot_sync(FULL, nt == g));
        }
        // pass 2: stable placement
        int base[4];
        base[0] = 0;
        base[1] = cnt[0];
        base[2] = base[1] + cnt[1];
        base[3] = base[2] + cnt[2];
        const long long off = col_off[i];
        const long long pitch = col_off[i + 1] - off;
        for (int r0 = 0; r0 < d; r0 += 32) {
            const int r = r0 + lane;
            SynRead s;
            s.nt = -1;
            if (r < d) s = make_read(workload, c, r, n_alt);
            int dst = -1;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const unsigned m = __ballot_sync(FULL, s.nt == g);
                if (s.nt == g) dst = base[g] + __popc(m & lt_mask);
                base[g] += __popc(m);
            }
            if (dst >= 0) {
                bq[off + dst] = (unsigned char)s.bq;
                mq[off + dst] = 60;
                if (baq) baq[off + dst] = (unsigned char)s.baq;
            }
        }
        // padding bytes are zero, like the numpy generator
        for (long long p = d + lane; p < pitch; p += 32) {
            bq[off + p] = 0;
            mq[off + p] = 0;
            if (baq) baq[off + p] = 0;
        }
        if (lane < 4) nt_cnt[4 * i + lane] = cnt[lane];
        if (lane == 0) ref[i] = "ACGT"[c & 3ull];
    }
}

void launch_synth_depths(int workload, long long c0, long long n, int *depth, cudaStream_t st)
{
    upload_tables();
    if (n <= 0) return;
    k_synth_depths<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(workload, c0, n, depth);
}

void launch_synth_columns(int workload, long long c0, long long n, const long long *col_off, int *nt_cnt, char *ref,
                          unsigned char *bq, unsigned char *mq, unsigned char *baq, cudaStream_t st)
{
    upload_tables();
    if (n <= 0) return;
    const long long want = (n + 7) / 8;
    const int grid = (int)(want < 148ll * 8 ? want : 148ll * 8);
    k_synth_columns<<<grid, 256, 0, st>>>(workload, c0, n, col_off, nt_cnt, ref, bq, mq, baq);
}

}  // namespace lfb
