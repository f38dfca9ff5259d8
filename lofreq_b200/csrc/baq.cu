// k_kpa_glocal: the BAQ profile HMM (kpa_ext_glocal, kprobaln_ext.c:80-277) for a batch of reads, one read per thread.
//
// bam_prob_realn_core_ext (bam_md_ext.c:407) runs this HMM once per read — 80 % of the default `lofreq call` time
//  — and the reads are independent.  The arithmetic is baq_core.cuh (the reference's operations in the
// reference's order: identical state[] and q[]); what this file adds is the layout.  A read's forward matrix is
// (l_query + 1) x (6 bw + 9) doubles, far too much for registers or shared memory, and a thread walks it cell by cell: the
// scratch is interleaved over the 32 reads of a warp (KpaDevMem), so that the threads of a warp — which walk their matrices
// in the same order, read lengths and bands allowing — touch 32 consecutive doubles per access.  Reads are handed out in
// launch-sized chunks sorted by nothing: a warp runs as long as its longest read.
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "internal.h"
#include "baq_core.cuh"

namespace lfb {

#ifndef KPA_CTAS
#define KPA_CTAS 6
#endif

// Scratch of a launch: every warp owns one contiguous block, its 32 reads interleaved inside it — cell c of row i of lane l
// at ((w * rows + i) * W3 + c) * 32 + l.  The lanes of a warp touch 32 consecutive doubles per access, and a warp walks its
// own few megabytes front to back instead of striding through the scratch of the whole launch.
struct KpaDevMem {
    double *f, *b, *s;
    size_t w, l;
    int w3, rows;
    __device__ __forceinline__ double &F(int i, int c) { return f[(((w * rows + i) * w3 + c) << 5) + l]; }
    __device__ __forceinline__ double &B(int p, int c) { return b[(((w * 2 + p) * w3 + c) << 5) + l]; }
    __device__ __forceinline__ double &S(int i) { return s[((w * (rows + 1) + i) << 5) + l]; }
};

__global__ void __launch_bounds__(128, KPA_CTAS) k_kpa_glocal(long long r0, int n_reads, const unsigned char *ref, const long long *ref_off,
                                                    const unsigned char *query, const long long *qry_off, const unsigned char *qual, float d,
                                                    float e, int bw, const float *q2p, double *f, double *b, double *s, int w3, int rows, int *state,
                                                    unsigned char *q, KpaFix *fix, int fix_cap, unsigned *n_fix)
{
    __shared__ float s_q2p[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_q2p[i] = q2p[i];
    __syncthreads();
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_reads) return;
    const long long r = r0 + t;
    KpaDevMem mem;
    mem.f = f; mem.b = b; mem.s = s;
    mem.w = (size_t)(t >> 5);
    mem.l = (size_t)(t & 31);
    mem.w3 = w3;
    mem.rows = rows;
    const long long ro = ref_off[r], qo = qry_off[r];
    kpa_glocal_core(ref + ro, (int)(ref_off[r + 1] - ro), query + qo, (int)(qry_off[r + 1] - qo), qual ? qual + qo : nullptr, d, e, bw, s_q2p,
                    mem, state + qo, q + qo, qo, fix, fix_cap, n_fix);
}

void launch_kpa_glocal(long long r0, int n_reads, const unsigned char *ref, const long long *ref_off, const unsigned char *query,
                       const long long *qry_off, const unsigned char *qual, float d, float e, int bw, const float *q2p, double *f, double *b,
                       double *s, int w3, int rows, int *state, unsigned char *q, KpaFix *fix, int fix_cap, unsigned *n_fix, cudaStream_t st)
{
    if (n_reads <= 0) return;
    k_kpa_glocal<<<(n_reads + 127) / 128, 128, 0, st>>>(r0, n_reads, ref, ref_off, query, qry_off, qual, d, e, bw, q2p, f, b, s, w3, rows, state, q,
                                                        fix, fix_cap, n_fix);
}

}  // namespace lfb
