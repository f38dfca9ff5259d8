// Strand-bias quality of a reported variant (the reporting tail): report_var()
// (lofreq_call.c:108-125) runs Fisher's exact test on the 2x2 table DP4 = (ref fw, ref rv, alt fw, alt rv) with
// kt_fisher_exact (fet.c:62-101) and stores SB = PROB_TO_PHREDQUAL_SAFE(two-tailed p) (utils.h:46).
//
// The test is the classic one (Langsrud's formulation, as in samtools): the hypergeometric probability of the
// observed table, then both tails summed from the extreme tables inwards while a table is not more probable than the
// observed one, each term obtained from its neighbour by the ratio of two hypergeometric masses, re-anchored with
// lgamma at every 11th table.  One thread per table; a variant site costs up to ~depth ratio steps, which on the
// host would be tens of milliseconds per million columns and here disappears behind the next batch.
//
// SB is an integer: (int)(-10 log10l(p)).  The device decides it in double with a guard band around the integer
// boundaries; the few tables inside the band (and p so small that it may have underflowed differently) are repeated
// by the host with the same routine compiled for the host, i.e. with glibc's lgamma/exp like the reference.
#include <cuda_runtime.h>
#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include "internal.h"

namespace lfb {

struct HgState {          // the table whose mass was evaluated last
    int a, row, col, tot;
    double mass;
};

__host__ __device__ inline double ln_choose(int n, int k)
{
    if (k == 0 || n == k) return 0;
    return lgamma((double)(n + 1)) - lgamma((double)(k + 1)) - lgamma((double)(n - k + 1));
}

__host__ __device__ inline double hg_mass(int a, int row, int col, int tot)
{
    return exp(ln_choose(row, a) + ln_choose(tot - row, col - a) - ln_choose(tot, col));
}

// mass of the table with top-left cell a and the margins of st; neighbours of the last table by ratio (fet.c:38-60)
__host__ __device__ inline double hg_step(int a, HgState &st)
{
    if (a % 11 && a + st.tot - st.row - st.col) {
        if (a == st.a + 1) {
            st.mass *= (double)(st.row - st.a) / a * (st.col - st.a) / (a + st.tot - st.row - st.col);
            st.a = a;
            return st.mass;
        }
        if (a == st.a - 1) {
            st.mass *= (double)st.a / (st.row - a) * (st.a + st.tot - st.row - st.col) / (st.col - a);
            st.a = a;
            return st.mass;
        }
    }
    st.a = a;
    st.mass = hg_mass(st.a, st.row, st.col, st.tot);
    return st.mass;
}

// two-tailed p of kt_fisher_exact(n11, n12, n21, n22) (fet.c:62-101)
__host__ __device__ inline double fisher_two_tailed(int n11, int n12, int n21, int n22)
{
    HgState st;
    st.row = n11 + n12;
    st.col = n11 + n21;
    st.tot = n11 + n12 + n21 + n22;
    const int hi = st.col < st.row ? st.col : st.row;          // largest possible top-left cell
    int lo = st.row + st.col - st.tot;                          // smallest
    if (lo < 0) lo = 0;
    if (lo == hi) return 1.;
    st.a = n11;
    st.mass = hg_mass(n11, st.row, st.col, st.tot);
    const double q = st.mass;                                   // the observed table
    double p = hg_step(lo, st), left = 0., right = 0.;
    int i, j;
    for (i = lo + 1; p < 0.99999999 * q; ++i) {
        left += p;
        p = hg_step(i, st);
    }
    if (p < 1.00000001 * q) left += p;
    p = hg_step(hi, st);
    for (j = hi - 1; p < 0.99999999 * q; --j) {
        right += p;
        p = hg_step(j, st);
    }
    if (p < 1.00000001 * q) right += p;
    double two = left + right;
    if (two > 1.) two = 1.;
    return two;
}

// SB of one table; *sure = false when the host must repeat it
__host__ __device__ inline int sb_of_table(const int4 t, bool *sure)
{
    *sure = true;
    // lofreq_call.c:112-113: no reference reads and the alt reads on one strand only
    if (t.x + t.y == 0 && (t.z == 0 || t.w == 0)) return INT_MAX;
    const double two = fisher_two_tailed(t.x, t.y, t.z, t.w);
    if (!(two > 1e-290)) { *sure = false; return INT_MAX; }     // PROB_TO_PHREDQUAL_SAFE: p <= 0 -> INT_MAX; near the underflow: host
    const double qd = -10.0 * log10(two);
    const double f = qd - floor(qd);
    if (!(f > 1e-6 && f < 1.0 - 1e-6) && two < 1.0) *sure = false;
    return (int)qd;
}

__global__ void __launch_bounds__(128) k_sb_qual(const int4 *dp4, long long n, int *sb, unsigned char *unsure)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        bool sure;
        sb[i] = sb_of_table(dp4[i], &sure);
        unsure[i] = sure ? 0 : 1;
    }
}

void launch_sb_qual(int sms, const int *dp4, long long n, int *sb, unsigned char *unsure, cudaStream_t st)
{
    if (n <= 0) return;
    const long long want = (n + 127) / 128;
    const int grid = (int)(want < (long long)sms * 8 ? want : (long long)sms * 8);
    k_sb_qual<<<grid, 128, 0, st>>>(reinterpret_cast<const int4 *>(dp4), n, sb, unsure);
}

// the same routine with the host's libm and long double log10l, for the tables the device was not sure about
int sb_qual_host(const int *t4)
{
    if (t4[0] + t4[1] == 0 && (t4[2] == 0 || t4[3] == 0)) return INT_MAX;
    const double two = fisher_two_tailed(t4[0], t4[1], t4[2], t4[3]);
    if (two <= 0.0) return INT_MAX;                              // PROB_TO_PHREDQUAL_SAFE, utils.h:46
    return (int)(-10.0 * log10l(two));
}

}  // namespace lfb
