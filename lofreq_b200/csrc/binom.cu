// k_binom: batched binom() — the binomial CDF / survival function behind `lofreq uniq`'s frequency test
// (binom.c:52-93 -> cdfbin(which=1), dcdflib.c:1727 -> cumbin :4966 -> cumbet -> bratio, TOMS 708).
//
// cdflib reduces the binomial to the incomplete beta function.  The device sums the probability mass function
// on the short side of the mode instead: one accurately computed term — Loader's saddle-point form, i.e. the
// Stirling-series error term stirlerr() plus the deviance bd0() — and from there the exact term ratio
// pmf(k-1)/pmf(k) = k/(n-k+1) * q/p (left of the mode) or pmf(k+1)/pmf(k) = (n-k)/(k+1) * p/q (right of it).
// Every term is positive and the terms shrink, so the relative error of the tail is that of the first term
// (~1e-15 * |ln pmf|); the other output is 1 - tail with tail <= ~1/2.  One thread per problem: the volume is one
// call per input variant (lofreq_uniq.c:381).  Status codes are cdfbin's (dcdflib.c:1860-1960).
// tests/algomodel.py:model_binom is the same arithmetic in Python, checked on the CPU against the compiled
// reference; tests/test_parity_gpu.py checks this kernel against it on the GPU.
#include <cuda_runtime.h>
#include <math.h>
#include "internal.h"

namespace lfb {

__constant__ double c_stirlerr[16];      // stirlerr(0..15), built by the host in long double

__device__ __forceinline__ double stirlerr(int n)
{
    if (n <= 15) return c_stirlerr[n];
    const double S0 = 1.0 / 12, S1 = 1.0 / 360, S2 = 1.0 / 1260, S3 = 1.0 / 1680, S4 = 1.0 / 1188;
    const double x = (double)n, nn = x * x;
    if (n > 500) return (S0 - S1 / nn) / x;
    if (n > 80) return (S0 - (S1 - S2 / nn) / nn) / x;
    if (n > 35) return (S0 - (S1 - (S2 - S3 / nn) / nn) / nn) / x;
    return (S0 - (S1 - (S2 - (S3 - S4 / nn) / nn) / nn) / nn) / x;
}

// x ln(x/np) + np - x without cancellation when x is close to np
__device__ double bd0(double x, double np)
{
    if (fabs(x - np) < 0.1 * (x + np)) {
        double v = (x - np) / (x + np);
        double s = (x - np) * v;
        double ej = 2.0 * x * v;
        v = v * v;
        for (int j = 1; j < 1000; ++j) {
            ej *= v;
            const double s1 = s + ej / (double)(2 * j + 1);
            if (s1 == s) return s1;
            s = s1;
        }
    }
    return x * log(x / np) + np - x;
}

// ln of the binomial mass at x, 0 < p < 1, q = 1 - p
__device__ double ln_binom_pmf(int x, int n, double p, double q)
{
    const double dn = (double)n, dx = (double)x;
    if (x == 0) return p < 0.1 ? -bd0(dn, dn * q) - dn * p : dn * log(q);
    if (x == n) return q < 0.1 ? -bd0(dn, dn * p) - dn * q : dn * log(p);
    const double lc = stirlerr(n) - stirlerr(x) - stirlerr(n - x) - bd0(dx, dn * p) - bd0(dn - dx, dn * q);
    const double lf = 1.8378770664093454835606594728112 + log(dx) + log1p(-dx / dn);
    return lc - 0.5 * lf;
}

__global__ void __launch_bounds__(128) k_binom(long long n_prob, const int *num_trials, const int *num_success, const double *prob,
                                               double *cum_out, double *ccum_out, int *status_out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_prob) return;
    const int n = num_trials[i], s = num_success[i];
    const double pr = prob[i];
    // argument checks in cdfbin's order (which = 1): XN, S, PR
    int status = 0;
    if (!(n > 0)) status = -5;
    else if (s < 0 || s > n) status = -4;
    else if (pr < 0.0 || pr > 1.0) status = -6;
    status_out[i] = status;
    if (status) return;                  // like the reference, outputs stay untouched
    double cum, ccum;
    const double q = 1.0 - pr;
    if (!(s < n) || pr == 0.0) {         // cumbin, dcdflib.c:5017-5026; pr == 0: all mass at 0
        cum = 1.0;
        ccum = 0.0;
    } else if (q == 0.0) {               // all mass at n > s
        cum = 0.0;
        ccum = 1.0;
    } else {
        const double mode = floor(((double)n + 1.0) * pr);
        if ((double)s < mode) {
            const double lt = ln_binom_pmf(s, n, pr, q);
            const double ratio = q / pr;
            double t = 1.0, tot = 1.0;
            for (int k = s; k > 0; --k) {
                t *= ((double)k / ((double)(n - k) + 1.0)) * ratio;
                tot += t;
                if (t < 1e-18 * tot) break;
            }
            cum = exp(lt + log(tot));
            ccum = 1.0 - cum;
        } else {
            const double lt = ln_binom_pmf(s + 1, n, pr, q);
            const double ratio = pr / q;
            double t = 1.0, tot = 1.0;
            for (int k = s + 1; k < n; ++k) {
                t *= ((double)(n - k) / ((double)k + 1.0)) * ratio;
                tot += t;
                if (t < 1e-18 * tot) break;
            }
            ccum = exp(lt + log(tot));
            cum = 1.0 - ccum;
        }
    }
    cum_out[i] = cum;
    ccum_out[i] = ccum;
}

int launch_binom(long long n_prob, const int *num_trials, const int *num_success, const double *prob, double *cum, double *ccum,
                 int *status, cudaStream_t st)
{
    static unsigned long long ready_mask = 0;       // __constant__ memory is per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (!(ready_mask >> (dev & 63) & 1ull)) {
        // stirlerr(n) = ln n! - ln(sqrt(2 pi n) (n/e)^n): heavy cancellation, hence long double
        double tab[16];
        tab[0] = 0.0;
        long double fact = 1.0L;
        for (int k = 1; k < 16; ++k) {
            fact *= (long double)k;
            const long double v = logl(fact) - ((long double)k + 0.5L) * logl((long double)k) + (long double)k -
                                  0.5L * logl(2.0L * 3.14159265358979323846264338327950288L);
            tab[k] = (double)v;
        }
        if (cudaMemcpyToSymbol(c_stirlerr, tab, sizeof(tab)) != cudaSuccess) return 1;
        ready_mask |= 1ull << (dev & 63);
    }
    if (n_prob <= 0) return 0;
    const int grid = (int)((n_prob + 127) / 128);
    k_binom<<<grid, 128, 0, st>>>(n_prob, num_trials, num_success, prob, cum, ccum, status);
    return 0;
}

}  // namespace lfb
