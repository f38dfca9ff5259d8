/* TEST INFRASTRUCTURE — not part of the shipped product.
 *
 * CPU restatement ("port") of LoFreq's per-pileup-column SNV test, written from
 * the behaviour of the reference (file:line cited per function, relative to
 * /root/reference/src/lofreq/).  It exists to check the CUDA path and to serve
 * as a CPU baseline when oracle/_ref is not available; the product never calls
 * it.  Parity is PINNED: tests/test_oracle.py compares every function below
 * bit-for-bit with the compiled, unmodified reference (oracle/_ref/libsnpref.so)
 * and with the fixtures under tests/golden/ that were generated from it
 * (tests/golden/make_golden.py), including the one known-answer value the
 * reference carries in a comment (snpcaller.c:1222-1232).
 *
 * It uses the same libm entry points as the reference (pow, log, log1p, exp,
 * expl, log10l) in the same order, so on the same glibc the results are
 * bit-identical, including which calls raise FE_UNDERFLOW.
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <float.h>
#include <limits.h>
#include <errno.h>
#include <fenv.h>

#include "column_batch.h"

#define LFO_USE_BAQ 1   /* defaults.h:76 */
#define LFO_USE_MQ  2   /* defaults.h:77 */
#define LFO_USE_SQ  4   /* defaults.h:78 */
#define LFO_LOGZERO (-1e100)      /* snpcaller.c:66 */
#define LFO_MQ0_ERRPROB 0.5       /* snpcaller.c:64 */

/* ---- A1: quality <-> probability (utils.h:42-46) ----------------------- */
double lfo_phred_to_prob(int q)
{
    if (q == INT_MAX) return DBL_MIN;
    return pow(10.0, -1.0 * q / 10.0);
}

int lfo_prob_to_phred(long double p)
{
    return (int)(-10.0 * log10l(p));
}

int lfo_prob_to_phred_safe(double p)
{
    if (p <= 0.0) return INT_MAX;
    return (int)(-10.0 * log10l(p));
}

/* ---- A2: four-way quality merge (snpcaller.c:302-341) ------------------- */
double lfo_merge_quals(int sq, int mq, int baq, int bq)
{
    double p_src = (sq == -1) ? 0.0 : lfo_phred_to_prob(sq);
    double p_map = (mq == -1) ? 0.0 : (mq == 0 ? LFO_MQ0_ERRPROB : lfo_phred_to_prob(mq));
    double p_aln = (baq == -1) ? 0.0 : lfo_phred_to_prob(baq);
    double p_base = (bq == -1) ? 0.0 : lfo_phred_to_prob(bq);
    /* association order of snpcaller.c:334, left to right, no contraction */
    double ok_map = 1.0 - p_map;
    double acc = p_map + ok_map * p_src;
    acc = acc + ok_map * (1 - p_src) * p_aln;
    acc = acc + ok_map * (1 - p_src) * (1 - p_aln) * p_base;
    return acc;
}

/* ---- median of an int vector, even size -> truncated mean (utils.c:435-458) */
static int cmp_int(const void *a, const void *b)
{
    int x = *(const int *)a, y = *(const int *)b;
    return (x > y) - (x < y);
}

int lfo_int_median(const int *v, int n)
{
    int *tmp, r;
    if (n == 0) return 0;
    tmp = malloc(sizeof(int) * n);
    memcpy(tmp, v, sizeof(int) * n);
    qsort(tmp, n, sizeof(int), cmp_int);
    r = (n % 2) ? tmp[n / 2] : (int)((tmp[n / 2] + tmp[n / 2 - 1]) / 2.0);
    free(tmp);
    return r;
}

/* ---- A3: error-probability vector of one packed column (snpcaller.c:345-498) */
static const char k_bases[4] = {'A', 'C', 'G', 'T'};

int lfo_column_errprobs(const oracle_conf_t *cf, const oracle_batch_t *b, long long c,
                        double *ep, int *alt_bases, int *alt_counts, int *alt_raw)
{
    long long base_idx = b->col_off[c];
    int n = 0, slot = -1, g;
    int ref_median = -1;
    char ref = b->ref_base[c];
    int use_baq = (cf->flag & LFO_USE_BAQ) && b->baq;
    int use_mq = (cf->flag & LFO_USE_MQ) && b->mq;
    int use_sq = (cf->flag & LFO_USE_SQ) && b->sq;

    if (cf->def_alt_bq == -1) {          /* snpcaller.c:363-379 */
        long long o = base_idx;
        for (g = 0; g < 4; g++) {
            int cnt = b->nt_cnt[4 * c + g];
            if (k_bases[g] == ref && cnt) {
                int *q = malloc(sizeof(int) * cnt), j;
                for (j = 0; j < cnt; j++) q[j] = b->bq[o + j];
                ref_median = lfo_int_median(q, cnt);
                free(q);
                break;
            }
            o += cnt;
        }
    }

    for (g = 0; g < 4; g++) {
        int cnt = b->nt_cnt[4 * c + g], j;
        int alt = (k_bases[g] != ref);
        if (alt) {
            slot++;
            alt_bases[slot] = k_bases[g];
            alt_counts[slot] = 0;
            alt_raw[slot] = 0;
        }
        /* a group without base qualities contributes nothing: the reference
         * loops j < base_quals[i].n (snpcaller.c:399) */
        for (j = 0; j < cnt; j++) {
            long long at = base_idx + j;
            int bq = b->bq[at], mq = -1, baq = -1, sq = -1, jq;
            double jp;
            if (alt) alt_raw[slot]++;                       /* raw count precedes all filters :418-420 */
            if (bq < cf->min_bq) continue;                  /* :426 */
            if (alt) {
                if (bq < cf->min_alt_bq) continue;          /* :433 */
                if (cf->def_alt_bq == -1) bq = ref_median;  /* :435 */
                else if (cf->def_alt_bq != 0) bq = cf->def_alt_bq;
            }
            if (use_baq) { baq = b->baq[at]; if (baq == 255) baq = -1; }
            if (use_mq)  { mq = b->mq[at];   if (mq == 255) mq = -1; }       /* :451-453 */
            if (use_sq)  { sq = b->sq[at];   if (sq == 255) sq = -1; }
            jp = lfo_merge_quals(sq, mq, baq, bq);
            jq = lfo_prob_to_phred_safe(jp);
            if (jq < cf->min_jq) continue;                  /* :469 */
            if (alt) {
                if (jq < cf->min_alt_jq) continue;          /* :480 */
                if (cf->def_alt_jq == -1) abort();          /* reference exits: not implemented (:482-484) */
                if (cf->def_alt_jq != 0) jp = lfo_phred_to_prob(cf->def_alt_jq);
                alt_counts[slot]++;
            }
            ep[n++] = jp;
        }
        base_idx += cnt;
    }
    return n;
}

/* ---- A5: log-space helpers (snpcaller.c:692-700, 729-741) ---------------- */
double lfo_log_sum(double a, double b)
{
    double hi = a > b ? a : b, lo = a > b ? b : a;
    return hi + log1p(exp(lo - hi));
}

double lfo_tailsum(const double *row, int from, int len)
{
    double t = row[from];
    int k;
    for (k = from + 1; k < len; k++) t = lfo_log_sum(t, row[k]);
    return t;
}

/* expl with the reference's exception clamp (snpcaller.c:924-936, 1047-1059,
 * 1169-1188). The caller decides where the flags were cleared. */
static long double clamp_after(long double p)
{
    if (errno || fetestexcept(FE_INVALID | FE_DIVBYZERO | FE_OVERFLOW | FE_UNDERFLOW))
        return (p < DBL_EPSILON) ? LDBL_MIN : LDBL_MAX;
    return p;
}

/* ---- A6: pruned Poisson-binomial DP (snpcaller.c:830-971) ---------------- *
 * One row updated in place from high k to low k, which reads exactly the
 * values the reference's two-row version reads (row n-1 entries k and k-1).
 * row[k<K] = ln P(k errors in the first n reads), row[K] = ln P(>=K errors). */
double *lfo_pruned_dp(const double *ep, int N, int K, long long bonf, double sig)
{
    double *row = malloc(sizeof(double) * (K + 1));
    int n, k;
    if (!row) return NULL;
    row[0] = 0.0;
    for (n = 1; n <= N; n++) {
        double p = ep[n - 1];
        double lp = (fabs(p) < DBL_EPSILON) ? log(DBL_EPSILON) : log(p);
        double lq = (fabs(p - 1.0) < DBL_EPSILON) ? log1p(-p + DBL_EPSILON) : log1p(-p);
        double below_top = (n >= K) ? row[K - 1] : 0.0;   /* row n-1, entry K-1, before it is overwritten */
        int top;

        if (n < K) row[n] = LFO_LOGZERO;
        top = n < K - 1 ? n : K - 1;
        for (k = top; k >= 1; k--)
            row[k] = lfo_log_sum(row[k] + lq, row[k - 1] + lp);
        row[0] = row[0] + lq;

        if (n == K) {
            row[K] = below_top + lp;
        } else if (n > K) {
            long double pv;
            row[K] = lfo_log_sum(row[K], below_top + lp);
            errno = 0;
            feclearexcept(FE_ALL_EXCEPT);
            pv = expl(row[K]);
            pv = clamp_after(pv);
            if (pv * (double)bonf > sig) return row;      /* early exit, row is partial */
        }
    }
    return row;
}

/* ---- A7: poissbin (snpcaller.c:1019-1062) -------------------------------- */
double *lfo_poissbin(long double *pvalue, const double *ep, int N, int K, long long bonf, double sig)
{
    double *row;
    *pvalue = LDBL_MAX;
    row = lfo_pruned_dp(ep, N, K, bonf, sig);
    errno = 0;
    feclearexcept(FE_ALL_EXCEPT);
    *pvalue = expl(row[K]);
    *pvalue = clamp_after(*pvalue);
    return row;
}

/* ---- A8: snpcaller (snpcaller.c:1074-1204), approx_threshold_n path absent
 * (GSL not compiled in: the reference exits when it is requested, :1118-1125) */
int lfo_snpcaller(long double *pv3, const double *ep, int N, const int *counts, long long bonf, double sig)
{
    int i, K = 0;
    long double p_top;
    double *row;
    for (i = 0; i < 3; i++) { pv3[i] = LDBL_MAX; if (counts[i] > K) K = counts[i]; }
    if (K == 0) return 0;
    row = lfo_poissbin(&p_top, ep, N, K, bonf, sig);
    if (!(p_top * (double)bonf > sig)) {
        for (i = 0; i < 3; i++) {
            long double p;
            if (counts[i] == 0) continue;
            errno = 0;
            feclearexcept(FE_ALL_EXCEPT);                 /* cleared BEFORE the tail sum: the exp()
                                                           * underflows inside log_sum count (:1169-1172) */
            p = expl(lfo_tailsum(row, counts[i], K + 1));
            pv3[i] = clamp_after(p);
        }
    }
    free(row);
    return 0;
}

/* ascending comparator with the reference's epsilon tie (utils.c:66-76) */
static int cmp_prob(const void *a, const void *b)
{
    double x = *(const double *)a, y = *(const double *)b;
    if (fabs(x - y) < DBL_EPSILON) return 0;
    return x < y ? -1 : (x > y ? 1 : 0);
}

void lfo_sort_probs(double *ep, int n) { qsort(ep, n, sizeof(double), cmp_prob); }

/* ---- A4: the column glue of call_vars/call_snvs (lofreq_call.c:734-879,
 * 886-935) without VCF output, over a packed batch ------------------------ */
int lfo_call_columns(oracle_conf_t *cf, const oracle_batch_t *b, oracle_out_t *o)
{
    long long c, bonf = cf->bonf_subst;
    double *ep = NULL;
    long long ep_cap = 0;

    for (c = 0; c < b->n_cols; c++) {
        int i, n, any = 0, nreads = 0, cov;
        int alt_bases[3] = {0, 0, 0}, alt_counts[3] = {0, 0, 0}, alt_raw[3] = {0, 0, 0};
        long double pv[3];
        char ref = b->ref_base[c];

        for (i = 0; i < 3; i++) {
            o->alt_counts[3 * c + i] = 0; o->alt_raw_counts[3 * c + i] = 0;
            o->pvalues[3 * c + i] = LDBL_MAX; o->called[3 * c + i] = 0; o->qual[3 * c + i] = -1;
        }
        o->tested[c] = 0; o->bonf_used[c] = 0;
        for (i = 0; i < 4; i++) nreads += b->nt_cnt[4 * c + i];
        cov = b->coverage ? b->coverage[c] : nreads;
        if (b->num_bases) nreads = b->num_bases[c];      /* only the gates below look at it from here on */

        if (ref != 'A' && ref != 'C' && ref != 'G' && ref != 'T') continue;  /* 'N' :892; others are 'N' at HEAD */
        if (nreads * 2 < cov) continue;                                     /* :931 */
        if (nreads < cf->min_cov) continue;                                 /* :747 */

        if (nreads > ep_cap) { free(ep); ep_cap = nreads; ep = malloc(sizeof(double) * ep_cap); }
        n = lfo_column_errprobs(cf, b, c, ep, alt_bases, alt_counts, alt_raw);
        for (i = 0; i < 3; i++) {
            o->alt_counts[3 * c + i] = alt_counts[i];
            o->alt_raw_counts[3 * c + i] = alt_raw[i];
            any |= (alt_counts[i] != 0);
        }
        if (!any) continue;                                                 /* not a test :768-780 */

        lfo_sort_probs(ep, n);                                              /* :784 */
        if (cf->bonf_dynamic) bonf = (bonf == 1) ? 3 : bonf + 3;            /* :794-800 */
        cf->num_snv_tests += 3;                                             /* :801 */
        o->tested[c] = 1;
        o->bonf_used[c] = bonf;

        lfo_snpcaller(pv, ep, n, alt_counts, bonf, cf->sig);
        for (i = 0; i < 3; i++) {
            o->pvalues[3 * c + i] = pv[i];
            if (pv[i] * (double)bonf < cf->sig) {                           /* :832 */
                o->called[3 * c + i] = 1;
                o->qual[3 * c + i] = lfo_prob_to_phred(pv[i]);              /* :863 */
            }
        }
    }
    free(ep);
    cf->bonf_subst = bonf;
    return 0;
}
