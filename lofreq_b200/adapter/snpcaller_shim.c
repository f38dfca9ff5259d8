/* Link-level swap: the un-prefixed symbols of the reference's snpcaller.h / binom.h on top of liblofreq_b200.so, for a
 * maintainer who wants every existing caller — call_snvs / call_alt_ins / call_alt_del (lofreq_call.c:319,384,758,807),
 * uniq_snv (lofreq_uniq.c:299,311,381), source_qual (plp.c:554) — on the GPU without touching a call site.
 *
 *   snpcaller()        snpcaller.h:97-102     -> lfb200_snpcaller
 *   poissbin()         snpcaller.h:93-96      -> lfb200_poissbin        (malloc'ed row, the caller free()s it)
 *   plp_to_errprobs()  snpcaller.h:72-75      -> lfb200_plp_to_errprobs (malloc'ed vector, the caller free()s it)
 *   binom()            binom.h:32             -> lfb200_binom
 *
 * Build: add this file to lofreq_SOURCES and -llofreq_b200 to LIBS (src/lofreq/Makefile.am:4-28,44); compile the
 * reference's snpcaller.c with  -Dsnpcaller=ref_snpcaller -Dpoissbin=ref_poissbin -Dplp_to_errprobs=ref_plp_to_errprobs
 * (and binom.c with -Dbinom=ref_binom) so that what else lives in those objects (init_varcall_conf, dump_varcall_conf,
 * plp_to_ins_errprobs, plp_to_del_errprobs) keeps linking.  One kernel launch per call: correct, and what the reference's
 * own tests exercise, but launch-latency-bound — the batched callback (lofreq_adapter.c) is the door for throughput.
 * oracle/Makefile builds exactly this configuration (_ref/libcallswap.so) and tests/test_vcf_boundary.py runs the
 * reference's unmodified call_vars() on top of it. */
#include <stdlib.h>

#include "snpcaller.h"
#include "plp.h"

#include "lofreq_b200.h"

int snpcaller(long double *snp_pvalues, const double *err_probs, const int num_err_probs, const int *noncons_counts,
              const long long int bonf_factor, const double sig_level, const int approx_treshold_n)
{
    return lfb200_snpcaller(snp_pvalues, err_probs, num_err_probs, noncons_counts, bonf_factor, sig_level, approx_treshold_n);
}

double *poissbin(long double *pvalue, const double *err_probs, const int num_err_probs, const int num_failures,
                 const long long int bonf, const double sig)
{
    return lfb200_poissbin(pvalue, err_probs, num_err_probs, num_failures, bonf, sig);
}

void plp_to_errprobs(double **err_probs, int *num_err_probs, int *alt_bases, int *alt_counts, int *alt_raw_counts,
                     const plp_col_t *p, varcall_conf_t *conf)
{
    lfb200_plp_col_t c;
    lfb200_conf_t gc;
    int i;
    lfb200_init_conf(&gc);
    gc.min_bq = conf->min_bq;   gc.min_alt_bq = conf->min_alt_bq;   gc.def_alt_bq = conf->def_alt_bq;
    gc.min_jq = conf->min_jq;   gc.min_alt_jq = conf->min_alt_jq;   gc.def_alt_jq = conf->def_alt_jq;
    gc.flag = conf->flag;
    c.ref_base = p->ref_base;
    c.coverage_plp = p->coverage_plp;
    for (i = 0; i < 4; i++) {                             /* nt4 A,C,G,T; N is skipped, snpcaller.c:386 */
        c.n[i] = (int)p->base_quals[i].n;
        c.base_quals[i] = p->base_quals[i].data;
        c.map_quals[i] = p->map_quals[i].n ? p->map_quals[i].data : NULL;          /* snpcaller.c:444-463 */
        c.baq_quals[i] = p->baq_quals[i].n ? p->baq_quals[i].data : NULL;
        c.source_quals[i] = p->source_quals[i].n ? p->source_quals[i].data : NULL;
    }
    lfb200_plp_to_errprobs(err_probs, num_err_probs, alt_bases, alt_counts, alt_raw_counts, &c, &gc);
}

int binom(double *p, double *q, int num_trials, int num_success, double succ_prob)
{
    return lfb200_binom(p, q, num_trials, num_success, succ_prob);
}
