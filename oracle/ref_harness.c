/* TEST INFRASTRUCTURE — not part of the shipped product.
 *
 * Driver for the UNMODIFIED reference implementation of the SNV-test path.
 * It is linked together with /root/reference/src/lofreq/{snpcaller.c,utils.c,
 * log.c} (compiled from where they lie, see oracle/Makefile) into
 * oracle/_ref/libsnpref.so.  Nothing of the reference is copied here: this file
 * only (1) unpacks a column batch (oracle/column_batch.h) into the reference's
 * own plp_col_t with the reference's own int_varray_add_value(), and (2) walks
 * the same steps call_vars()/call_snvs() take around the reference's
 * plp_to_errprobs() / qsort(dbl_cmp) / snpcaller() (src/lofreq/lofreq_call.c:
 * 734-879, 886-935), minus the VCF writing, recording what those functions
 * return.
 */
#include <stdlib.h>
#include <string.h>
#include <float.h>
#include <limits.h>

#include "snpcaller.h"   /* reference header: varcall_conf_t, plp_col_t, snpcaller() ... */
#include "utils.h"       /* reference header: int_varray_*, dbl_cmp, PROB_TO_PHREDQUAL */

#include "column_batch.h"

/* normally defined in plp.c (src/lofreq/plp.c:49), which needs real htslib */
const char *bam_nt4_rev_table = "ACGTN";

static void col_init(plp_col_t *p)
{
    int b;
    memset(p, 0, sizeof(*p));
    for (b = 0; b < NUM_NT4; b++) {
        int_varray_init(&p->base_quals[b], 0);
        int_varray_init(&p->baq_quals[b], 0);
        int_varray_init(&p->map_quals[b], 0);
        int_varray_init(&p->source_quals[b], 0);
    }
}

static void col_release(plp_col_t *p)
{
    int b;
    for (b = 0; b < NUM_NT4; b++) {
        int_varray_free(&p->base_quals[b]);
        int_varray_free(&p->baq_quals[b]);
        int_varray_free(&p->map_quals[b]);
        int_varray_free(&p->source_quals[b]);
    }
}

static int plane_val(const unsigned char *plane, long long idx, int sentinel_is_na)
{
    int v = plane[idx];
    if (sentinel_is_na && v == 255) return -1;
    return v;
}

/* Run the reference path over a batch. Returns 0 on success. */
int lfref_call_columns(oracle_conf_t *oc, const oracle_batch_t *b, oracle_out_t *o)
{
    varcall_conf_t conf;
    long long c;
    static char target[] = "synthetic";

    init_varcall_conf(&conf);
    conf.min_bq = oc->min_bq;       conf.min_alt_bq = oc->min_alt_bq;   conf.def_alt_bq = oc->def_alt_bq;
    conf.min_jq = oc->min_jq;       conf.min_alt_jq = oc->min_alt_jq;   conf.def_alt_jq = oc->def_alt_jq;
    conf.min_cov = oc->min_cov;     conf.bonf_dynamic = oc->bonf_dynamic;
    conf.flag = oc->flag;           conf.sig = oc->sig;
    conf.bonf_subst = oc->bonf_subst;
    conf.no_indels = 1;             /* what main_call does by default (lofreq_call.c:1013) */

    for (c = 0; c < b->n_cols; c++) {
        plp_col_t col;
        long long idx = b->col_off[c];
        int g, j, nreads = 0;
        double *ep = NULL;
        int n_ep = 0, i, got_alt = 0;
        int alt_bases[NUM_NONCONS_BASES], alt_counts[NUM_NONCONS_BASES], alt_raw[NUM_NONCONS_BASES];
        long double pv[NUM_NONCONS_BASES];

        for (i = 0; i < 3; i++) {
            o->alt_counts[3*c+i] = 0; o->alt_raw_counts[3*c+i] = 0;
            o->pvalues[3*c+i] = LDBL_MAX; o->called[3*c+i] = 0; o->qual[3*c+i] = -1;
        }
        o->tested[c] = 0; o->bonf_used[c] = 0;

        col_init(&col);
        col.target = target;
        col.pos = (int)c;
        col.ref_base = b->ref_base[c];
        for (g = 0; g < 4; g++) {
            for (j = 0; j < b->nt_cnt[4*c+g]; j++, idx++) {
                int_varray_add_value(&col.base_quals[g], b->bq[idx]);
                if (b->mq)  int_varray_add_value(&col.map_quals[g], b->mq[idx]);
                if (b->baq) int_varray_add_value(&col.baq_quals[g], plane_val(b->baq, idx, 1));
                if (b->sq)  int_varray_add_value(&col.source_quals[g], plane_val(b->sq, idx, 1));
                nreads++;
            }
        }
        col.num_bases = b->num_bases ? b->num_bases[c] : nreads;
        col.coverage_plp = b->coverage ? b->coverage[c] : nreads;
        col.cons_base[0] = col.ref_base; col.cons_base[1] = '\0';

        /* call_vars gates (lofreq_call.c:892, 928-932), indel branch off */
        if (col.ref_base == 'N' || col.num_bases * 2 < col.coverage_plp) goto next;
        /* HEAD maps every non-ACGT reference base to 'N' upstream of the callback;
         * plp_to_errprobs() would overrun alt_bases[3] otherwise */
        if (!strchr("ACGT", col.ref_base)) goto next;
        /* call_snvs gates (lofreq_call.c:747, 754) */
        if (col.num_bases < conf.min_cov) goto next;

        plp_to_errprobs(&ep, &n_ep, alt_bases, alt_counts, alt_raw, &col, &conf);
        for (i = 0; i < 3; i++) {
            o->alt_counts[3*c+i] = alt_counts[i];
            o->alt_raw_counts[3*c+i] = alt_raw[i];
            if (alt_counts[i]) got_alt = 1;
        }
        if (!got_alt) { free(ep); goto next; }                    /* lofreq_call.c:768-780 */

        qsort(ep, n_ep, sizeof(double), dbl_cmp);                  /* :784 */
        if (conf.bonf_dynamic) {                                   /* :794-800 */
            if (conf.bonf_subst == 1) conf.bonf_subst = NUM_NONCONS_BASES;
            else conf.bonf_subst += NUM_NONCONS_BASES;
        }
        oc->num_snv_tests += NUM_NONCONS_BASES;                    /* :801 */
        o->tested[c] = 1;
        o->bonf_used[c] = conf.bonf_subst;

        if (snpcaller(pv, ep, n_ep, alt_counts, conf.bonf_subst, conf.sig, conf.approx_threshold_n)) {
            free(ep); col_release(&col); return 1;
        }
        for (i = 0; i < 3; i++) {
            o->pvalues[3*c+i] = pv[i];
            if (alt_bases[i] == col.ref_base) continue;
            if (pv[i] * (double)conf.bonf_subst < conf.sig) {      /* :832 */
                o->called[3*c+i] = 1;
                o->qual[3*c+i] = PROB_TO_PHREDQUAL(pv[i]);         /* :863 */
            }
        }
        free(ep);
    next:
        col_release(&col);
    }
    oc->bonf_subst = conf.bonf_subst;
    return 0;
}

/* error-prob vector of one column exactly as plp_to_errprobs() returns it
 * (unsorted); out_ep must hold sum(nt_cnt) doubles. Returns n. */
int lfref_column_errprobs(const oracle_conf_t *oc, const oracle_batch_t *b, long long c,
                          double *out_ep, int *alt_bases, int *alt_counts, int *alt_raw)
{
    varcall_conf_t conf;
    plp_col_t col;
    long long idx = b->col_off[c];
    int g, j, n_ep = 0, nreads = 0;
    double *ep = NULL;
    static char target[] = "synthetic";

    init_varcall_conf(&conf);
    conf.min_bq = oc->min_bq;  conf.min_alt_bq = oc->min_alt_bq;  conf.def_alt_bq = oc->def_alt_bq;
    conf.min_jq = oc->min_jq;  conf.min_alt_jq = oc->min_alt_jq;  conf.def_alt_jq = oc->def_alt_jq;
    conf.flag = oc->flag;
    col_init(&col);
    col.target = target; col.pos = (int)c; col.ref_base = b->ref_base[c];
    for (g = 0; g < 4; g++) {
        for (j = 0; j < b->nt_cnt[4*c+g]; j++, idx++) {
            int_varray_add_value(&col.base_quals[g], b->bq[idx]);
            if (b->mq)  int_varray_add_value(&col.map_quals[g], b->mq[idx]);
            if (b->baq) int_varray_add_value(&col.baq_quals[g], plane_val(b->baq, idx, 1));
            if (b->sq)  int_varray_add_value(&col.source_quals[g], plane_val(b->sq, idx, 1));
            nreads++;
        }
    }
    col.num_bases = nreads;
    col.coverage_plp = b->coverage ? b->coverage[c] : nreads;
    plp_to_errprobs(&ep, &n_ep, alt_bases, alt_counts, alt_raw, &col, &conf);
    if (ep) { memcpy(out_ep, ep, sizeof(double) * n_ep); free(ep); }
    col_release(&col);
    return n_ep;
}

int lfref_prob_to_phredqual(long double p) { return PROB_TO_PHREDQUAL(p); }
int lfref_prob_to_phredqual_safe(double p) { return PROB_TO_PHREDQUAL_SAFE(p); }
double lfref_phredqual_to_prob(int q) { return PHREDQUAL_TO_PROB(q); }
int lfref_sizeof_plp_col(void) { return (int)sizeof(plp_col_t); }
int lfref_sizeof_varcall_conf(void) { return (int)sizeof(varcall_conf_t); }

/* One indel test through the reference's own functions: the column is rebuilt with the reference's add_ins_sequence /
 * add_del_sequence (utils.c:559,619) and int_varray_add_value, then plp_to_ins_errprobs / plp_to_del_errprobs
 * (snpcaller.c:501-623), qsort(dbl_cmp) and snpcaller() with (event count, 0, 0) — the steps of call_indels /
 * call_alt_ins / call_alt_del (lofreq_call.c:305-426, 618-726) minus the VCF writing.
 * Non-indel reads: other_q / other_mq [n_other].  Events: event e owns reads [ev_off[e], ev_off[e+1]) of ev_q / ev_aq /
 * ev_mq / ev_sq (-1 = not available); the event under test is test_event.  Returns the number of error probabilities. */
int lfref_indel_test(int is_del, int flag, double sig, long long bonf, int n_other, const int *other_q, const int *other_mq,
                     int n_events, const int *ev_off, const int *ev_q, const int *ev_aq, const int *ev_mq, const int *ev_sq,
                     int test_event, long double *pvalue, int *event_count)
{
    varcall_conf_t conf;
    plp_col_t p;
    double *ep = NULL;
    int n_ep = 0, i, e, counts[3] = {0, 0, 0};
    long double pv[3];
    char key[MAX_INDELSIZE], test_key[MAX_INDELSIZE];
    init_varcall_conf(&conf);
    conf.flag = flag;
    memset(&p, 0, sizeof(p));
    int_varray_init(&p.ins_quals, 0); int_varray_init(&p.ins_map_quals, 0);
    int_varray_init(&p.del_quals, 0); int_varray_init(&p.del_map_quals, 0);
    p.coverage_plp = n_other + ev_off[n_events];
    for (i = 0; i < n_other; i++) {
        int_varray_add_value(is_del ? &p.del_quals : &p.ins_quals, other_q[i]);
        int_varray_add_value(is_del ? &p.del_map_quals : &p.ins_map_quals, other_mq[i]);
    }
    test_key[0] = '\0';
    for (e = 0; e < n_events; e++) {
        int len = 1 + e % (MAX_INDELSIZE - 2), j;
        for (j = 0; j < len; j++) key[j] = "ACGT"[(e + j) & 3];
        key[len] = '\0';
        if (e >= 4) key[0] = "ACGT"[(e / 4) & 3];
        if (e == test_event) strcpy(test_key, key);
        for (i = ev_off[e]; i < ev_off[e + 1]; i++) {
            if (is_del) add_del_sequence(&p.del_event_counts, key, ev_q[i], ev_aq[i], ev_mq[i], ev_sq[i], i & 1);
            else add_ins_sequence(&p.ins_event_counts, key, ev_q[i], ev_aq[i], ev_mq[i], ev_sq[i], i & 1);
        }
    }
    if (is_del) {
        del_event *it = find_del_sequence(&p.del_event_counts, test_key);
        plp_to_del_errprobs(&ep, &n_ep, &p, &conf, test_key);
        counts[0] = it ? it->count : 0;
    } else {
        ins_event *it = find_ins_sequence(&p.ins_event_counts, test_key);
        plp_to_ins_errprobs(&ep, &n_ep, &p, &conf, test_key);
        counts[0] = it ? it->count : 0;
    }
    qsort(ep, n_ep, sizeof(double), dbl_cmp);
    snpcaller(pv, ep, n_ep, counts, bonf, sig, -1);
    *pvalue = pv[0];
    *event_count = counts[0];
    free(ep);
    if (is_del) destruct_del_event_counts(&p.del_event_counts); else destruct_ins_event_counts(&p.ins_event_counts);
    int_varray_free(&p.ins_quals); int_varray_free(&p.ins_map_quals);
    int_varray_free(&p.del_quals); int_varray_free(&p.del_map_quals);
    return n_ep;
}
