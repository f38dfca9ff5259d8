/* The adapter a LoFreq maintainer compiles into src/lofreq/ (add this file to lofreq_SOURCES, link with
 * -llofreq_b200): the reference's per-column callback with its exact signature,
 *
 *      void (*plp_proc_func)(const plp_col_t *, void *)                  plp.h:159-163, plp.c:1443
 *
 * in front of the batched GPU door.  In main_call() (lofreq_call.c:1474-1489):
 *
 *      plp_proc_func = &lfb200_call_vars;                      // was &call_vars
 *      rc = mpileup(&mplp_conf, plp_proc_func, (void*)&varcall_conf, 1, (const char **) argv + optind + 1);
 *      if (lfb200_flush((void*)&varcall_conf)) rc = 1;         // the one added line
 *
 * lfb200_call_vars() applies call_vars()'s own gates that need the plp_col_t (lofreq_call.c:892, 928-932), copies what
 * the SNV test reads out of the column (the pointer is only valid during the call, plp.c:1445) and batches it;
 * every LFB200_BATCH_COLS columns (default 65536) and at lfb200_flush() the batch runs on the GPU and each called allele
 * is written through the reference's own vcf_write_var() with the INFO string report_var() would have produced
 * (DP, AF, SB, DP4, HQA; lofreq_call.c:92-137).  conf->bonf_subst and the global num_snv_tests advance exactly as
 * call_snvs() advances them (lofreq_call.c:794-801), so the dynamic-Bonferroni filter step and the
 * "Number of substitution tests performed" line (lofreq_call.c:1524, 1562) see the same numbers.
 * Indel calling (conf->no_indels == 0): the bookkeeping of call_indels() (lofreq_call.c:618-726: min_cov gate, the
 * poly-AT filter, one test per event in hash order, conf->bonf_indel / num_indel_tests advanced per test) is done here
 * per column, the tests themselves are batched through lfb200_indel_tests, and the records of a column come out in the
 * reference's order (insertions, deletions, then substitutions).
 *
 * This file includes the reference's own headers, so it is compiled inside the reference tree — or, for the parity
 * test, by oracle/Makefile against /root/reference (tests/test_vcf_boundary.py diffs the VCF text of both paths). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "snpcaller.h"
#include "plp.h"
#include "vcf.h"
#include "log.h"

#include "lofreq_b200.h"

#include "uthash.h"

extern long long int num_snv_tests;                              /* lofreq_call.c:84 */
extern long long int num_indel_tests;                            /* lofreq_call.c:85 */
extern long int indel_calls_wo_idaq;                             /* lofreq_call.c:88 */

/* one buffered indel test = one iteration of the HASH_ITER loops of call_indels (lofreq_call.c:684-725) */
typedef struct {
    long long col;                   /* adapter column (tag) */
    int is_del, count, dp, hrun, has_aqs;
    char ref_base;
    char key[MAX_INDELSIZE];
    lfb200_dp4_t dp4;
    long long bonf;
} indel_test_t;

typedef struct {
    varcall_conf_t *conf;            /* the caller's conf: vcf_out, bonf_subst */
    lfb200_ctx *ctx;
    lfb200_builder *bld;
    lfb200_conf_t cf;
    /* per buffered column: target name (interned) and position */
    char **targets;
    int n_targets;
    int *col_tid, *col_pos;
    long long n_cols, cap_cols, batch_cols;
    int failed;
    /* indel tests of the buffered columns: reads of test t are [read_off[t], read_off[t+1]) of the planes, the reads of
     * the tested event last (lfb200_indel_tests) */
    indel_test_t *tests;
    long long n_tests, cap_tests;
    long long *read_off;
    unsigned char *iq, *mq, *aq, *sq;
    long long n_reads, cap_reads;
    /* called indels of the batch being flushed, in column order, and how far they have been written */
    long long *hit;                  /* indices into tests */
    int *hit_qual, *hit_sb;
    long long n_hit, next_hit;
} adapter_t;

static adapter_t A;

static int intern_target(const char *t)
{
    int i;
    if (A.n_targets && 0 == strcmp(A.targets[A.n_targets - 1], t)) return A.n_targets - 1;
    for (i = 0; i < A.n_targets; i++)
        if (0 == strcmp(A.targets[i], t)) return i;
    A.targets = realloc(A.targets, (A.n_targets + 1) * sizeof(char *));
    A.targets[A.n_targets] = strdup(t);
    return A.n_targets++;
}

/* a called indel -> one VCF record (call_alt_ins / call_alt_del + report_var, lofreq_call.c:305-426, 92-137) */
static void write_indel(long long h)
{
    const indel_test_t *t = &A.tests[A.hit[h]];
    var_t *var;
    char info[512], ref[MAX_INDELSIZE + 2], alt[MAX_INDELSIZE + 2];
    const float af = t->count / ((float)t->dp);               /* count / (coverage_plp - num_tails), :327,393 */
    vcf_new_var(&var);
    var->chrom = strdup(A.targets[A.col_tid[t->col]]);
    var->pos = A.col_pos[t->col];
    if (!t->has_aqs) indel_calls_wo_idaq += 1;                  /* report_var, :106-108 */
    ref[0] = alt[0] = t->ref_base;
    ref[1] = alt[1] = '\0';
    strcpy(t->is_del ? ref + 1 : alt + 1, t->key);              /* del_to_str / ins_to_str, :258-303 */
    var->ref = strdup(ref);
    var->alt = strdup(alt);
    if (A.hit_qual[h] > -1) var->qual = A.hit_qual[h];
    if (lfb200_format_indel_info(info, sizeof(info), t->dp, af, A.hit_sb[h], &t->dp4, t->hrun) < 0) info[0] = '\0';
    var->info = strdup(info);
    vcf_write_var(&A.conf->vcf_out, var);
    vcf_free_var(&var);
}

/* one called allele -> one VCF record, through the reference's own writer */
static void on_variant(const lfb200_variant_t *v, void *user)
{
    var_t *var;
    char info[512], ref[2], alt[2];
    (void)user;
    /* the indel records of every column up to this one come first (call_vars: call_indels before call_snvs) */
    while (A.next_hit < A.n_hit && A.tests[A.hit[A.next_hit]].col <= v->tag) write_indel(A.next_hit++);
    vcf_new_var(&var);
    var->chrom = strdup(A.targets[A.col_tid[v->tag]]);
    var->pos = A.col_pos[v->tag];
    ref[0] = v->ref_base; alt[0] = v->alt_base; ref[1] = alt[1] = '\0';
    var->ref = strdup(ref);
    var->alt = strdup(alt);
    if (v->qual > -1) var->qual = v->qual;
    if (lfb200_format_snv_info(info, sizeof(info), v->dp, v->af, v->sb, &v->dp4, v->hqa) < 0) info[0] = '\0';
    var->info = strdup(info);
    vcf_write_var(&A.conf->vcf_out, var);
    vcf_free_var(&var);
}

static int adapter_init(varcall_conf_t *conf)
{
    const char *bc = getenv("LFB200_BATCH_COLS");
    long long batch = bc ? atoll(bc) : 65536;
    memset(&A, 0, sizeof(A));
    A.conf = conf;
    if (lfb200_create(&A.ctx, 0)) {
        LOG_FATAL("lofreq_b200: %s\n", lfb200_last_error());
        return 1;
    }
    lfb200_init_conf(&A.cf);
    A.cf.min_bq = conf->min_bq;         A.cf.min_alt_bq = conf->min_alt_bq;   A.cf.def_alt_bq = conf->def_alt_bq;
    A.cf.min_jq = conf->min_jq;         A.cf.min_alt_jq = conf->min_alt_jq;   A.cf.def_alt_jq = conf->def_alt_jq;
    A.cf.min_cov = conf->min_cov;       A.cf.bonf_dynamic = conf->bonf_dynamic;
    A.cf.flag = conf->flag;             A.cf.sig = conf->sig;
    A.cf.bonf_subst = conf->bonf_subst; A.cf.num_snv_tests = num_snv_tests;
    if (conf->approx_threshold_n > 0) {
        LOG_FATAL("%s\n", "lofreq_b200: --approx-threshold is not provided by this path (snpcaller.c:1118-1125)");
        return 1;
    }
    if (batch < 1) batch = 1;
    A.batch_cols = batch;
    /* the adapter flushes itself (indel records have to be interleaved): the builder never does */
    if (lfb200_builder_create(&A.bld, A.ctx, &A.cf, 1ll << 40, NULL, NULL) || lfb200_builder_on_variant(A.bld, on_variant, NULL)) {
        LOG_FATAL("lofreq_b200: %s\n", lfb200_last_error());
        return 1;
    }
    return 0;
}

static void sync_counters(void)
{
    A.conf->bonf_subst = A.cf.bonf_subst;            /* lofreq_call.c:794-800 */
    num_snv_tests = A.cf.num_snv_tests;               /* lofreq_call.c:801 */
}

static unsigned char qb(int q) { return q < 0 ? 255 : q > 254 ? 254 : (unsigned char)q; }

static void push_read(int iq, int mq, int aq, int sq)
{
    if (A.n_reads == A.cap_reads) {
        A.cap_reads = A.cap_reads ? 2 * A.cap_reads : 1 << 16;
        A.iq = realloc(A.iq, A.cap_reads + 32); A.mq = realloc(A.mq, A.cap_reads + 32);
        A.aq = realloc(A.aq, A.cap_reads + 32); A.sq = realloc(A.sq, A.cap_reads + 32);
    }
    A.iq[A.n_reads] = qb(iq); A.mq[A.n_reads] = qb(mq); A.aq[A.n_reads] = qb(aq); A.sq[A.n_reads] = qb(sq);
    A.n_reads++;
}

/* the reads of one event as plp_to_ins_errprobs / plp_to_del_errprobs see them (snpcaller.c:531-560, 592-621) */
static void push_event(const int_varray_t *q, const int_varray_t *aq, const int_varray_t *mq, const int_varray_t *sq, int tested,
                       const varcall_conf_t *conf)
{
    size_t j;
    for (j = 0; j < q->n; j++) {
        int m = -1, a = -1, s = -1;
        if ((conf->flag & VARCALL_USE_IDAQ) && tested) a = aq->data[j];
        if ((conf->flag & VARCALL_USE_MQ) && mq->n) { m = mq->data[j]; if (m == 255) m = -1; }
        if ((conf->flag & VARCALL_USE_SQ) && sq->n) s = sq->data[j];
        push_read(q->data[j], m, a, s);
    }
}

static indel_test_t *new_test(const plp_col_t *p, varcall_conf_t *conf, int is_del, const char *key, int count, const long *fw_rv)
{
    indel_test_t *t;
    if (A.n_tests == A.cap_tests) {
        A.cap_tests = A.cap_tests ? 2 * A.cap_tests : 1024;
        A.tests = realloc(A.tests, A.cap_tests * sizeof(indel_test_t));
        A.read_off = realloc(A.read_off, (A.cap_tests + 1) * sizeof(long long));
    }
    if (A.n_tests == 0) A.read_off[0] = 0;
    t = &A.tests[A.n_tests];
    t->col = A.n_cols;
    t->is_del = is_del;
    t->count = count;
    t->dp = p->coverage_plp - p->num_tails;
    t->hrun = p->hrun;
    t->has_aqs = p->has_indel_aqs;
    t->ref_base = p->ref_base;
    strncpy(t->key, key, MAX_INDELSIZE - 1);
    t->key[MAX_INDELSIZE - 1] = '\0';
    t->dp4.ref_fw = (int)(is_del ? p->non_del_fw_rv[0] : p->non_ins_fw_rv[0]);
    t->dp4.ref_rv = (int)(is_del ? p->non_del_fw_rv[1] : p->non_ins_fw_rv[1]);
    t->dp4.alt_fw = (int)fw_rv[0];
    t->dp4.alt_rv = (int)fw_rv[1];
    if (conf->bonf_dynamic) conf->bonf_indel += 1;                /* lofreq_call.c:693-695, 715-717 */
    num_indel_tests += 1;                                           /* :696, 718 */
    t->bonf = conf->bonf_indel;
    return t;
}

/* call_indels() (lofreq_call.c:618-726) for one column: gates and bookkeeping here, the tests batched */
static void buffer_indel_tests(const plp_col_t *p, varcall_conf_t *conf)
{
    int ign_indels[NUM_NT4] = {0};
    size_t i;
    if (p->num_non_indels + p->num_ins + p->num_dels < conf->min_cov) return;           /* :626 */
    if (p->num_ins && p->ins_quals.n && p->num_dels && p->del_quals.n) {                 /* poly-AT filter, :650-680 */
        const float max_af = 0.05;
        ins_event *ie, *iet;
        del_event *de, *det;
        int ins_dict[NUM_NT4] = {0}, del_dict[NUM_NT4] = {0}, k;
        HASH_ITER(hh_ins, p->ins_event_counts, ie, iet)
            if (strlen(ie->key) == 1 && strchr("AT", ie->key[0]) != NULL) ins_dict[bam_nt4_table[(int)ie->key[0]]] = ie->count;
        HASH_ITER(hh_del, p->del_event_counts, de, det)
            if (strlen(de->key) == 1 && strchr("AT", de->key[0]) != NULL) del_dict[bam_nt4_table[(int)de->key[0]]] = de->count;
        for (k = 0; k < NUM_NT4; k++)
            if (ins_dict[k] && del_dict[k]) {
                const float ins_af = ins_dict[k] / ((float)(p->coverage_plp - p->num_tails));
                const float del_af = del_dict[k] / ((float)(p->coverage_plp - p->num_tails));
                if (ins_af < max_af && del_af < max_af) ign_indels[k] = 1;
            }
    }
    if (p->num_ins) {                                                                    /* :684-703 */
        ins_event *it, *tmp, *ev, *tmp2;
        HASH_ITER(hh_ins, p->ins_event_counts, it, tmp) {
            if (strlen(it->key) == 1 && ign_indels[bam_nt4_table[(int)it->key[0]]]) continue;
            new_test(p, conf, 0, it->key, it->count, it->fw_rv);
            for (i = 0; i < p->ins_quals.n; i++)                                         /* reads without the event: snpcaller.c:521-529 */
                push_read(p->ins_quals.data[i], (conf->flag & VARCALL_USE_MQ) ? (p->ins_map_quals.data[i] == 255 ? 254 : p->ins_map_quals.data[i]) : -1, -1, -1);
            HASH_ITER(hh_ins, p->ins_event_counts, ev, tmp2)
                if (ev != it) push_event(&ev->ins_quals, &ev->ins_aln_quals, &ev->ins_map_quals, &ev->ins_source_quals, 0, conf);
            push_event(&it->ins_quals, &it->ins_aln_quals, &it->ins_map_quals, &it->ins_source_quals, 1, conf);
            A.read_off[++A.n_tests] = A.n_reads;
        }
    }
    if (p->num_dels) {                                                                   /* :706-725 */
        del_event *it, *tmp, *ev, *tmp2;
        HASH_ITER(hh_del, p->del_event_counts, it, tmp) {
            if (strlen(it->key) == 1 && ign_indels[bam_nt4_table[(int)it->key[0]]]) continue;
            new_test(p, conf, 1, it->key, it->count, it->fw_rv);
            for (i = 0; i < p->del_quals.n; i++)
                push_read(p->del_quals.data[i], (conf->flag & VARCALL_USE_MQ) ? (p->del_map_quals.data[i] == 255 ? 254 : p->del_map_quals.data[i]) : -1, -1, -1);
            HASH_ITER(hh_del, p->del_event_counts, ev, tmp2)
                if (ev != it) push_event(&ev->del_quals, &ev->del_aln_quals, &ev->del_map_quals, &ev->del_source_quals, 0, conf);
            push_event(&it->del_quals, &it->del_aln_quals, &it->del_map_quals, &it->del_source_quals, 1, conf);
            A.read_off[++A.n_tests] = A.n_reads;
        }
    }
}

/* the buffered batch: indel tests, then the substitution test with the indel records interleaved in column order */
static int adapter_flush_batch(void)
{
    int rc = 0;
    long long t;
    A.n_hit = A.next_hit = 0;
    if (A.n_tests) {
        long double *pv = malloc(A.n_tests * sizeof(long double));
        unsigned char *called = malloc(A.n_tests);
        int *qual = malloc(A.n_tests * sizeof(int)), *cnt = malloc(A.n_tests * sizeof(int));
        long long *bonf = malloc(A.n_tests * sizeof(long long));
        lfb200_dp4_t *tabs;
        for (t = 0; t < A.n_tests; t++) { cnt[t] = A.tests[t].count; bonf[t] = A.tests[t].bonf; }
        A.cf.sig = A.conf->sig;
        if (lfb200_indel_tests(A.ctx, &A.cf, A.n_tests, A.read_off, A.iq, A.mq, A.aq, A.sq, cnt, bonf, pv, NULL, NULL, called, qual)) {
            LOG_FATAL("lofreq_b200: %s\n", lfb200_last_error());
            rc = 1;
        } else {
            A.hit = realloc(A.hit, A.n_tests * sizeof(long long));
            A.hit_qual = realloc(A.hit_qual, A.n_tests * sizeof(int));
            A.hit_sb = realloc(A.hit_sb, A.n_tests * sizeof(int));
            tabs = malloc(A.n_tests * sizeof(lfb200_dp4_t));
            for (t = 0; t < A.n_tests; t++)
                if (called[t]) {                                    /* pvalue * bonf_indel < sig, lofreq_call.c:326,392 */
                    A.hit[A.n_hit] = t;
                    A.hit_qual[A.n_hit] = qual[t];
                    tabs[A.n_hit++] = A.tests[t].dp4;
                }
            if (A.n_hit && lfb200_sb_qual_batch(A.ctx, A.n_hit, tabs, A.hit_sb)) {
                LOG_FATAL("lofreq_b200: %s\n", lfb200_last_error());
                rc = 1;
            }
            free(tabs);
        }
        free(pv); free(called); free(qual); free(cnt); free(bonf);
    }
    if (!rc && lfb200_builder_flush(A.bld)) {
        LOG_FATAL("lofreq_b200: %s\n", lfb200_last_error());
        rc = 1;
    }
    if (!rc)
        while (A.next_hit < A.n_hit) write_indel(A.next_hit++);     /* indels of the columns after the last called substitution */
    A.n_tests = 0;
    A.n_reads = 0;
    A.n_cols = 0;
    return rc;
}

/* drop-in for call_vars() (lofreq_call.c:886-935) */
void lfb200_call_vars(const plp_col_t *p, void *confp)
{
    varcall_conf_t *conf = (varcall_conf_t *)confp;
    const int *bq[4], *mq[4], *baq[4], *sq[4];
    int n[4], g;
    if (A.failed) return;
    if (!A.bld && adapter_init(conf)) { A.failed = 1; return; }

    if (p->ref_base == 'N') return;                                       /* :892 */
    if (A.n_cols == A.cap_cols) {
        A.cap_cols = A.cap_cols ? 2 * A.cap_cols : 4096;
        A.col_tid = realloc(A.col_tid, A.cap_cols * sizeof(int));
        A.col_pos = realloc(A.col_pos, A.cap_cols * sizeof(int));
    }
    A.col_tid[A.n_cols] = intern_target(p->target);
    A.col_pos[A.n_cols] = p->pos;
    if (!conf->no_indels) buffer_indel_tests(p, conf);                    /* :896 */
    if (conf->only_indels || p->cons_base[0] == '+' || p->cons_base[0] == '-' ||
        p->num_bases * 2 < p->coverage_plp) {                             /* :928-932: no substitution test on this column */
        if (!conf->no_indels) {                                           /* but its indel records need a column slot */
            A.n_cols++;
            if (A.n_cols >= A.batch_cols && adapter_flush_batch()) A.failed = 1;
        }
        return;
    }

    for (g = 0; g < 4; g++) {
        n[g] = (int)p->base_quals[g].n;
        bq[g] = p->base_quals[g].data;
        mq[g] = p->map_quals[g].n ? p->map_quals[g].data : NULL;         /* snpcaller.c:444-463: empty varray = absent */
        baq[g] = p->baq_quals[g].n ? p->baq_quals[g].data : NULL;
        sq[g] = p->source_quals[g].n ? p->source_quals[g].data : NULL;
    }
    if (lfb200_builder_add_column_strands(A.bld, A.n_cols, p->ref_base, p->coverage_plp, p->num_bases, bq, mq, baq, sq, n,
                                          p->fw_counts, p->rv_counts)) {
        LOG_FATAL("lofreq_b200: %s\n", lfb200_last_error());
        A.failed = 1;
        return;
    }
    A.n_cols++;
    if (A.n_cols >= A.batch_cols && adapter_flush_batch()) A.failed = 1;
    sync_counters();
}

/* once after mpileup() has returned (lofreq_call.c:1477): the last, partial batch.  Returns non-zero on failure. */
int lfb200_flush(void *confp)
{
    int i, rc = A.failed;
    (void)confp;
    if (A.bld) {
        if (!rc && adapter_flush_batch()) rc = 1;
        sync_counters();
        lfb200_builder_destroy(A.bld);
        lfb200_destroy(A.ctx);
    }
    for (i = 0; i < A.n_targets; i++) free(A.targets[i]);
    free(A.targets); free(A.col_tid); free(A.col_pos);
    free(A.tests); free(A.read_off); free(A.iq); free(A.mq); free(A.aq); free(A.sq); free(A.hit); free(A.hit_qual); free(A.hit_sb);
    memset(&A, 0, sizeof(A));
    return rc;
}
