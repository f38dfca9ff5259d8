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
 * Indel calling (conf->no_indels == 0) stays with the reference's call_indels(), called in place.
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

extern long long int num_snv_tests;                              /* lofreq_call.c:84 */
void call_indels(const plp_col_t *p, varcall_conf_t *conf);      /* lofreq_call.c:618 */

typedef struct {
    varcall_conf_t *conf;            /* the caller's conf: vcf_out, bonf_subst */
    lfb200_ctx *ctx;
    lfb200_builder *bld;
    lfb200_conf_t cf;
    /* per buffered column: target name (interned) and position */
    char **targets;
    int n_targets;
    int *col_tid, *col_pos;
    long long n_cols, cap_cols;
    int failed;
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

/* one called allele -> one VCF record, through the reference's own writer */
static void on_variant(const lfb200_variant_t *v, void *user)
{
    var_t *var;
    char info[512], ref[2], alt[2];
    (void)user;
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
    if (lfb200_builder_create(&A.bld, A.ctx, &A.cf, batch, NULL, NULL) || lfb200_builder_on_variant(A.bld, on_variant, NULL)) {
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

/* drop-in for call_vars() (lofreq_call.c:886-935) */
void lfb200_call_vars(const plp_col_t *p, void *confp)
{
    varcall_conf_t *conf = (varcall_conf_t *)confp;
    const int *bq[4], *mq[4], *baq[4], *sq[4];
    int n[4], g;
    if (A.failed) return;
    if (!A.bld && adapter_init(conf)) { A.failed = 1; return; }

    if (p->ref_base == 'N') return;                                       /* :892 */
    if (!conf->no_indels) call_indels(p, conf);                           /* :896 — stays reference code */
    if (conf->only_indels || p->cons_base[0] == '+' || p->cons_base[0] == '-' ||
        p->num_bases * 2 < p->coverage_plp) return;                       /* :928-932 */

    for (g = 0; g < 4; g++) {
        n[g] = (int)p->base_quals[g].n;
        bq[g] = p->base_quals[g].data;
        mq[g] = p->map_quals[g].n ? p->map_quals[g].data : NULL;         /* snpcaller.c:444-463: empty varray = absent */
        baq[g] = p->baq_quals[g].n ? p->baq_quals[g].data : NULL;
        sq[g] = p->source_quals[g].n ? p->source_quals[g].data : NULL;
    }
    if (lfb200_builder_pending(A.bld) == 0) A.n_cols = 0;                 /* the builder flushed: tags start over */
    if (A.n_cols == A.cap_cols) {
        A.cap_cols = A.cap_cols ? 2 * A.cap_cols : 4096;
        A.col_tid = realloc(A.col_tid, A.cap_cols * sizeof(int));
        A.col_pos = realloc(A.col_pos, A.cap_cols * sizeof(int));
    }
    A.col_tid[A.n_cols] = intern_target(p->target);
    A.col_pos[A.n_cols] = p->pos;
    if (lfb200_builder_add_column_strands(A.bld, A.n_cols, p->ref_base, p->coverage_plp, p->num_bases, bq, mq, baq, sq, n,
                                          p->fw_counts, p->rv_counts)) {
        LOG_FATAL("lofreq_b200: %s\n", lfb200_last_error());
        A.failed = 1;
        return;
    }
    A.n_cols++;
    sync_counters();
}

/* once after mpileup() has returned (lofreq_call.c:1477): the last, partial batch.  Returns non-zero on failure. */
int lfb200_flush(void *confp)
{
    int i, rc = A.failed;
    (void)confp;
    if (A.bld) {
        if (!rc && lfb200_builder_flush(A.bld)) {
            LOG_FATAL("lofreq_b200: %s\n", lfb200_last_error());
            rc = 1;
        }
        sync_counters();
        lfb200_builder_destroy(A.bld);
        lfb200_destroy(A.ctx);
    }
    for (i = 0; i < A.n_targets; i++) free(A.targets[i]);
    free(A.targets); free(A.col_tid); free(A.col_pos);
    memset(&A, 0, sizeof(A));
    return rc;
}
