/* TEST INFRASTRUCTURE — not part of the shipped product.
 *
 * Callback-boundary oracle: the reference's REAL call_vars() / call_snvs() / report_var()
 * (src/lofreq/lofreq_call.c:92-137, 734-935), vcf.c, fet.c, snpcaller.c, utils.c, log.c, compiled unmodified from where
 * they lie and linked against the stub htslib headers of oracle/refshim2 plus the aborting stubs below (those files name
 * a dozen htslib / plp.c symbols that the post-pileup path never calls).  This file is the fake pileup: it unpacks a
 * column batch (oracle/column_batch.h + strand counts) into the reference's own plp_col_t with the reference's own
 * int_varray_add_value(), hands each column to call_vars(&col, &conf) exactly like mpileup() does (plp.c:1440-1445) and
 * lets the reference write its VCF lines itself.  What comes back: the raw VCF text, conf.bonf_subst, num_snv_tests. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "snpcaller.h"
#include "utils.h"
#include "vcf.h"
#include "plp.h"

#include "column_batch.h"

/* ---- what plp.c / bedidx.c / htslib would provide --------------------------------------------------------------- */
const char *bam_nt4_rev_table = "ACGTN";                                   /* plp.c:49 */
const unsigned char bam_nt4_table[256] = {                                 /* plp.c:71-88: A0 C1 G2 T3 else 4 */
    4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4, 4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4, 4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4, 4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,
    4,0,4,1,4,4,4,2,4,4,4,4,4,4,4,4, 4,4,4,4,3,4,4,4,4,4,4,4,4,4,4,4, 4,0,4,1,4,4,4,2,4,4,4,4,4,4,4,4, 4,4,4,4,3,4,4,4,4,4,4,4,4,4,4,4,
    4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4, 4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4, 4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4, 4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,
    4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4, 4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4, 4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4, 4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4};
static void never(const char *what) { fprintf(stderr, "call_harness: stub %s reached\n", what); abort(); }
faidx_t *fai_load(const char *fn) { never("fai_load"); return NULL; }
void fai_destroy(faidx_t *fai) { never("fai_destroy"); }
int faidx_nseq(const faidx_t *fai) { never("faidx_nseq"); return 0; }
const char *faidx_iseq(const faidx_t *fai, int i) { never("faidx_iseq"); return NULL; }
int faidx_seq_len(const faidx_t *fai, const char *seq) { never("faidx_seq_len"); return 0; }
BGZF *bgzf_open(const char *path, const char *mode) { never("bgzf_open"); return NULL; }
int bgzf_close(BGZF *fp) { never("bgzf_close"); return 0; }
int bgzf_flush(BGZF *fp) { never("bgzf_flush"); return 0; }
ssize_t bgzf_write(BGZF *fp, const void *data, size_t length) { never("bgzf_write"); return 0; }
int64_t bgzf_seek(BGZF *fp, int64_t pos, int whence) { never("bgzf_seek"); return 0; }
int bgzf_getline(BGZF *fp, int delim, kstring_t *str) { never("bgzf_getline"); return 0; }
#include "htslib/tbx.h"
const tbx_conf_t tbx_conf_vcf = {2, 1, 2, 0, '#', 0};
int tbx_index_build(const char *fn, int min_shift, const tbx_conf_t *conf) { never("tbx_index_build"); return 0; }
void init_mplp_conf(mplp_conf_t *c) { memset(c, 0, sizeof(*c)); }
void dump_mplp_conf(const mplp_conf_t *c, FILE *stream) {}
int mpileup(const mplp_conf_t *mplp_conf, void (*plp_proc_func)(const plp_col_t *, void *), void *plp_proc_conf, const int n,
            const char **fn) { never("mpileup"); return 1; }
int source_qual_load_ign_vcf(const char *vcf_path, void *bed) { never("source_qual_load_ign_vcf"); return 1; }
void source_qual_free_ign_vars() {}
void *bed_read(const char *fn) { never("bed_read"); return NULL; }
void bed_destroy(void *_h) {}
int bed_overlap(const void *_h, const char *chr, int beg, int end) { never("bed_overlap"); return 0; }

/* ---- the reference's own entry points and counters (lofreq_call.c:84-85, 886) ------------------------------------- */
extern long long int num_snv_tests, num_indel_tests;
void call_vars(const plp_col_t *p, void *confp);
/* the product's drop-in callback (lofreq_b200/adapter/lofreq_adapter.c), present only in _ref/libcallb200.so */
void lfb200_call_vars(const plp_col_t *p, void *confp) __attribute__((weak));
int lfb200_flush(void *confp) __attribute__((weak));

/* Columns [0, n_cols) through the reference's call_vars(); raw VCF records (no header) go to vcf_path.
 * strand8: per column fw[A,C,G,T] then rv[A,C,G,T] (plp_col_t.fw_counts / rv_counts); pos: 0-based position (NULL: the
 * column index); cons0: first character of cons_base per column (NULL: the reference base); '+' / '-' = consensus indel.
 * Indel events of a column (optional, n_ev_cols > 0): see lfref_call_vars_indels below. */
typedef struct {
    long long n_events;              /* insertion events first, then deletion events, grouped by column in column order */
    const long long *ev_col;         /* column of the event */
    const unsigned char *ev_is_del;
    const char *ev_key;              /* MAX_INDELSIZE bytes per event, NUL-terminated */
    const long long *ev_read_off;    /* n_events + 1: reads of the event in ev_q / ev_aq / ev_mq / ev_sq / ev_rv */
    const int *ev_q, *ev_aq, *ev_mq, *ev_sq;      /* -1 = not available */
    const unsigned char *ev_rv;      /* strand of the read */
    /* reads without an indel, per column: qualities for insertions and for deletions (plp.c:1086-1150) */
    const long long *non_off;        /* n_cols + 1 */
    const int *non_iq, *non_dq, *non_mq;
    const int *hrun;                 /* n_cols */
    const int *num_tails;            /* n_cols */
    const int *non_ins_fw_rv, *non_del_fw_rv;     /* 2 * n_cols each */
} oracle_indels_t;

int lfref_call_vars_vcf(oracle_conf_t *oc, const oracle_batch_t *b, const int *strand8, const int *pos, const char *cons0,
                        const char *target, const char *vcf_path, const oracle_indels_t *ind, long long *bonf_indel_io,
                        long long *num_indel_tests_out, int use_adapter)
{
    void (*callback)(const plp_col_t *, void *) = use_adapter ? lfb200_call_vars : call_vars;
    int rc = 0;
    varcall_conf_t conf;
    long long c, ev = 0;
    char *tgt = strdup(target ? target : "synthetic");

    init_varcall_conf(&conf);
    conf.min_bq = oc->min_bq;       conf.min_alt_bq = oc->min_alt_bq;   conf.def_alt_bq = oc->def_alt_bq;
    conf.min_jq = oc->min_jq;       conf.min_alt_jq = oc->min_alt_jq;   conf.def_alt_jq = oc->def_alt_jq;
    conf.min_cov = oc->min_cov;     conf.bonf_dynamic = oc->bonf_dynamic;
    conf.flag = oc->flag;           conf.sig = oc->sig;
    conf.bonf_subst = oc->bonf_subst;
    conf.no_indels = ind ? 0 : 1;   /* main_call's default is 1 (lofreq_call.c:1013) */
    if (bonf_indel_io) conf.bonf_indel = *bonf_indel_io;
    num_snv_tests = oc->num_snv_tests;
    num_indel_tests = 0;
    if (!callback) { free(tgt); return 2; }                      /* adapter asked for, but this build has none */
    if (vcf_file_open(&conf.vcf_out, vcf_path, 0, 'w')) { free(tgt); return 1; }

    for (c = 0; c < b->n_cols; c++) {
        plp_col_t col;
        long long idx = b->col_off[c];
        int g, j, nreads = 0;
        memset(&col, 0, sizeof(col));
        for (g = 0; g < NUM_NT4; g++) {
            int_varray_init(&col.base_quals[g], 0);
            int_varray_init(&col.baq_quals[g], 0);
            int_varray_init(&col.map_quals[g], 0);
            int_varray_init(&col.source_quals[g], 0);
        }
        int_varray_init(&col.ins_quals, 0); int_varray_init(&col.ins_map_quals, 0); int_varray_init(&col.ins_source_quals, 0);
        int_varray_init(&col.del_quals, 0); int_varray_init(&col.del_map_quals, 0); int_varray_init(&col.del_source_quals, 0);
        col.target = tgt;
        col.pos = pos ? pos[c] : (int)c;
        col.ref_base = b->ref_base[c];
        for (g = 0; g < 4; g++) {
            for (j = 0; j < b->nt_cnt[4*c+g]; j++, idx++) {
                int_varray_add_value(&col.base_quals[g], b->bq[idx]);
                if (b->mq)  int_varray_add_value(&col.map_quals[g], b->mq[idx]);
                if (b->baq) int_varray_add_value(&col.baq_quals[g], b->baq[idx] == 255 ? -1 : b->baq[idx]);
                if (b->sq)  int_varray_add_value(&col.source_quals[g], b->sq[idx] == 255 ? -1 : b->sq[idx]);
                nreads++;
            }
            if (strand8) { col.fw_counts[g] = strand8[8*c+g]; col.rv_counts[g] = strand8[8*c+4+g]; }
        }
        col.num_bases = b->num_bases ? b->num_bases[c] : nreads;
        col.coverage_plp = b->coverage ? b->coverage[c] : nreads;
        col.cons_base[0] = cons0 ? cons0[c] : col.ref_base;
        col.cons_base[1] = '\0';
        if (ind) {
            long long i;
            col.hrun = ind->hrun ? ind->hrun[c] : 0;
            col.num_tails = ind->num_tails ? ind->num_tails[c] : 0;
            col.has_indel_aqs = 1;
            for (i = ind->non_off[c]; i < ind->non_off[c + 1]; i++) {
                int_varray_add_value(&col.ins_quals, ind->non_iq[i]);
                int_varray_add_value(&col.del_quals, ind->non_dq[i]);
                int_varray_add_value(&col.ins_map_quals, ind->non_mq[i]);
                int_varray_add_value(&col.del_map_quals, ind->non_mq[i]);
            }
            col.num_non_indels = (int)(ind->non_off[c + 1] - ind->non_off[c]);
            if (ind->non_ins_fw_rv) { col.non_ins_fw_rv[0] = ind->non_ins_fw_rv[2*c]; col.non_ins_fw_rv[1] = ind->non_ins_fw_rv[2*c+1]; }
            if (ind->non_del_fw_rv) { col.non_del_fw_rv[0] = ind->non_del_fw_rv[2*c]; col.non_del_fw_rv[1] = ind->non_del_fw_rv[2*c+1]; }
            for (; ev < ind->n_events && ind->ev_col[ev] == c; ev++) {
                const char *key = ind->ev_key + (size_t)ev * MAX_INDELSIZE;
                for (i = ind->ev_read_off[ev]; i < ind->ev_read_off[ev + 1]; i++) {
                    if (ind->ev_is_del[ev]) {
                        add_del_sequence(&col.del_event_counts, (char *)key, ind->ev_q[i], ind->ev_aq[i], ind->ev_mq[i], ind->ev_sq[i], ind->ev_rv[i]);
                        col.num_dels++;
                    } else {
                        add_ins_sequence(&col.ins_event_counts, (char *)key, ind->ev_q[i], ind->ev_aq[i], ind->ev_mq[i], ind->ev_sq[i], ind->ev_rv[i]);
                        col.num_ins++;
                    }
                }
            }
        }
        /* HEAD maps every non-ACGT reference base to 'N' upstream of the callback (plp.c) */
        if (!strchr("ACGT", col.ref_base)) col.ref_base = 'N';

        (*callback)(&col, &conf);                                 /* plp.c:1443 */

        for (g = 0; g < NUM_NT4; g++) {
            int_varray_free(&col.base_quals[g]); int_varray_free(&col.baq_quals[g]);
            int_varray_free(&col.map_quals[g]); int_varray_free(&col.source_quals[g]);
        }
        int_varray_free(&col.ins_quals); int_varray_free(&col.ins_map_quals); int_varray_free(&col.ins_source_quals);
        int_varray_free(&col.del_quals); int_varray_free(&col.del_map_quals); int_varray_free(&col.del_source_quals);
        if (col.ins_event_counts) destruct_ins_event_counts(&col.ins_event_counts);
        if (col.del_event_counts) destruct_del_event_counts(&col.del_event_counts);
    }
    if (use_adapter) rc = lfb200_flush(&conf);                    /* the one line added after mpileup() (lofreq_call.c:1477) */
    vcf_file_close(&conf.vcf_out);
    oc->bonf_subst = conf.bonf_subst;
    oc->num_snv_tests = num_snv_tests;
    if (bonf_indel_io) *bonf_indel_io = conf.bonf_indel;
    if (num_indel_tests_out) *num_indel_tests_out = num_indel_tests;
    free(tgt);
    return rc;
}

/* kt_fisher_exact + PROB_TO_PHREDQUAL_SAFE as report_var() uses them (lofreq_call.c:115-125) */
#include "fet.h"
#include <limits.h>
int lfref_sb_qual(int ref_fw, int ref_rv, int alt_fw, int alt_rv, double *two)
{
    double l, r, t;
    if ((ref_fw + ref_rv) == 0 && (alt_fw == 0 || alt_rv == 0)) { if (two) *two = -1.0; return INT_MAX; }
    (void)kt_fisher_exact(ref_fw, ref_rv, alt_fw, alt_rv, &l, &r, &t);
    if (two) *two = t;
    return PROB_TO_PHREDQUAL_SAFE(t);
}
