/* TEST INFRASTRUCTURE — not part of the shipped product.
 *
 * Shared argument conventions for the two CPU checkers under oracle/:
 *   oracle/_ref/libsnpref.so   the UNMODIFIED reference sources compiled where
 *                              they lie (/root/reference/src/lofreq/...), driven
 *                              by oracle/ref_harness.c          (prefix lfref_)
 *   oracle/libsnvoracle.so     a from-scratch C restatement, oracle/snv_oracle.c
 *                              (prefix lfo_)
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference
 * arm may load either of them.
 *
 * A "column batch" is the packed form of what the reference's per-column
 * callback receives (plp_col_t, src/lofreq/plp.h:73-145): for column c the
 * reads are grouped by called base in the order A,C,G,T (the order
 * plp_to_errprobs walks them, src/lofreq/snpcaller.c:383-388), group sizes in
 * nt_cnt[4*c .. 4*c+3], and the quality bytes of the reads at
 * plane[col_off[c] + i], i in [0, sum(nt_cnt[4c..4c+3])).  col_off has
 * n_cols+1 entries; bytes between the end of a column's reads and
 * col_off[c+1] are padding.  A plane pointer may be NULL (quality absent for
 * every read).  Byte value 255 in the baq/sq planes means "-1 / not available"
 * (plp.c:961 stores -1 when BAQ failed); 255 in the mq plane is the SAM
 * "unknown" that the reference itself maps to -1 (snpcaller.c:451-453).
 */
#ifndef LFB200_ORACLE_COLUMN_BATCH_H
#define LFB200_ORACLE_COLUMN_BATCH_H

typedef struct {
    int min_bq, min_alt_bq, def_alt_bq;     /* snpcaller.h:39-41 */
    int min_jq, min_alt_jq, def_alt_jq;     /* snpcaller.h:43-45 */
    int min_cov;                            /* snpcaller.h:50 */
    int bonf_dynamic;                       /* snpcaller.h:47 */
    int flag;                               /* VARCALL_USE_* bits, defaults.h:76-80 */
    float sig;                              /* snpcaller.h:53 (float!) */
    long long bonf_subst;                   /* in: start value; out: final value */
    long long num_snv_tests;                /* in/out running counter (lofreq_call.c:84) */
} oracle_conf_t;

typedef struct {
    long long n_cols;
    const long long *col_off;       /* n_cols+1 */
    const int *nt_cnt;              /* 4*n_cols: A,C,G,T group sizes */
    const char *ref_base;           /* n_cols, uppercase */
    const int *coverage;            /* n_cols or NULL (= sum of nt_cnt) : plp_col_t.coverage_plp */
    const unsigned char *bq, *mq, *baq, *sq;   /* planes; mq/baq/sq may be NULL */
    const int *num_bases;           /* n_cols or NULL (= sum of nt_cnt): plp_col_t.num_bases, which also counts
                                     * reads showing N (plp.c:1019-1022) */
} oracle_batch_t;

typedef struct {
    int *alt_counts;                /* 3*n_cols (filtered), order = A,C,G,T minus ref */
    int *alt_raw_counts;            /* 3*n_cols */
    unsigned char *tested;          /* n_cols: column reached snpcaller() */
    long long *bonf_used;           /* n_cols: bonf_subst handed to snpcaller (0 if untested) */
    long double *pvalues;           /* 3*n_cols: snpcaller output incl. LDBL_MAX / LDBL_MIN sentinels */
    unsigned char *called;          /* 3*n_cols: pvalue*bonf < sig (lofreq_call.c:832) */
    int *qual;                      /* 3*n_cols: PROB_TO_PHREDQUAL(pvalue) where called, else -1 */
} oracle_out_t;

#endif
