/* lofreq_b200 — C ABI of the B200-native per-pileup-column SNV test.
 *
 * This is the boundary a LoFreq maintainer binds to.  Plain C types only (no
 * torch, no C++), one shared library: lofreq_b200/lib/liblofreq_b200.so.
 * Every entry point names the reference interface it stands in for; paths are
 * relative to the reference tree (src/lofreq/...).
 *
 * The test itself (quality merge -> error probabilities -> Poisson-binomial
 * tail -> per-allele p-values) runs in hand-written sm_100a CUDA kernels.
 * There is no CPU implementation behind these calls: without a CUDA device
 * lfb200_create() fails and every compute entry point returns an error.
 *
 * Host-side work that stays on the CPU, because it needs x87 long double to
 * make the same decisions as the reference: expl() of the natural-log
 * p-values, the FE-exception clamp of snpcaller.c:1174-1188, the comparison
 * `pvalue * bonf < sig` (lofreq_call.c:832) and PROB_TO_PHREDQUAL
 * (utils.h:45).  It runs only on the few columns the device could not rule
 * out ("sites").
 */
#ifndef LOFREQ_B200_H
#define LOFREQ_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define LFB200_NUM_NONCONS 3          /* NUM_NONCONS_BASES, defaults.h:38 */

/* varcall_conf_t.flag bits (defaults.h:76-80) */
#define LFB200_USE_BAQ  1
#define LFB200_USE_MQ   2
#define LFB200_USE_SQ   4
#define LFB200_USE_IDAQ 8

/* per-allele result status */
#define LFB200_ST_VALUE   0           /* ln p / pvalue hold a computed value              */
#define LFB200_ST_LDBLMAX 1           /* reference returns LDBL_MAX: not computed / pruned / clamped high */
#define LFB200_ST_LDBLMIN 2           /* reference returns LDBL_MIN: clamped low (snpcaller.c:1177)      */
#define LFB200_ST_UNSUPPORTED 3       /* not computed: the allele's column has an alt count above what this build's kernels take;
                                       * the site is reported so that the caller sees it (summary.n_unsupported counts them) */

/* lfb200_site_t.flags */
#define LFB200_SITE_HOST_FINISHED 1   /* a comparison of the decision fell inside its guard band on the device: the host
                                       * repeated the reference's long double sequence for this site */
#define LFB200_SITE_UNSUPPORTED   2

/* The fields of varcall_conf_t (snpcaller.h:38-63) this path reads, same names
 * and meaning; defaults = init_varcall_conf (snpcaller.c:626-651) via
 * lfb200_init_conf().  bonf_subst and num_snv_tests are in/out: the running
 * Bonferroni factor (lofreq_call.c:794-800) and the global test counter
 * (lofreq_call.c:84,801). */
typedef struct {
    int min_bq, min_alt_bq, def_alt_bq;
    int min_jq, min_alt_jq, def_alt_jq;
    int min_cov;
    int bonf_dynamic;
    int flag;
    float sig;
    long long bonf_subst;
    long long num_snv_tests;
} lfb200_conf_t;

/* A batch of pileup columns = the packed form of the plp_col_t objects
 * (plp.h:73-145) the reference hands to its per-column callback
 * (plp.c:1443).  For column c the reads are grouped by called base in A,C,G,T
 * order (the order plp_to_errprobs walks base_quals[], snpcaller.c:383-388),
 * group sizes in nt_cnt[4c..4c+3]; the quality bytes of its reads start at
 * plane[col_off[c]].  Bytes between the end of a column and col_off[c+1] are
 * padding.  mq / baq / sq may be NULL (quality absent).  Byte 255 in mq, baq
 * and sq means "not available" (-1 in the reference; mq==255 is mapped to -1
 * by snpcaller.c:451-453 itself).  coverage may be NULL (= reads in the
 * column); it is plp_col_t.coverage_plp.
 * Device variants: every pointer is a device pointer, planes 16-byte aligned
 * and readable up to the next multiple of 16 bytes past the last read. */
typedef struct {
    long long n_cols;
    const long long *col_off;          /* n_cols + 1 */
    const int *nt_cnt;                 /* 4 * n_cols */
    const char *ref_base;              /* n_cols, uppercase */
    const int *coverage;               /* n_cols or NULL */
    const unsigned char *bq, *mq, *baq, *sq;
    const int *num_bases;              /* n_cols or NULL (= reads in the column): plp_col_t.num_bases, which also
                                        * counts reads showing N (plp.c:1019-1022); only the gates use it */
} lfb200_batch_t;

/* One column the device could not rule out (a candidate variant site), after
 * the decision.  status[] / called[] / qual[] are decided on the device from the
 * natural-log p-values (the comparisons of snpcaller.c:1047-1059, 1155, 1166-1196,
 * lofreq_call.c:832 and the truncation of PROB_TO_PHREDQUAL, utils.h:45, are
 * comparisons of ln p with constants; a site with one of them inside its guard
 * band is re-decided by the host in long double, flags & LFB200_SITE_HOST_FINISHED).
 * pvalue[] are what snpcaller() would have written to snp_pvalues[]
 * (snpcaller.h:97-102) including the LDBL_MAX / LDBL_MIN sentinels — x87 long
 * double images computed by the host, for callers that ask for them
 * (lfb200_set_site_pvalues, default on; zero when off); called[] is the test of
 * lofreq_call.c:832; qual[] is PROB_TO_PHREDQUAL(pvalue) where called, else -1
 * (lofreq_call.c:863). */
typedef struct {
    long long col;                     /* index into the batch */
    long long bonf;                    /* Bonferroni factor handed to the test */
    double lnp[LFB200_NUM_NONCONS];    /* natural log of the tail probability from the device */
    double ln_floor;                   /* min(ln P(X = K-1), ln P(X >= K)), K = largest alt count: input of the
                                        * FE-underflow clamp rule of probvec_tailsum (snpcaller.c:1169-1188) */
    long double pvalue[LFB200_NUM_NONCONS];
    int alt_count[LFB200_NUM_NONCONS]; /* filtered counts, A,C,G,T-minus-ref order (snpcaller.c:489) */
    int alt_raw_count[LFB200_NUM_NONCONS];
    int qual[LFB200_NUM_NONCONS];
    unsigned char status[LFB200_NUM_NONCONS];
    unsigned char called[LFB200_NUM_NONCONS];
    unsigned char flags;               /* LFB200_SITE_* */
    unsigned char reserved;
} lfb200_site_t;

/* Optional dense per-column outputs of the host entry point; any pointer may
 * be NULL.  Layout [n_cols] or [n_cols][3]. */
typedef struct {
    int *alt_counts;                   /* plp_to_errprobs alt_counts  (snpcaller.h:72-75) */
    int *alt_raw_counts;               /* plp_to_errprobs alt_raw_counts */
    unsigned char *tested;             /* column reached snpcaller() (lofreq_call.c:794-807) */
    long long *bonf_used;              /* bonf_subst passed to snpcaller(), 0 if untested */
    double *lnp;                       /* ln p where status == LFB200_ST_VALUE */
    unsigned char *status;
    long double *pvalues;              /* snpcaller() snp_pvalues[] */
    unsigned char *called;
    int *qual;
} lfb200_dense_out_t;

typedef struct {
    long long n_cols, n_tested, n_sites, n_heavy;   /* n_heavy: columns that needed the O(depth*K) kernel */
    long long bonf_subst_final, num_snv_tests;
    long long n_unsupported;           /* columns reported with LFB200_ST_UNSUPPORTED (alt count above 16384) */
} lfb200_summary_t;

typedef struct lfb200_ctx lfb200_ctx;

/* ---- life cycle ------------------------------------------------------- */
int lfb200_create(lfb200_ctx **ctx, int device);     /* 0 = ok; fails without a CUDA device */
void lfb200_destroy(lfb200_ctx *ctx);
const char *lfb200_last_error(void);
/* init_varcall_conf (snpcaller.c:626-651) */
void lfb200_init_conf(lfb200_conf_t *conf);

/* ---- the batched door (new; replaces one call_vars() per column,
 *      lofreq_call.c:886-935 / 734-879, minus the VCF writing) ------------ */
/* Host buffers in, sites out (sorted by column).  conf->bonf_subst and
 * conf->num_snv_tests are advanced exactly as call_snvs() would over the same
 * columns in order.  Returns 0, or non-zero with lfb200_last_error() set;
 * if more than max_sites sites exist the call fails with that message. */
int lfb200_call_columns(lfb200_ctx *ctx, lfb200_conf_t *conf, const lfb200_batch_t *host_batch,
                        const lfb200_dense_out_t *dense, lfb200_site_t *sites, long long max_sites,
                        lfb200_summary_t *summary);
/* How lfb200_call_columns (and the column builder) hands the quality planes to the device: 1 (default) = planes that
 * lie in pinned host memory (cudaHostAlloc / cudaHostRegister) are read in place over PCIe — the kernels only touch the
 * reads that decide a column, so far fewer bytes cross the bus than a bulk copy moves; planes in pageable memory are
 * copied; 0 = every plane is copied to the device first.  Results are identical. */
int lfb200_set_host_planes(lfb200_ctx *ctx, int mode);

/* Device-resident batch, two phases so that region shards on several GPUs can
 * exchange their tested-column counts in between (the running Bonferroni of a
 * shard starts where the previous shard's ends; lofreq2_call_pparallel.py
 * instead restarts it per region and sums the counts at the end, :131-161).
 * stream is a cudaStream_t (or NULL).
 *   screen : gates (lofreq_call.c:747,754,892,931), alt counts and raw counts
 *            from the reads showing a non-reference base, tested flags and the
 *            place of every tested column in the running count, and the first
 *            stage of the early exit under a lower bound of the factor
 *   ntested: number of tested columns found by screen (synchronises)
 *   test   : running Bonferroni from conf->bonf_subst (or from device memory),
 *            the reference's early exit lane-per-column, the O(depth*K) kernels for
 *            the columns it cannot rule out, then the decision (status / called /
 *            QUAL) per site on the device and the sites in column order into
 *            pinned host memory.  Nothing of this needs the host.
 *   sites  : wait for the test, re-decide the few guard-band sites in long
 *            double, fill the long double p-values when asked for (synchronises)
 * One batch per context at a time: screen fails while a lfb200_sites_begin
 * request is pending on the same context.
 * When a phase is queued with the same arguments, buffers and stream as the
 * time before (a resident batch timed repeatedly, a builder refilling its
 * planes), its launches are captured once and replayed as a CUDA graph; this
 * needs a stream other than the legacy default stream and can be switched off
 * with LFB200_NO_GRAPH=1 in the environment.  Results do not depend on it. */
int lfb200_screen_device(lfb200_ctx *ctx, const lfb200_conf_t *conf, const lfb200_batch_t *dev_batch, void *stream);
int lfb200_ntested_device(lfb200_ctx *ctx, void *stream, long long *n_tested);
int lfb200_test_device(lfb200_ctx *ctx, const lfb200_conf_t *conf, void *stream);
/* the same exchange without a host round trip: copy the tested count into caller-owned device memory
 * (e.g. the send buffer of an NCCL all_gather on the same stream), and start the test from a running factor
 * that a kernel of the caller left in device memory */
int lfb200_ntested_copy_device(lfb200_ctx *ctx, void *stream, long long *dst_dev);
int lfb200_test_device_from(lfb200_ctx *ctx, const lfb200_conf_t *conf, void *stream, const long long *bonf_start_dev);
/* start_dev[0] = the running factor shard `rank` starts from, given the all-gathered tested counts of every
 * shard in device memory (lofreq_call.c:794-800 continued across shards) */
int lfb200_bonf_start_device(void *stream, const long long *tested_counts_dev, int rank, long long bonf_subst,
                             long long *start_dev);
/* The exchange between region shards done by the library itself, directly on the kernels' stream (one process per
 * GPU, all on one node):
 *   comm_unique_id : rank 0 creates the 128-byte id (an NCCL unique id; libnccl is dlopen'ed), the caller broadcasts
 *                    it to all ranks by any means
 *   comm_init      : one communicator per context (collective: every rank calls it); also maps the count mailbox, a
 *                    POSIX shared-memory segment named after the id and registered with CUDA
 *   comm_exchange  : after screen — one warp posts {tested columns of this batch, sites_prev_batch} in the mailbox and
 *                    waits for the shards BEFORE this one (a shard continues their running Bonferroni factor; rank 0
 *                    never waits); the starting factor of this shard is left in device memory (*bonf_start_dev, for
 *                    test_device_from).  No collective, no host round trip.  LFB200_EXCHANGE_NCCL=1 in the
 *                    environment selects one ncclAllGather per batch instead.
 *   comm_gathered  : the counts of every shard from the last exchange: one ncclAllGather of 2 x int64 per rank (the
 *                    final gather of the per-region counts) and a host copy; synchronises the stream */
int lfb200_comm_unique_id(unsigned char id[128]);
int lfb200_comm_init(lfb200_ctx *ctx, int world, int rank, const unsigned char id[128]);
int lfb200_comm_exchange(lfb200_ctx *ctx, void *stream, long long bonf_subst, long long sites_prev_batch,
                         const long long **bonf_start_dev);
int lfb200_comm_gathered(lfb200_ctx *ctx, void *stream, long long *tested_all, long long *sites_prev_all);
int lfb200_sites_device(lfb200_ctx *ctx, lfb200_conf_t *conf, void *stream, lfb200_site_t *sites,
                        long long max_sites, lfb200_summary_t *summary);
/* The same without a copy: *sites points at the context's own pinned buffer (summary->n_sites entries, column
 * order), valid until the next lfb200_test_device* on this context. */
int lfb200_sites_view(lfb200_ctx *ctx, lfb200_conf_t *conf, void *stream, const lfb200_site_t **sites,
                      lfb200_summary_t *summary);
/* lfb200_sites_device split in two: begin registers the request and returns at once (with the long double p-values
 * switched on they are computed by a thread owned by the context meanwhile); end waits for the test and hands the
 * sites over.  conf, sites must stay valid in between; one request per context at a time, and the NEXT batch must go
 * to another context until lfb200_sites_end has returned (lfb200_screen_device refuses otherwise: the request reads
 * this context's workspace). */
int lfb200_sites_begin(lfb200_ctx *ctx, lfb200_conf_t *conf, void *stream, lfb200_site_t *sites, long long max_sites);
int lfb200_sites_end(lfb200_ctx *ctx, lfb200_summary_t *summary);
/* sites == NULL in lfb200_sites_device / lfb200_sites_begin: nothing is copied; the sites of the batch are read in
 * place from the context's pinned buffer (summary->n_sites entries, valid until the next test on this context) */
int lfb200_sites_buffer(lfb200_ctx *ctx, const lfb200_site_t **sites);
/* lfb200_site_t.pvalue[] (x87 long double images of the p-values, the only per-site work left on the host):
 * on = 1 (default) fills them for every site; on = 0 leaves them zero — status / called / qual / lnp are complete
 * either way, and lfb200_site_fill_pvalues() computes them later for the sites a caller keeps. */
int lfb200_set_site_pvalues(lfb200_ctx *ctx, int on);
void lfb200_site_fill_pvalues(lfb200_site_t *sites, long long n);
/* optional per-phase device timing with CUDA events on the launching stream (benchmark / roofline):
 * ms4 = { k_front + k_scan_tiles, 0, k_prune2, the O(depth*K) kernels (k_mid | k_dp<0..2> | k_xl, then the fallbacks) }
 * of the last screen + test */
int lfb200_set_profiling(lfb200_ctx *ctx, int on);
int lfb200_get_profile(lfb200_ctx *ctx, float *ms4);
/* copy alt_counts | alt_raw_counts ([n_cols][6] ints) of the last screen into caller-owned device memory */
int lfb200_copy_counts_device(lfb200_ctx *ctx, void *stream, int *dst_dev);
/* which kernels the tested columns of the last finished batch (lfb200_sites_device / _end / lfb200_call_columns) took:
 * out[0] k_dp (8 < K <= 2048, several columns per warp), out[1] the per-column fallback k_heavy_all (columns k_dp or
 * k_mid handed back), out[2] k_heavy_xl (K > 2048, one CTA per column), out[3] k_mid (K <= 8 survivors of the prune) */
int lfb200_last_job_counts(lfb200_ctx *ctx, long long out[4]);
/* The BAQ profile HMM for a batch of reads: replaces one kpa_ext_glocal() call per read (kprobaln_ext.h:38-40,
 * kprobaln_ext.c:80-277) as bam_prob_realn_core_ext makes it (bam_md_ext.c:407) with pd == NULL — the banded glocal
 * forward-backward that yields, per query base, the reference position / state it most probably aligns to and the
 * phred-scaled posterior of that being wrong.  Read r: reference window ref[ref_off[r] .. ref_off[r+1]) and query
 * query[qry_off[r] .. qry_off[r+1]) as 0..3 (4 = ambiguous), base qualities qual (NULL: Q30, as the reference assumes);
 * d, e, bw are kpa_ext_par_t (gap open, gap extension, band width).  state[] and q[] (qry_off[n] entries each, host
 * memory) get exactly the reference's values; reads with an empty window or query are left untouched, as the reference
 * leaves them.  Not covered: the pd matrix of LoFreq's indel-alignment qualities (idaq, bam_md_ext.c:73). */
int lfb200_kpa_glocal_batch(lfb200_ctx *ctx, long long n, const unsigned char *ref, const long long *ref_off,
                            const unsigned char *query, const long long *qry_off, const unsigned char *qual, float d, float e,
                            int bw, int *state, unsigned char *q);
/* how many phases of this context were queued by replaying a captured CUDA graph so far (diagnostics) */
long long lfb200_graph_replays(lfb200_ctx *ctx);
/* measured DFMA/s of this GPU (8 independent chains per thread, ~20 ms): the fp64-pipe roofline denominator */
double lfb200_dfma_peak(lfb200_ctx *ctx, void *stream);
/* device pointers of the per-column results of the last screen/test (valid
 * until the next screen on this ctx): alt_counts and alt_raw_counts are interleaved, stride 6 ints per
 * column (alt_raw_counts = alt_counts + 3); tested u8[n]; bonf_used i64[n] */
int lfb200_device_results(lfb200_ctx *ctx, const int **alt_counts, const int **alt_raw_counts,
                          const unsigned char **tested, const long long **bonf_used);

/* ---- the per-column callback surface ---------------------------------------
 * The reference hands one plp_col_t at a time to a callback
 *   void (*plp_proc_func)(const plp_col_t *, void *)      (plp.h:159-163, plp.c:1443)
 * and the pointer is only valid during the call (plp.c:1445).  A column
 * builder copies what the SNV test needs out of the plp_col_t, batches
 * batch_cols columns, runs them through lfb200_call_columns() and reports
 * every site through on_site in input order — call_vars() (lofreq_call.c:886)
 * becomes lfb200_builder_add_column(), and one lfb200_builder_flush() goes
 * after mpileup() returns (lofreq_call.c:1477).  INTEGRATION.md has the adapter.
 * conf is kept by reference: bonf_subst / num_snv_tests advance at each flush.
 *
 * add_column arguments, per nt4 index A,C,G,T (NUM_NT4 minus N):
 *   quals[i] = plp_col_t.base_quals[i].data, n[i] = plp_col_t.base_quals[i].n   (plp.h:89, utils.h:60-66)
 *   map_quals / baq_quals / source_quals likewise; the array pointer or any
 *   entry may be NULL when that varray is empty (snpcaller.c:444-463 then
 *   treats the quality as absent); a stored -1 means "not available".
 * num_bases = plp_col_t.num_bases (gates of lofreq_call.c:747,931); tag is returned with the site
 * (e.g. plp_col_t.pos). */
typedef void (*lfb200_site_fn)(const lfb200_site_t *site, long long tag, char ref_base, int coverage_plp, void *user);
typedef struct lfb200_builder lfb200_builder;
int lfb200_builder_create(lfb200_builder **bld, lfb200_ctx *ctx, lfb200_conf_t *conf, long long batch_cols,
                          lfb200_site_fn on_site, void *user);
int lfb200_builder_add_column(lfb200_builder *bld, long long tag, char ref_base, int coverage_plp, int num_bases,
                              const int *const base_quals[4], const int *const map_quals[4],
                              const int *const baq_quals[4], const int *const source_quals[4], const int n[4]);
int lfb200_builder_flush(lfb200_builder *bld);
void lfb200_builder_destroy(lfb200_builder *bld);
long long lfb200_builder_pending(const lfb200_builder *bld);      /* columns buffered since the last flush */

/* ---- the reporting tail of a called allele: what report_var() (lofreq_call.c:92-137) adds ------
 * DP4 = strand counts of the reference base and of the alt base (plp_col_t.fw_counts / rv_counts, lofreq_call.c:853-857);
 * SB  = PROB_TO_PHREDQUAL_SAFE of the two-tailed p of Fisher's exact test on DP4 (kt_fisher_exact, fet.c:62-101),
 *       INT_MAX when there are no reference reads and the alt reads sit on one strand (lofreq_call.c:112-113);
 * AF  = alt_raw_count / (float)coverage_plp (lofreq_call.c:835); DP = coverage_plp; HQA = filtered alt count (:860). */
typedef struct { int ref_fw, ref_rv, alt_fw, alt_rv; } lfb200_dp4_t;          /* dp4_counts_t, vcf.h:60-65 */
/* n tables at once, one GPU thread per table; the few results within a guard band of an integer boundary of
 * -10 log10(p) are repeated on the host with glibc (same routine), so SB is the reference's integer */
int lfb200_sb_qual_batch(lfb200_ctx *ctx, long long n, const lfb200_dp4_t *dp4, int *sb_qual);
/* the INFO string of a SNV record exactly as vcf_var_sprintf_info writes it (vcf.c:608-629):
 * "DP=%d;AF=%f;SB=%d;DP4=%d,%d,%d,%d;HQA=%d"; returns its length or -1 when buf is too small */
int lfb200_format_snv_info(char *buf, unsigned long size, int dp, float af, int sb, const lfb200_dp4_t *dp4, int hqa);
/* the INFO string of an indel record (vcf.c:608-629, indel branch): "DP=%d;AF=%f;SB=%d;DP4=%d,%d,%d,%d;INDEL;HRUN=%d" */
int lfb200_format_indel_info(char *buf, unsigned long size, int dp, float af, int sb, const lfb200_dp4_t *dp4, int hrun);
/* the record line as vcf_write_var writes it (vcf.c:469-495): CHROM POS(1-based) . REF ALT QUAL . INFO \n */
int lfb200_format_snv_record(char *buf, unsigned long size, const char *chrom, long pos0, char ref, char alt, int qual,
                             const char *info);
/* One reported variant = one iteration of the loop of call_snvs that passes `pvalue * bonf < sig`
 * (lofreq_call.c:818-871), with everything report_var() needs. */
typedef struct {
    long long tag;                     /* as given to lfb200_builder_add_column_strands */
    long long bonf;
    double lnp;
    float af;
    int qual, dp, sb, hqa;
    lfb200_dp4_t dp4;
    char ref_base, alt_base;
} lfb200_variant_t;
typedef void (*lfb200_variant_fn)(const lfb200_variant_t *v, void *user);
/* Column with its strand counts: fw_counts / rv_counts = plp_col_t.fw_counts / rv_counts (long int[NUM_NT4], the first
 * four entries A,C,G,T are read).  With a variant callback set, every flush runs the SB kernel over the called alleles of
 * the batch and reports them in input order, alleles in A,C,G,T-minus-ref order like call_snvs. */
int lfb200_builder_add_column_strands(lfb200_builder *bld, long long tag, char ref_base, int coverage_plp, int num_bases,
                                      const int *const base_quals[4], const int *const map_quals[4],
                                      const int *const baq_quals[4], const int *const source_quals[4], const int n[4],
                                      const long *fw_counts, const long *rv_counts);
int lfb200_builder_on_variant(lfb200_builder *bld, lfb200_variant_fn fn, void *user);

/* ---- link-compatible single-column symbols ----------------------------- */
/* snpcaller() of snpcaller.h:97-102 (callers: lofreq_call.c:319,384,807,
 * lofreq_uniq.c:311): same arguments, same sentinels, returns 0 on success.
 * A batch of one through the same kernels, on a lazily created default ctx.
 * approx_threshold_n > 0 is refused like a reference build without GSL
 * (snpcaller.c:1118-1125) but by returning 1 instead of exit(1). */
int lfb200_snpcaller(long double *snp_pvalues, const double *err_probs, const int num_err_probs,
                     const int *noncons_counts, const long long int bonf_factor,
                     const double sig_level, const int approx_threshold_n);
/* many independent snpcaller() problems in one launch: problem i owns
 * err_probs[ep_off[i] .. ep_off[i+1]), noncons_counts[3i..], bonf[i]; writes
 * snp_pvalues[3i..], lnp[3i..] (nullable), status[3i..] (nullable). */
int lfb200_snpcaller_batch(lfb200_ctx *ctx, long long n, const double *err_probs, const long long *ep_off,
                           const int *noncons_counts, const long long *bonf, double sig_level,
                           long double *snp_pvalues, double *lnp, unsigned char *status);

/* poissbin() of snpcaller.h:93-96 (snpcaller.c:1019-1062; called by snpcaller() and by source_qual, plp.c:554):
 * same arguments; returns a malloc()ed row of num_failures+1 natural-log probabilities the caller free()s —
 * row[k < K] = ln P(k errors), row[K] = ln P(>= K errors) among the reads seen when pruned_calc_prob_dist
 * (snpcaller.c:830-971) stopped, i.e. a partial row when its Bonferroni-aware early exit fired — and
 * *pvalue = expl(row[K]) with the reference's LDBL_MIN / LDBL_MAX clamp.  NULL on error (the reference: on
 * failed malloc).  The reads are walked in the caller's order with the reference's log-space recurrence, one CTA
 * per problem.  num_failures must be >= 1; for num_failures > num_err_probs the reference leaves row[K]
 * uninitialised, here it is -1e100 (LOGZERO). */
double *lfb200_poissbin(long double *pvalue, const double *err_probs, const int num_err_probs, const int num_failures,
                        const long long int bonf, const double sig);
/* n problems at once: problem i has err_probs[ep_off[i] .. ep_off[i+1]), num_failures[i], bonf[i]; its row goes to
 * rows[row_off[i] .. row_off[i] + num_failures[i]], row_off[n] = total.  n_end (optional): the read at which the
 * recurrence stopped (= the number of reads when it was not pruned). */
int lfb200_poissbin_batch(lfb200_ctx *ctx, long long n, const double *err_probs, const long long *ep_off,
                          const int *num_failures, const long long *bonf, double sig, const long long *row_off,
                          double *rows, long double *pvalues, int *n_end);

/* plp_to_errprobs() of snpcaller.h:72-75 (snpcaller.c:345-498) for every column of a host batch: the merged error
 * probabilities of the reads that pass the filters, in pileup order (A, C, G, T groups, the order the reference
 * walks them), column c at err_probs[col_off[c] ..] (so err_probs needs col_off[n_cols] doubles), their number in
 * num_err_probs[c], and alt_bases / alt_counts / alt_raw_counts[3c .. 3c+2] as the reference fills them.
 * Columns whose ref_base is not A/C/G/T get num_err_probs 0 and zero counts (call_snvs never gets that far:
 * lofreq_call.c:754,892). */
int lfb200_batch_errprobs(lfb200_ctx *ctx, const lfb200_conf_t *conf, const lfb200_batch_t *host_batch, double *err_probs,
                          int *num_err_probs, int *alt_bases, int *alt_counts, int *alt_raw_counts);
/* The part of plp_col_t (plp.h:73-145) plp_to_errprobs reads: per nt4 index A,C,G,T the int_varray_t data pointers and
 * lengths of base_quals / map_quals / baq_quals / source_quals (pointer NULL = that varray is empty). */
typedef struct {
    char ref_base;
    int coverage_plp;
    int n[4];
    const int *base_quals[4], *map_quals[4], *baq_quals[4], *source_quals[4];
} lfb200_plp_col_t;
/* Same contract as the reference function: *err_probs is malloc()ed with room for coverage_plp doubles and the caller
 * free()s it (lofreq_call.c:778,878); the three int arrays have 3 slots; on failure a FATAL line goes to stderr and
 * *err_probs is NULL (snpcaller.c:354-359).  A batch of one on the default ctx. */
void lfb200_plp_to_errprobs(double **err_probs, int *num_err_probs, int *alt_bases, int *alt_counts, int *alt_raw_counts,
                            const lfb200_plp_col_t *p, const lfb200_conf_t *conf);

/* ---- indel tests: call_indels -> plp_to_ins_errprobs / plp_to_del_errprobs -> snpcaller ------
 * One test per indel event of a column (lofreq_call.c:618-726).  Its error probabilities are, for every read of the
 * column, merge_srcq_mapq_baq_and_bq(sq, mq, aq, iq) (snpcaller.c:501-623): iq = the read's insertion / deletion
 * quality, mq its mapping quality, aq its indel alignment quality — only for the reads of the event under test
 * (snpcaller.c:541,597) — and sq its source quality (event reads only); the count handed to snpcaller() is
 * (event count, 0, 0) with bonf_indel (lofreq_call.c:305-320, 360-385).
 * Test t owns reads [read_off[t], read_off[t+1]) of the planes, the reads of its event LAST: event_count[t] of them.
 * Byte 255 = quality not available (aq of every read that is not of the event; sq of non-indel reads; mq 255).
 * mq / aq / sq may be NULL; conf->flag selects MQ / IDAQ / SQ as in the reference; conf->sig is the level.
 * Outputs per test (any may be NULL except pvalues): pvalues = snp_pvalues[0], lnp, status (LFB200_ST_*),
 * called = pvalue * bonf < sig (lofreq_call.c:326,392), qual = PROB_TO_PHREDQUAL(pvalue) where called, else -1.
 * Everything between the quality bytes and ln p stays on the device (k_errprobs -> k_prob_jobs). */
int lfb200_indel_tests(lfb200_ctx *ctx, const lfb200_conf_t *conf, long long n_tests, const long long *read_off,
                       const unsigned char *iq, const unsigned char *mq, const unsigned char *aq, const unsigned char *sq,
                       const int *event_count, const long long *bonf_indel, long double *pvalues, double *lnp,
                       unsigned char *status, unsigned char *called, int *qual);

/* binom() of binom.c:52-93 (caller: lofreq_uniq.c:381; note binom.h:32 names the
 * last two arguments the other way round): *p = P(X <= num_success), *q = 1 - *p
 * for X ~ Binomial(num_trials, prob_success); either pointer may be NULL.
 * Returns cdfbin's status (dcdflib.c:1860-1960): 0 ok, -5 num_trials <= 0,
 * -4 num_success outside [0, num_trials], -6 prob_success outside [0, 1]; on a
 * non-zero status *p and *q are left untouched, like the reference.  The
 * reference reaches the value through cdflib's incomplete beta function
 * (cumbin -> cumbet -> bratio); the device sums the mass function on the short
 * side of the mode (csrc/binom.cu).  Same default context as lfb200_snpcaller. */
int lfb200_binom(double *p, double *q, int num_trials, int num_success, double prob_success);
/* many binom() calls in one launch; status[i] as above (required), p / q nullable.
 * Returns non-zero only for CUDA / memory failures (lfb200_last_error). */
int lfb200_binom_batch(lfb200_ctx *ctx, long long n, const int *num_trials, const int *num_success,
                       const double *prob_success, double *p, double *q, int *status);

/* ---- synthetic pileup columns on the device (benchmark input,
 *      bit-identical to oracle/synth_np.py) --------------- */
/* workload: 2..5 = C2..C5.  Writes columns [c0, c0+n_cols) with the given
 * per-column pitch into device buffers; col_off must already hold the
 * offsets (n_cols+1).  baq may be NULL. */
int lfb200_synth_depths(int workload, long long c0, long long n_cols, int *depth_dev, void *stream);
int lfb200_synth_columns(int workload, long long c0, long long n_cols, const long long *col_off_dev,
                         int *nt_cnt_dev, char *ref_base_dev, unsigned char *bq_dev, unsigned char *mq_dev,
                         unsigned char *baq_dev, void *stream);

#ifdef __cplusplus
}
#endif
#endif
