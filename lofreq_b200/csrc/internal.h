// Internal declarations shared by the CUDA translation units and the host side
// of liblofreq_b200.so.  Not part of the public ABI (that is include/lofreq_b200.h).
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

namespace lfb {

constexpr int KS = 8;              // columns whose largest alt count is <= KS are finished by the screen kernel
constexpr int NCLASS = 10;         // job lists: 0 = K <= 32 (k_mid), 1..6 = register tiles R = 2..64 (k_heavy<R>), 7 = XL (CTA per
                                   // column), 8 = K <= 32 columns k_mid hands back to k_heavy<1> (tail outside the untilted range)
constexpr int CLS_XL = 7, CLS_FALLBACK = 8;
constexpr int CLS_XLFB = 1;         // K > 2048 columns k_xl hands to k_heavy_xl (a step parameter above 2^20: rescaling after every read)
constexpr int CLS_PRUNE2 = 9;       // K <= 8 columns still alive after the first reads of k_front's prune: k_prune2's input
constexpr int MAXK_WARP = 2048;    // 32 lanes * 64 cells

// k_dp (dp_fused.cu): columns with 8 < K <= 2048 share a warp — G = 4, 8, 16 or 32 lanes per column, R = 8 .. 64 cells
// per lane (class), all columns of a warp advancing read by read in lock step.  Job lists per (class, depth bin) so
// that the columns of a warp have similar depths; a binned list that overflows spills into the class's unbinned list.
constexpr int DP_NCLS = 7;         // (G, R): (4,8) (8,8) (16,8) (32,8) (32,16) (32,32) (32,64) -> K <= 32 .. 2048
constexpr int DP_NBIN = 14;        // depth bins: <= 128, 256, ..., 2^20 reads (defaults.h:60: max depth 1e6)
constexpr int DP_NBIN1 = DP_NBIN + 1;   // + the unbinned overflow list of the class
constexpr int DP_NL = DP_NCLS * DP_NBIN1;
constexpr int DP_MAXK = 2048;
constexpr int FRONT_MAXROUNDS = 32;  // round counters of k_front (later rounds count nothing: the prune's factor stays a lower bound)

// what the kernels need from varcall_conf_t, pre-digested on the host
struct DevConf {
    int min_bq, min_alt_bq;
    int alt_bq_mode;               // 0 keep, 1 constant (def_alt_bq > 0), 2 median of the ref-base qualities (def_alt_bq == -1)
    double alt_bq_prob;            // pow(10, -def_alt_bq/10) for mode 1
    int min_cov;
    int use_mq, use_baq, use_sq;   // flag bit set AND plane present
    int jq_filters;                // min_jq or min_alt_jq can reject a read
    double skip_jp, skip_alt_jp;   // a read is dropped when its merged probability is >= this (see host_api.cpp: jq_cut)
    int def_alt_jq_on;
    double def_alt_jq_prob;
    double sig;                    // (double)conf->sig, as call_snvs passes it (lofreq_call.c:807)
    double ln_sig;                 // glibc log(sig): k_emit_sites decides `pvalue * bonf < sig` on ln p (lofreq_call.c:832)
    int qual_ldblmin;              // PROB_TO_PHREDQUAL(LDBL_MIN) as the host evaluates it (utils.h:45)
    int bonf_dynamic;
    long long bonf_start;          // conf->bonf_subst on entry
};

struct DevBatch {
    long long n_cols;
    const long long *col_off;
    const int *nt_cnt;
    const char *ref_base;
    const int *coverage;
    const unsigned char *bq, *mq, *baq, *sq;
    const int *num_bases;
};

// pow(10,-q/10) tables built on the host with glibc pow() so that the device
// works on bit-identical probabilities (utils.h:42)
struct Lut {
    double bq[256];                // plain phred
    double mq[256];                // [0] = 0.5 (MQ0_ERRPROB, snpcaller.c:64), [255] = 0 (unknown -> -1 -> prob 0)
    double aq[256];                // baq / sq: [255] = 0 (-1: not available)
    // 1 / (1 - bq[i]), 1 / (1 - mq[i]) and 1 / (1 - aq[i]) (inf where the probability is 1): 1 / (1 - merged probability) of a (bq, mq) pair
    // is their product — the step parameters of k_dp without a division (read from global memory: not part of the 768
    // doubles the kernels stage in shared memory)
    double rbq[256];
    double rmq[256];
    double raq[256];
};

// a column (or a stand-alone snpcaller problem) the device could not rule out
struct Cand {
    long long col;
    long long bonf;
    double lnp[3];                 // ln P(X >= count_i)
    double ln_floor;               // min(ln P(X = K-1), ln P(X >= K)), K = max count: input of the clamp rule
    int cnt[3];
    int raw[3];
    int flags;
    int pad;
};
enum { CF_INSIG = 1, CF_RANGE = 2, CF_UNSUPPORTED = 4 };

// A site as the host sees it: byte image of lfb200_site_t (include/lofreq_b200.h) on x86-64, written by k_emit_sites
// straight into pinned host memory in column order.  The long double p-values are the host's business (x87).
struct alignas(16) SiteRec {
    long long col, bonf;
    double lnp[3];
    double ln_floor;
    unsigned long long pvalue_bits[6];   // long double pvalue[3]: zeroed here, filled by the host on request
    int cnt[3], raw[3], qual[3];
    unsigned char status[3], called[3];
    unsigned char flags;           // SITE_NEEDS_HOST: a comparison fell inside its guard band, the host repeats it in long double
    unsigned char pad;
};
static_assert(sizeof(SiteRec) == 144, "SiteRec must mirror lfb200_site_t");
enum { SITE_NEEDS_HOST = 1, SITE_UNSUPPORTED = 2 };
constexpr int EMIT_FIX_MAX = 60;

struct Counters {
    unsigned long long n_tested;   // written by the block-sum scan
    unsigned int n_cand;
    unsigned int n_jobs[NCLASS];
    unsigned int next_job[NCLASS];
    unsigned int err_flags;
    unsigned int n_pjobs[DP_NL];   // k_dp job lists, index class * DP_NBIN1 + bin (a binned count may exceed Workspace::pcap:
                                   // the excess went to the class's unbinned list, bin == DP_NBIN)
    unsigned int next_ptask[3];    // k_dp<RC>
    long long bonf_start_used;     // running factor the last test started from (host or device supplied)
    unsigned int front_round[FRONT_MAXROUNDS];   // tested columns counted so far per round of k_front (lower bounds for its prune)
    // written by k_emit_sites
    unsigned int n_fix;            // sites whose decision the host must repeat (may exceed EMIT_FIX_MAX: then all are rechecked)
    unsigned int emit_overflow;    // more sites than the host buffer holds: the host grows it and emits again
    unsigned int n_unsupported;    // columns with an alt count no kernel of this build takes
    unsigned int fix[EMIT_FIX_MAX];
};

struct Workspace {
    long long cap_cols;
    int *cnt6;                     // [n][6]: alt_counts[3], alt_raw_counts[3]
    unsigned char *tested;         // [n]
    unsigned char *rank;           // [n]: 1-based rank among the tested columns of its warp's 32 columns, 0 = untested (k_front);
                                   // rank in the batch = blocksum[tile] + wcount[warps before it in the tile] + rank (col_rank)
    unsigned char *wcount;         // [ceil(n/256)][8]: tested columns per warp of a tile (k_front)
    long long *bonf_used;          // [n]: materialised on request (k_bonf_used)
    long long *blocksum;           // [ceil(n/256)]: tested columns before each tile (k_scan_tiles)
    int *jobs;                     // [NCLASS][n]
    Cand *cand;                    // [n]
    unsigned char *is_cand;        // [n rounded up to 256]: column emitted a candidate
    unsigned int *candtile;        // [ceil(n/256)]: candidates per tile, zeroed again by the scan
    long long *candpre;            // [ceil(n/256)]: candidates before each tile
    int *cand_perm;                // [n]: cand_perm[rank in column order] = index into cand
    Counters *counters;
    int *pjobs;                    // [DP_NCLS * DP_NBIN][pcap]
    long long pcap;
    int *ujobs;                    // [DP_NCLS][n]: unbinned overflow lists
};

// stand-alone snpcaller problems (link-compatible path)
struct ProbBatch {
    long long n;
    const double *err_probs;
    const long long *ep_off;
    const int *counts;             // [n][3]
    const long long *bonf;         // [n]
    double sig;
};

// per-context launch state: side streams and fork/join events of the job classes, SM count of the context's device.
// Owned by lfb200_ctx (created after cudaSetDevice), so that contexts on different devices or driven by different
// host threads never share a stream or an event.
constexpr int NSIDE = 8;
struct LaunchState {
    int sms = 148;
    cudaStream_t side[NSIDE] = {};
    cudaEvent_t ev_fork = nullptr, ev_fork2 = nullptr, ev_join[NSIDE] = {};
    bool ready = false;
    // which of k_mid (1), k_dp<1> (2), k_xl (4), k_dp<2> (8) get a side stream for the next test: those whose job list
    // was not empty in this context's previous batch (everything, before the first batch)
    unsigned side_mask = 0xfu;
};
int launch_state_init(LaunchState &ls, int device);
void launch_state_destroy(LaunchState &ls);

// launchers (snv_kernels.cu)
// front.cu: gates, alt counts, running count, first stage of the prune, job lists — one pass
void launch_front(const LaunchState &ls, const DevConf &cf, const DevBatch &b, const Lut *lut, const Workspace &ws, cudaStream_t st);
void launch_bonf_used(const LaunchState &ls, const DevConf &cf, const Workspace &ws, long long n, cudaStream_t st);
void launch_test(const LaunchState &ls, const DevConf &cf, const DevBatch &b, const Lut *lut, const Workspace &ws, cudaStream_t st,
                 cudaEvent_t after_finalize, cudaEvent_t after_heavy, const long long *bonf_start_dev);
// sites in column order, decided on the device, into (mapped pinned) host memory
void launch_emit_sites(const LaunchState &ls, const DevConf &cf, const Workspace &ws, SiteRec *out, unsigned cap, cudaStream_t st);
void launch_bonf_start(const long long *counts, int rank, long long bonf_subst, long long *start, cudaStream_t st);
void launch_bonf_start_strided(const long long *counts, int stride, int rank, long long bonf_subst, long long *start,
                               cudaStream_t st);
void launch_set_i64(long long *dst, long long v, cudaStream_t st);
double measure_dfma_per_second(int sms, cudaStream_t st);
// dp_fused.cu
int dp_smem_optin();
void launch_dp(const LaunchState &ls, const DevConf &cf, const DevBatch &b, const Lut *lut, const Workspace &ws, cudaStream_t st,
               cudaStream_t st1, cudaStream_t st2);
void launch_prob_jobs(int sms, const ProbBatch &pb, Cand *out, cudaStream_t st);
// xl.cu: K > 2048, one CTA per column, the row spread over 8 warps running as a wavefront over the reads
int xl_smem_optin();
void launch_xl(const LaunchState &ls, const DevConf &cf, const DevBatch &b, const Lut *lut, const Workspace &ws, cudaStream_t st);
// mailbox.cu: the per-batch count exchange between shards through shared host memory
constexpr int MAIL_DEPTH = 64;
struct MailSlot {
    unsigned long long seq;        // written last (release): the exchange this slot holds
    long long tested, sites, pad;
};
void launch_mail_exchange(MailSlot *slots, unsigned long long *ack, int world, int rank, unsigned long long seq,
                          const unsigned long long *n_tested_dev, long long sites_prev, long long bonf_subst, long long *d_mine,
                          long long *d_start, int *d_err, cudaStream_t st);
// poissbin.cu
void launch_poissbin_rows(const ProbBatch &pb, const int *num_failures, const long long *row_off, double *buf, double *rows,
                          int *n_end, cudaStream_t st);
void launch_errprobs(const DevConf &cf, const DevBatch &b, const Lut *lut, double *ep_out, int *n_out, int *counts9, cudaStream_t st);
// binom.cu
int launch_binom(long long n_prob, const int *num_trials, const int *num_success, const double *prob, double *cum, double *ccum,
                 int *status, cudaStream_t st);
// fisher.cu: SB = PROB_TO_PHREDQUAL_SAFE(Fisher two-tailed p) per DP4 table (lofreq_call.c:108-125, fet.c:62-101)
void launch_sb_qual(int sms, const int *dp4, long long n, int *sb, unsigned char *unsure, cudaStream_t st);
// BAQ HMM (baq.cu): reads [r0, r0 + n_reads) of the batch; scratch for ceil(n_reads / 32) warps of `rows` forward rows each
struct KpaFix;
void launch_kpa_glocal(long long r0, int n_reads, const unsigned char *ref, const long long *ref_off, const unsigned char *query,
                       const long long *qry_off, const unsigned char *qual, float d, float e, int bw, const float *q2p, double *f, double *b,
                       double *s, int w3, int rows, int *state, unsigned char *q, KpaFix *fix, int fix_cap, unsigned *n_fix, cudaStream_t st);
int sb_qual_host(const int *dp4);
// synth.cu
void launch_synth_depths(int workload, long long c0, long long n, int *depth, cudaStream_t st);
void launch_synth_columns(int workload, long long c0, long long n, const long long *col_off, int *nt_cnt, char *ref,
                          unsigned char *bq, unsigned char *mq, unsigned char *baq, cudaStream_t st);

}  // namespace lfb
