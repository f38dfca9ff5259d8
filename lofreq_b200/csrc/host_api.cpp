// Host side of liblofreq_b200.so: context and workspace management, the C ABI of
// include/lofreq_b200.h, and the long double finishing of the few columns the device keeps
// ("sites").  No SNV arithmetic happens here except what needs x87 long double to decide exactly
// like the reference: expl(), the FE-exception clamp, pvalue*bonf < sig, PROB_TO_PHREDQUAL.
#include <algorithm>
#include <cerrno>
#include <chrono>
#include <cfenv>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdarg>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <utility>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/lofreq_b200.h"
#include "internal.h"
#include "baq_core.cuh"       // KpaFix (the routine itself is only ever instantiated for the device, baq.cu)

using namespace lfb;

void lfb_comm_release(lfb200_ctx *ctx);      // shard_comm.cpp
int lfb_comm_error(lfb200_ctx *ctx);

// ------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int fail(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return 1;
}

#define CU(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess) return fail("CUDA error %s at %s:%d", cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

extern "C" const char *lfb200_last_error(void) { return g_err; }

struct DevBuf;
static int upload(DevBuf &buf, const void *src, size_t bytes, size_t pad, cudaStream_t st, const void **dev);

// ------------------------------------------------------------------------------------------------
// A few persistent worker threads for the long double finishing (FE flags and errno are per thread).
class WorkerPool {
public:
    explicit WorkerPool(unsigned n) : stop_(false), gen_(0), pending_(0)
    {
        for (unsigned i = 0; i < n; ++i) threads_.emplace_back([this, i] { loop(i); });
    }
    ~WorkerPool()
    {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
            ++gen_;
        }
        cv_.notify_all();
        for (auto &t : threads_) t.join();
    }
    unsigned size() const { return (unsigned)threads_.size(); }
    // runs fn(part, nparts) on every worker and on the caller; returns when all are done
    void run(const std::function<void(unsigned, unsigned)> &fn)
    {
        const unsigned parts = size() + 1;
        {
            std::lock_guard<std::mutex> lk(m_);
            fn_ = &fn;
            pending_ = size();
            ++gen_;
        }
        cv_.notify_all();
        fn(parts - 1, parts);
        std::unique_lock<std::mutex> lk(m_);
        done_.wait(lk, [this] { return pending_ == 0; });
    }

private:
    void loop(unsigned id)
    {
        unsigned long long seen = 0;
        for (;;) {
            const std::function<void(unsigned, unsigned)> *fn;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
                fn = fn_;
            }
            (*fn)(id, size() + 1);
            {
                std::lock_guard<std::mutex> lk(m_);
                if (--pending_ == 0) done_.notify_one();
            }
        }
    }
    std::vector<std::thread> threads_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    const std::function<void(unsigned, unsigned)> *fn_ = nullptr;
    bool stop_;
    unsigned long long gen_;
    unsigned pending_;
};

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes)
    {
        if (bytes <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        if (cudaMalloc(&p, want) != cudaSuccess) {
            cudaGetLastError();
            return 1;
        }
        cap = want;
        return 0;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

// The launches of a phase (screen: 4, test: ~20 with the fork / join events of the side streams) as one CUDA graph.  A
// pipeline that sends batches of one shape through the same buffers (a resident batch timed again and again; a builder
// that refills its planes) repeats the same launch sequence byte for byte: the second time in a row a sequence is seen it
// is captured, from then on replayed with one cudaGraphLaunch — host time per batch is what limits 8 processes sharing
// one box's cores.  Anything else (a new configuration, a running Bonferroni factor passed by value, grown buffers)
// simply takes the plain launches.
struct GraphCache {
    std::vector<unsigned char> key, seen;
    cudaGraphExec_t exec = nullptr;
    bool disabled = false;
    long long replays = 0;
    void release()
    {
        if (exec) cudaGraphExecDestroy(exec);
        exec = nullptr;
        key.clear();
        seen.clear();
    }
};

struct KeyBuilder {
    std::vector<unsigned char> v;
    template <class T> void add(const T &x)
    {
        const unsigned char *p = reinterpret_cast<const unsigned char *>(&x);
        v.insert(v.end(), p, p + sizeof(T));
    }
};

template <class F> static int run_graphed(GraphCache &gc, bool allowed, const std::vector<unsigned char> &key, cudaStream_t st, F &&enqueue)
{
    static const bool off = getenv("LFB200_NO_GRAPH") != nullptr;
    if (off || !allowed || gc.disabled || st == nullptr || st == cudaStreamLegacy || st == cudaStreamPerThread) {
        enqueue();
        return 0;
    }
    if (gc.exec && key == gc.key) {
        if (cudaGraphLaunch(gc.exec, st) == cudaSuccess) { ++gc.replays; return 0; }
        cudaGetLastError();
        gc.disabled = true;
        enqueue();
        return 0;
    }
    if (key != gc.seen) {
        gc.seen = key;
        enqueue();
        return 0;
    }
    // the same sequence twice in a row: capture it (nothing runs during the capture), then launch what was captured
    cudaGraph_t g = nullptr;
    if (cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed) != cudaSuccess) {
        cudaGetLastError();
        gc.disabled = true;
        enqueue();
        return 0;
    }
    enqueue();
    cudaError_t rc = cudaStreamEndCapture(st, &g);
    cudaGraphExec_t ex = nullptr;
    if (rc == cudaSuccess && g) rc = cudaGraphInstantiate(&ex, g, 0);
    if (g) cudaGraphDestroy(g);
    if (rc != cudaSuccess || !ex) {
        cudaGetLastError();
        gc.disabled = true;
        enqueue();
        return 0;
    }
    if (gc.exec) cudaGraphExecDestroy(gc.exec);
    gc.exec = ex;
    gc.key = key;
    if (cudaGraphLaunch(gc.exec, st) != cudaSuccess) {
        cudaGetLastError();
        gc.disabled = true;
        enqueue();
    } else {
        ++gc.replays;
    }
    return 0;
}

struct lfb200_ctx {
    int device = 0;
    GraphCache g_front, g_test;
    cudaStream_t stream = nullptr;       // used by the host entry points
    Lut *d_lut = nullptr;
    // workspace of the current batch
    DevBuf w_cnt6, w_tested, w_bonf, w_rank, w_wcount, w_blocksum, w_jobs, w_cand, w_counters, w_pjobs, w_ujobs, w_iscand, w_candtile, w_candpre, w_perm;
    Workspace ws{};
    // device copies of host batches (host entry point)
    DevBuf in_off, in_cnt, in_ref, in_cov, in_nb, in_bq, in_mq, in_baq, in_sq;
    // single problems
    DevBuf p_ep, p_off, p_cnt, p_bonf, p_out;
    // BAQ HMM (lfb200_kpa_glocal_batch)
    DevBuf k_ref, k_roff, k_qry, k_qoff, k_qual, k_state, k_q, k_f, k_b, k_s, k_fix, k_q2p;
    // pinned scratch for small D2H transfers
    Counters *h_counters = nullptr;
    Cand *h_cand = nullptr;              // pinned
    size_t h_cand_cap = 0;
    std::unique_ptr<WorkerPool> pool;
    int ensure_cand(size_t n)
    {
        if (n <= h_cand_cap) return 0;
        if (h_cand) cudaFreeHost(h_cand);
        h_cand = nullptr;
        h_cand_cap = 0;
        const size_t want = n + n / 4 + 1024;
        if (cudaMallocHost(&h_cand, want * sizeof(Cand)) != cudaSuccess) { cudaGetLastError(); return 1; }
        h_cand_cap = want;
        return 0;
    }
    // optional per-phase timing (lfb200_set_profiling): events on the launching stream
    bool profiling = false;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    // finisher thread (lfb200_sites_begin / lfb200_sites_end)
    std::thread fin_thread;
    std::mutex fin_m;
    std::condition_variable fin_cv, fin_done;
    int fin_state = 0;                   // 0 idle, 1 requested, 2 done
    bool fin_stop = false;
    lfb200_conf_t *fin_conf = nullptr;
    void *fin_stream = nullptr;
    lfb200_site_t *fin_sites = nullptr;
    long long fin_max = 0;
    lfb200_summary_t fin_summary{};
    int fin_rc = 0;
    char fin_err[512] = "";
    void finisher_loop();
    void *comm_state = nullptr;          // owned by shard_comm.cpp
    int host_planes = 1;                 // lfb200_set_host_planes: pinned planes are read in place unless the caller says otherwise
    LaunchState ls;                      // side streams / events / SM count of this context's device
    // sites of the last test, written by k_emit_sites in column order straight into mapped pinned memory
    lfb200_site_t *h_sites = nullptr;
    SiteRec *d_sites = nullptr;          // device view of h_sites
    size_t h_sites_cap = 0;
    cudaEvent_t ev_done = nullptr;       // recorded after the sites and the counters of a test have been written
    DevConf last_dc{};                   // configuration of the last test (to emit again after growing h_sites)
    bool test_enqueued = false;
    bool bonf_used_valid = false;        // ws.bonf_used has been materialised for the current batch
    int site_pvalues = 1;                // lfb200_set_site_pvalues
    int ensure_sites(size_t n)
    {
        if (n <= h_sites_cap) return 0;
        if (h_sites) cudaFreeHost(h_sites);
        h_sites = nullptr;
        d_sites = nullptr;
        h_sites_cap = 0;
        const size_t want = n + n / 4 + 1024;
        if (cudaHostAlloc((void **)&h_sites, want * sizeof(lfb200_site_t), cudaHostAllocMapped) != cudaSuccess) { cudaGetLastError(); return 1; }
        if (cudaHostGetDevicePointer((void **)&d_sites, h_sites, 0) != cudaSuccess) { cudaGetLastError(); return 1; }
        h_sites_cap = want;
        return 0;
    }
    // state of the last screen
    DevBatch cur{};
    bool have_batch = false;
    long long n_tested = -1;
};

static void build_lut(Lut &l)
{
    for (int q = 0; q < 256; ++q) {
        const double p = pow(10.0, -1.0 * q / 10.0);     // PHREDQUAL_TO_PROB, utils.h:42
        l.bq[q] = p;
        l.mq[q] = p;
        l.aq[q] = p;
    }
    l.mq[0] = 0.5;      // MQ0_ERRPROB, snpcaller.c:64,315
    l.mq[255] = 0.0;    // unknown -> -1 -> probability 0, snpcaller.c:451-453,311
    l.aq[255] = 0.0;    // -1: quality not available (plp.c:961)
    for (int q = 0; q < 256; ++q) {
        l.rbq[q] = 1.0 / (1.0 - l.bq[q]);
        l.rmq[q] = 1.0 / (1.0 - l.mq[q]);
        l.raq[q] = 1.0 / (1.0 - l.aq[q]);
    }
}

// PROB_TO_PHREDQUAL_SAFE (utils.h:46)
static int prob_to_phred_safe(double p)
{
    if (p <= 0.0) return INT_MAX;
    return (int)(-10.0 * log10l(p));
}

// Smallest double jp for which the reference's filter `merged_qual < min_q` (snpcaller.c:469,480) fires,
// found by bisection on the bit pattern with the reference's own expression, so that the device can
// filter with one comparison and still agree at every integer boundary of log10l.
static double jq_cut(int min_q)
{
    if (min_q <= 0) return INFINITY;          // merged_qual >= 0 always (jp <= 1)
    if (!(prob_to_phred_safe(1.0) < min_q)) return INFINITY;
    unsigned long long lo = 1, hi;            // lo: predicate false (tiny jp, huge quality); hi: true
    double one = 1.0;
    memcpy(&hi, &one, 8);
    double dlo;
    memcpy(&dlo, &lo, 8);
    if (prob_to_phred_safe(dlo) < min_q) return dlo;
    while (hi - lo > 1) {
        const unsigned long long mid = lo + (hi - lo) / 2;
        double d;
        memcpy(&d, &mid, 8);
        if (prob_to_phred_safe(d) < min_q) hi = mid; else lo = mid;
    }
    double d;
    memcpy(&d, &hi, 8);
    return d;
}

static_assert(sizeof(lfb200_site_t) == sizeof(SiteRec), "k_emit_sites writes byte images of lfb200_site_t");
static_assert(offsetof(lfb200_site_t, lnp) == offsetof(SiteRec, lnp) && offsetof(lfb200_site_t, pvalue) == offsetof(SiteRec, pvalue_bits) &&
              offsetof(lfb200_site_t, alt_count) == offsetof(SiteRec, cnt) && offsetof(lfb200_site_t, alt_raw_count) == offsetof(SiteRec, raw) &&
              offsetof(lfb200_site_t, qual) == offsetof(SiteRec, qual) && offsetof(lfb200_site_t, status) == offsetof(SiteRec, status) &&
              offsetof(lfb200_site_t, called) == offsetof(SiteRec, called) && offsetof(lfb200_site_t, flags) == offsetof(SiteRec, flags) &&
              offsetof(lfb200_site_t, ln_floor) == offsetof(SiteRec, ln_floor), "lfb200_site_t / SiteRec layout");

static int make_devconf(const lfb200_conf_t *c, const lfb200_batch_t *b, DevConf &d)
{
    memset(&d, 0, sizeof(d));
    d.min_bq = c->min_bq;
    d.min_alt_bq = c->min_alt_bq;
    if (c->def_alt_bq == -1) d.alt_bq_mode = 2;
    else if (c->def_alt_bq != 0) {
        d.alt_bq_mode = 1;
        d.alt_bq_prob = pow(10.0, -1.0 * c->def_alt_bq / 10.0);
    }
    d.min_cov = c->min_cov;
    d.use_mq = (c->flag & LFB200_USE_MQ) && b->mq;
    d.use_baq = (c->flag & LFB200_USE_BAQ) && b->baq;
    d.use_sq = (c->flag & LFB200_USE_SQ) && b->sq;
    d.skip_jp = jq_cut(c->min_jq);
    d.skip_alt_jp = jq_cut(c->min_alt_jq);
    d.jq_filters = std::isfinite(d.skip_jp) || std::isfinite(d.skip_alt_jp);
    if (c->def_alt_jq == -1) return fail("def_alt_jq = -1 is not implemented (neither in the reference, snpcaller.c:482-484)");
    d.def_alt_jq_on = c->def_alt_jq != 0;
    d.def_alt_jq_prob = d.def_alt_jq_on ? pow(10.0, -1.0 * c->def_alt_jq / 10.0) : 0.0;
    d.sig = (double)c->sig;
    d.ln_sig = log(d.sig);
    d.qual_ldblmin = (int)(-10.0 * log10l(LDBL_MIN));      // PROB_TO_PHREDQUAL(LDBL_MIN), utils.h:45
    d.bonf_dynamic = c->bonf_dynamic;
    d.bonf_start = c->bonf_subst;
    return 0;
}

extern "C" void lfb200_init_conf(lfb200_conf_t *c)
{
    // init_varcall_conf, snpcaller.c:626-651 with defaults.h:40-80
    memset(c, 0, sizeof(*c));
    c->min_bq = 6;
    c->min_alt_bq = 6;
    c->def_alt_bq = 0;
    c->min_jq = 0;
    c->min_alt_jq = 0;
    c->def_alt_jq = 0;
    c->min_cov = 1;
    c->bonf_dynamic = 1;
    c->flag = LFB200_USE_MQ | LFB200_USE_BAQ | LFB200_USE_IDAQ;
    c->sig = 0.01f;
    c->bonf_subst = 1;
    c->num_snv_tests = 0;
}

extern "C" int lfb200_create(lfb200_ctx **out, int device)
{
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        cudaGetLastError();
        return fail("no CUDA device: lofreq_b200 has no CPU path");
    }
    if (device < 0 || device >= ndev) return fail("device %d out of range (%d visible)", device, ndev);
    CU(cudaSetDevice(device));
    lfb200_ctx *ctx = new lfb200_ctx();
    ctx->device = device;
    CU(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    Lut l;
    build_lut(l);
    CU(cudaMalloc(&ctx->d_lut, sizeof(Lut)));
    CU(cudaMemcpy(ctx->d_lut, &l, sizeof(Lut), cudaMemcpyHostToDevice));
    CU(cudaMallocHost(&ctx->h_counters, sizeof(Counters)));
    memset(ctx->h_counters, 0, sizeof(Counters));
    if (ctx->w_counters.ensure(sizeof(Counters))) return fail("out of device memory");
    // streams, events, SM count and shared-memory opt-ins belong to this context and this device
    if (launch_state_init(ctx->ls, device)) return fail("could not create the streams / events of the context");
    CU(cudaEventCreateWithFlags(&ctx->ev_done, cudaEventDisableTiming));
    *out = ctx;
    return 0;
}

extern "C" void lfb200_destroy(lfb200_ctx *ctx)
{
    if (!ctx) return;
    lfb_comm_release(ctx);
    if (ctx->fin_thread.joinable()) {
        {
            std::lock_guard<std::mutex> lk(ctx->fin_m);
            ctx->fin_stop = true;
        }
        ctx->fin_cv.notify_all();
        ctx->fin_thread.join();
    }
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    DevBuf *bufs[] = {&ctx->w_cnt6, &ctx->w_tested, &ctx->w_bonf, &ctx->w_rank, &ctx->w_wcount, &ctx->w_blocksum, &ctx->w_jobs,
                      &ctx->w_cand, &ctx->w_counters, &ctx->w_pjobs, &ctx->w_ujobs, &ctx->w_iscand, &ctx->w_candtile, &ctx->w_candpre, &ctx->w_perm, &ctx->in_off, &ctx->in_cnt, &ctx->in_ref, &ctx->in_cov, &ctx->in_nb,
                      &ctx->in_bq, &ctx->in_mq, &ctx->in_baq, &ctx->in_sq, &ctx->p_ep, &ctx->p_off, &ctx->p_cnt,
                      &ctx->p_bonf, &ctx->p_out, &ctx->k_ref, &ctx->k_roff, &ctx->k_qry, &ctx->k_qoff, &ctx->k_qual, &ctx->k_state, &ctx->k_q,
                      &ctx->k_f, &ctx->k_b, &ctx->k_s, &ctx->k_fix, &ctx->k_q2p};
    for (DevBuf *b : bufs) b->release();
    if (ctx->d_lut) cudaFree(ctx->d_lut);
    if (ctx->h_counters) cudaFreeHost(ctx->h_counters);
    if (ctx->h_cand) cudaFreeHost(ctx->h_cand);
    if (ctx->h_sites) cudaFreeHost(ctx->h_sites);
    if (ctx->ev_done) cudaEventDestroy(ctx->ev_done);
    ctx->g_front.release();
    ctx->g_test.release();
    launch_state_destroy(ctx->ls);
    ctx->pool.reset();
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

static int ensure_workspace(lfb200_ctx *ctx, long long n)
{
    const size_t nn = (size_t)std::max<long long>(n, 1);
    int bad = 0;
    bad |= ctx->w_cnt6.ensure(nn * 6 * sizeof(int));
    bad |= ctx->w_tested.ensure(nn + 1024);
    bad |= ctx->w_bonf.ensure(nn * sizeof(long long));
    bad |= ctx->w_blocksum.ensure(((nn + 255) / 256 + 1) * sizeof(long long));
    bad |= ctx->w_rank.ensure(nn);
    bad |= ctx->w_wcount.ensure(((nn + 255) / 256 + 1) * 8);
    bad |= ctx->w_jobs.ensure(nn * NCLASS * sizeof(int));
    bad |= ctx->w_cand.ensure(nn * sizeof(Cand));
    bad |= ctx->w_iscand.ensure(((nn + 255) / 256 + 1) * 256);
    bad |= ctx->w_candpre.ensure(((nn + 255) / 256 + 1) * sizeof(long long));
    bad |= ctx->w_perm.ensure(nn * sizeof(int));
    {
        const void *before = ctx->w_candtile.p;
        bad |= ctx->w_candtile.ensure(((nn + 255) / 256 + 1) * sizeof(unsigned int));
        if (!bad && ctx->w_candtile.p != before) cudaMemset(ctx->w_candtile.p, 0, ctx->w_candtile.cap);   // the scan keeps it zero afterwards
    }
    // k_dp job lists: a binned list holds an eighth of the batch (a fuller one spills into its class's unbinned list,
    // which holds the whole batch)
    const size_t pcap = std::max<size_t>(nn / 8, 4096);
    bad |= ctx->w_pjobs.ensure(pcap * DP_NCLS * DP_NBIN * sizeof(int));
    bad |= ctx->w_ujobs.ensure(nn * DP_NCLS * sizeof(int));
    if (bad) return fail("out of device memory for a batch of %lld columns", n);
    Workspace &w = ctx->ws;
    w.cap_cols = n;
    w.cnt6 = (int *)ctx->w_cnt6.p;
    w.tested = (unsigned char *)ctx->w_tested.p;
    w.bonf_used = (long long *)ctx->w_bonf.p;
    w.blocksum = (long long *)ctx->w_blocksum.p;
    w.rank = (unsigned char *)ctx->w_rank.p;
    w.wcount = (unsigned char *)ctx->w_wcount.p;
    w.jobs = (int *)ctx->w_jobs.p;
    w.cand = (Cand *)ctx->w_cand.p;
    w.is_cand = (unsigned char *)ctx->w_iscand.p;
    w.candtile = (unsigned int *)ctx->w_candtile.p;
    w.candpre = (long long *)ctx->w_candpre.p;
    w.cand_perm = (int *)ctx->w_perm.p;
    w.counters = (Counters *)ctx->w_counters.p;
    w.pjobs = (int *)ctx->w_pjobs.p;
    w.pcap = (long long)pcap;
    w.ujobs = (int *)ctx->w_ujobs.p;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// long double finishing of one site: snpcaller.c:1144-1196 + lofreq_call.c:832,863
// ------------------------------------------------------------------------------------------------
static const double LN_EXP_UNDERFLOW = 708.3964185322641;   // glibc exp() raises FE_UNDERFLOW below -this

// expl() with the reference's clamp (snpcaller.c:1047-1059, 1169-1188); `pre_flag` = an exp() inside the
// preceding probvec_tailsum would already have raised FE_UNDERFLOW.
// The reference tests errno and the FE flags after expl().  For an argument <= 0 the only exception expl can
// raise is underflow, which IEEE 754 signals exactly when the (inexact) result is below LDBL_MIN — so the test
// is done on the result, which costs a third of the fenv round trip.  LFB200_STRICT_FENV=1 switches to the
// literal feclearexcept/fetestexcept sequence (tests/test_parity_gpu.py runs the golden grid both ways).
static const bool g_strict_fenv = getenv("LFB200_STRICT_FENV") != nullptr;

static long double expl_clamped(double t, bool pre_flag)
{
    long double p;
    bool flagged = pre_flag;
    if (g_strict_fenv) {
        errno = 0;
        feclearexcept(FE_ALL_EXCEPT);
        p = expl((long double)t);
        flagged = flagged || errno || fetestexcept(FE_INVALID | FE_DIVBYZERO | FE_OVERFLOW | FE_UNDERFLOW);
        errno = 0;
        feclearexcept(FE_ALL_EXCEPT);
    } else {
        p = expl((long double)t);
        flagged = flagged || p < LDBL_MIN || !(p <= LDBL_MAX);
    }
    if (flagged) p = (p < DBL_EPSILON) ? LDBL_MIN : LDBL_MAX;
    return p;
}

// PROB_TO_PHREDQUAL(p) = (int)(-10 * log10l(p)) (utils.h:45) for p = expl(t).  Only the integer part is used, and
// -10 t / ln 10 in double is within |q| * 5e-16 of the long double expression (expl and log10l are each good to
// ~1e-19 relative), so whenever its fractional part is further than 1e-7 from an integer the truncation is decided
// and log10l — a third of the finishing time — is not needed.  Sentinels and near-integer values take the literal path.
static int phredqual_of(long double p, double t)
{
    if (p != LDBL_MIN && p != LDBL_MAX && t < 0.0 && t > -11000.0) {
        const double q = t * -4.3429448190325182765;       // -10 / ln 10
        const double f = q - floor(q);
        if (f > 1e-7 && f < 1.0 - 1e-7) return (int)q;
    }
    return (int)(-10.0 * log10l(p));
}

static void finish_site(const Cand &cd, double sig, lfb200_site_t &s)
{
    s.col = cd.col;
    s.bonf = cd.bonf;
    s.flags = 0;
    s.reserved = 0;
    s.ln_floor = cd.ln_floor;
    int K = 0, imax = 0;
    for (int i = 0; i < 3; ++i) {
        s.lnp[i] = cd.lnp[i];
        s.pvalue[i] = LDBL_MAX;
        s.alt_count[i] = cd.cnt[i];
        s.alt_raw_count[i] = cd.raw[i];
        s.qual[i] = -1;
        s.status[i] = LFB200_ST_LDBLMAX;
        s.called[i] = 0;
        if (cd.cnt[i] > K) { K = cd.cnt[i]; imax = i; }
    }
    if (K == 0 || (cd.flags & CF_INSIG)) return;
    // poissbin: pvalue = expl(probvec[K]) with clamp, then the significance gate (snpcaller.c:1155)
    const long double pK = expl_clamped(cd.lnp[imax], false);
    if (pK * (double)cd.bonf > sig) return;
    for (int i = 0; i < 3; ++i) {
        const int c = cd.cnt[i];
        if (c == 0) continue;
        const double t = cd.lnp[i];
        // probvec_tailsum folds row[c..K] with log_sum; its exp() underflows exactly when the running sum
        // and the next term are more than 708.396 nats apart.  The row is log-concave, so the widest gap is
        // against the last two entries, min(row[K-1], row[K]) = ln_floor.
        const bool pre = (c < K) && (t - cd.ln_floor > LN_EXP_UNDERFLOW);
        const long double p = (c == K) ? pK : expl_clamped(t, pre);      // c == K: the same expl() as poissbin's
        s.pvalue[i] = p;
        s.status[i] = (p == LDBL_MAX) ? LFB200_ST_LDBLMAX : (p == LDBL_MIN) ? LFB200_ST_LDBLMIN : LFB200_ST_VALUE;
        if (p * (double)cd.bonf < sig) {                 // lofreq_call.c:832
            s.called[i] = 1;
            s.qual[i] = phredqual_of(p, t);              // PROB_TO_PHREDQUAL, lofreq_call.c:863
        }
    }
}

// ------------------------------------------------------------------------------------------------
// device-resident batches
// ------------------------------------------------------------------------------------------------
static void to_devbatch(const lfb200_batch_t *b, DevBatch &d)
{
    d.n_cols = b->n_cols;
    d.col_off = b->col_off;
    d.nt_cnt = b->nt_cnt;
    d.ref_base = b->ref_base;
    d.coverage = b->coverage;
    d.bq = b->bq;
    d.mq = b->mq;
    d.baq = b->baq;
    d.sq = b->sq;
    d.num_bases = b->num_bases;
}

extern "C" int lfb200_screen_device(lfb200_ctx *ctx, const lfb200_conf_t *conf, const lfb200_batch_t *b, void *stream)
{
    if (!ctx) return fail("no context");
    if (b->n_cols < 0 || b->n_cols > 2000000000ll) return fail("n_cols %lld out of range", b->n_cols);
    if (!b->bq) return fail("the bq plane is required");
    {
        // the finisher of the previous batch reads this context's workspace: the next batch must wait for lfb200_sites_end
        std::lock_guard<std::mutex> lk(ctx->fin_m);
        if (ctx->fin_state != 0) return fail("a sites request is pending on this context: call lfb200_sites_end first (or use a second context)");
    }
    CU(cudaSetDevice(ctx->device));
    DevConf dc;
    memset(&dc, 0, sizeof(dc));
    if (make_devconf(conf, b, dc)) return 1;
    if (ensure_workspace(ctx, b->n_cols)) return 1;
    ctx->test_enqueued = false;
    to_devbatch(b, ctx->cur);
    ctx->have_batch = true;
    ctx->n_tested = -1;
    cudaStream_t st = (cudaStream_t)stream;
    ctx->bonf_used_valid = false;
    if (ctx->profiling) cudaEventRecord(ctx->ev[0], st);
    {
        KeyBuilder kb;
        kb.add(dc); kb.add(ctx->cur); kb.add(ctx->ws); kb.add(ctx->d_lut); kb.add(st);
        run_graphed(ctx->g_front, !ctx->profiling && ctx->cur.n_cols > 0, kb.v, st,
                    [&]() { launch_front(ctx->ls, dc, ctx->cur, ctx->d_lut, ctx->ws, st); });
    }
    if (ctx->profiling) cudaEventRecord(ctx->ev[1], st);
    if (ctx->profiling) cudaEventRecord(ctx->ev[2], st);
    CU(cudaGetLastError());
    return 0;
}

extern "C" int lfb200_ntested_device(lfb200_ctx *ctx, void *stream, long long *n_tested)
{
    if (!ctx || !ctx->have_batch) return fail("no screened batch");
    cudaStream_t st = (cudaStream_t)stream;
    if (ctx->cur.n_cols == 0) { *n_tested = ctx->n_tested = 0; return 0; }
    CU(cudaMemcpyAsync(ctx->h_counters, ctx->ws.counters, sizeof(Counters), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    *n_tested = ctx->n_tested = (long long)ctx->h_counters->n_tested;
    return 0;
}

static int test_device_impl(lfb200_ctx *ctx, const lfb200_conf_t *conf, void *stream, const long long *bonf_start_dev)
{
    if (!ctx || !ctx->have_batch) return fail("no screened batch");
    lfb200_batch_t hb;
    memset(&hb, 0, sizeof(hb));
    hb.mq = ctx->cur.mq; hb.baq = ctx->cur.baq; hb.sq = ctx->cur.sq;
    DevConf dc;
    memset(&dc, 0, sizeof(dc));
    if (make_devconf(conf, &hb, dc)) return 1;
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (ctx->profiling) cudaEventRecord(ctx->ev[3], st);
    // the sites, decided on the device, in column order into pinned host memory; then the counters; then the event the
    // host waits for: nothing of this needs the host before lfb200_sites_*
    ctx->last_dc = dc;
    ctx->test_enqueued = false;
    if (ctx->cur.n_cols > 0 && ctx->ensure_sites(std::max<size_t>(65536, (size_t)ctx->cur.n_cols / 16))) return fail("out of pinned host memory");
    {
        const unsigned site_cap = (unsigned)std::min<size_t>(ctx->h_sites_cap, 0xffffffffu);
        KeyBuilder kb;
        kb.add(dc); kb.add(ctx->cur); kb.add(ctx->ws); kb.add(ctx->d_lut); kb.add(st); kb.add(ctx->ls.side_mask); kb.add(bonf_start_dev);
        kb.add(ctx->d_sites); kb.add(site_cap); kb.add(ctx->h_counters);
        bool bad = false;
        run_graphed(ctx->g_test, !ctx->profiling && ctx->cur.n_cols > 0, kb.v, st, [&]() {
            launch_test(ctx->ls, dc, ctx->cur, ctx->d_lut, ctx->ws, st, ctx->profiling ? ctx->ev[4] : nullptr, ctx->profiling ? ctx->ev[5] : nullptr,
                        bonf_start_dev);
            if (ctx->cur.n_cols > 0) {
                launch_emit_sites(ctx->ls, dc, ctx->ws, ctx->d_sites, site_cap, st);
                if (cudaMemcpyAsync(ctx->h_counters, ctx->ws.counters, sizeof(Counters), cudaMemcpyDeviceToHost, st) != cudaSuccess) bad = true;
            }
        });
        if (bad) { cudaGetLastError(); return fail("could not queue the copy of the counters"); }
    }
    CU(cudaEventRecord(ctx->ev_done, st));
    ctx->test_enqueued = true;
    CU(cudaGetLastError());
    return 0;
}

extern "C" int lfb200_test_device(lfb200_ctx *ctx, const lfb200_conf_t *conf, void *stream)
{
    return test_device_impl(ctx, conf, stream, nullptr);
}

extern "C" int lfb200_test_device_from(lfb200_ctx *ctx, const lfb200_conf_t *conf, void *stream, const long long *bonf_start_dev)
{
    if (!bonf_start_dev) return fail("bonf_start_dev is NULL");
    return test_device_impl(ctx, conf, stream, bonf_start_dev);
}

extern "C" int lfb200_bonf_start_device(void *stream, const long long *tested_counts_dev, int rank, long long bonf_subst,
                                        long long *start_dev)
{
    if (!tested_counts_dev || !start_dev || rank < 0) return fail("bad arguments");
    launch_bonf_start(tested_counts_dev, rank, bonf_subst, start_dev, (cudaStream_t)stream);
    CU(cudaGetLastError());
    return 0;
}

// accessors for shard_comm.cpp (the context layout is private to this file)
const unsigned long long *lfb_ctx_ntested_dev(lfb200_ctx *ctx) { return (ctx && ctx->have_batch) ? &ctx->ws.counters->n_tested : nullptr; }
int lfb_ctx_device(lfb200_ctx *ctx) { return ctx ? ctx->device : -1; }
void **lfb_ctx_comm_slot(lfb200_ctx *ctx) { return ctx ? &ctx->comm_state : nullptr; }
int lfb_fail(const char *msg) { return fail("%s", msg); }

extern "C" int lfb200_ntested_copy_device(lfb200_ctx *ctx, void *stream, long long *dst_dev)
{
    if (!ctx || !ctx->have_batch) return fail("no screened batch");
    CU(cudaMemcpyAsync(dst_dev, &ctx->ws.counters->n_tested, sizeof(long long), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return 0;
}

static long long final_bonf(const lfb200_conf_t *conf, long long start, long long n_tested)
{
    if (!conf->bonf_dynamic || n_tested == 0) return start;
    return (start == 1 ? 0 : start) + 3 * n_tested;     // lofreq_call.c:794-800
}

// the long double images of the p-values of one site whose status the device has decided (snpcaller.c:1047-1059, 1166-1196)
static inline void fill_pvalues(lfb200_site_t &s)
{
    for (int i = 0; i < 3; ++i) {
        const unsigned char st = s.status[i];
        s.pvalue[i] = st == LFB200_ST_VALUE ? expl((long double)s.lnp[i]) : st == LFB200_ST_LDBLMIN ? LDBL_MIN : LDBL_MAX;
    }
}

// a site with a comparison inside its guard band: the reference's own long double sequence
static void refinish_site(lfb200_site_t &s, double sig)
{
    Cand cd;
    cd.col = s.col;
    cd.bonf = s.bonf;
    for (int i = 0; i < 3; ++i) {
        cd.lnp[i] = s.lnp[i];
        cd.cnt[i] = s.alt_count[i];
        cd.raw[i] = s.alt_raw_count[i];
    }
    cd.ln_floor = s.ln_floor;
    cd.flags = 0;
    cd.pad = 0;
    const unsigned char fl = s.flags;
    finish_site(cd, sig, s);
    s.flags = fl;
    s.ln_floor = cd.ln_floor;
}

// Waits for the test enqueued by lfb200_test_device*, finishes what is left for the host and advances the counters.
// view != NULL: *view points at the context's own pinned buffer (valid until the next test on this context) and nothing
// is copied; otherwise the sites are copied to `sites`.
static int sites_sync(lfb200_ctx *ctx, lfb200_conf_t *conf, void *stream, lfb200_site_t *sites, long long max_sites,
                      const lfb200_site_t **view, lfb200_summary_t *summary)
{
    if (!ctx || !ctx->have_batch) return fail("no screened batch");
    if (!ctx->test_enqueued) return fail("lfb200_test_device has not been called for this batch");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    lfb200_summary_t sm;
    memset(&sm, 0, sizeof(sm));
    sm.n_cols = ctx->cur.n_cols;
    long long n_cand = 0;
    static const bool dbg = getenv("LFB200_DEBUG_TIMING") != nullptr;
    auto now = [] { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = dbg ? now() : 0.0;
    double t1 = 0, t2 = 0, t3 = 0;
    CU(cudaEventSynchronize(ctx->ev_done));
    if (lfb_comm_error(ctx)) return fail("count exchange: a shard did not post its counts within the timeout (this batch started from a poisoned factor)");
    if (dbg) t1 = now();
    const double sig = (double)conf->sig;
    if (ctx->cur.n_cols > 0) {
        const Counters *c = ctx->h_counters;
        if (c->emit_overflow) {
            // more sites than the pinned buffer held: grow it and emit again (the candidates are still on the device)
            if (ctx->ensure_sites((size_t)c->n_cand)) return fail("out of pinned host memory for %u sites", c->n_cand);
            CU(cudaMemsetAsync(&ctx->ws.counters->n_fix, 0, sizeof(Counters) - offsetof(Counters, n_fix), st));
            launch_emit_sites(ctx->ls, ctx->last_dc, ctx->ws, ctx->d_sites, (unsigned)std::min<size_t>(ctx->h_sites_cap, 0xffffffffu), st);
            CU(cudaMemcpyAsync(ctx->h_counters, ctx->ws.counters, sizeof(Counters), cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            if (c->emit_overflow) return fail("internal error: site buffer");
        }
        if (c->err_flags & CF_RANGE) return fail("a column's tail lies outside the representable range");
        sm.n_tested = (long long)c->n_tested;
        sm.n_unsupported = c->n_unsupported;
        n_cand = c->n_cand;
        {
            // side streams for the next batch on this context only where this batch had work for them
            unsigned long long dp1 = 0, dp2 = 0;
            for (int bin = 0; bin < DP_NBIN1; ++bin) {
                dp1 += c->n_pjobs[4 * DP_NBIN1 + bin] + c->n_pjobs[5 * DP_NBIN1 + bin];
                dp2 += c->n_pjobs[6 * DP_NBIN1 + bin];
            }
            ctx->ls.side_mask = (c->n_jobs[0] >= 256 ? 1u : 0u) | (dp1 ? 2u : 0u) | (c->n_jobs[CLS_XL] ? 4u : 0u) | (dp2 ? 8u : 0u);
        }
        sm.n_heavy = c->n_jobs[0] + c->n_jobs[CLS_XL];        // k_mid, k_xl
        for (int i = 0; i < DP_NL; ++i)                        // k_dp (binned lists are capped, the excess is in the unbinned ones)
            sm.n_heavy += (i % DP_NBIN1) == DP_NBIN ? (long long)c->n_pjobs[i] : std::min<long long>(c->n_pjobs[i], ctx->ws.pcap);
        lfb200_site_t *hs = ctx->h_sites;
        // the few sites with a comparison inside its guard band
        if (c->n_fix) {
            if (c->n_fix <= (unsigned)EMIT_FIX_MAX) {
                for (unsigned k = 0; k < c->n_fix; ++k) {
                    if (c->fix[k] >= (unsigned)n_cand) return fail("internal error: site index");
                    refinish_site(hs[c->fix[k]], sig);
                }
            } else {
                for (long long i = 0; i < n_cand; ++i)
                    if (hs[i].flags & SITE_NEEDS_HOST) refinish_site(hs[i], sig);
            }
        }
        if (dbg) t2 = now();
        // long double images of the p-values, for callers that want them (lfb200_set_site_pvalues)
        if (!ctx->site_pvalues && n_cand) {
            // the device does not write the p-value bytes: clear what an earlier batch may have left there
            for (long long i = 0; i < n_cand; ++i) memset((void *)hs[i].pvalue, 0, sizeof(hs[i].pvalue));
        }
        if (ctx->site_pvalues && n_cand) {
            if (n_cand < 1024) {
                for (long long i = 0; i < n_cand; ++i) fill_pvalues(hs[i]);
            } else {
                if (!ctx->pool) {
                    // share the host cores with the other shards of this node (torchrun exports LOCAL_WORLD_SIZE)
                    unsigned hw = std::thread::hardware_concurrency();
                    const char *lws = getenv("LOCAL_WORLD_SIZE");
                    const unsigned procs = lws ? (unsigned)std::max(1, atoi(lws)) : 1u;
                    const unsigned want = std::max(2u, std::min(hw / std::max(1u, 2 * procs), 8u));
                    ctx->pool.reset(new WorkerPool(want - 1));
                }
                auto part_fn = [&](unsigned part, unsigned parts) {
                    const long long per = (n_cand + parts - 1) / parts;
                    const long long lo = (long long)part * per, hi = std::min<long long>(n_cand, lo + per);
                    for (long long i = lo; i < hi; ++i) fill_pvalues(hs[i]);
                };
                ctx->pool->run(part_fn);
            }
            errno = 0;
            feclearexcept(FE_ALL_EXCEPT);
        }
        if (view) {
            *view = hs;
        } else if (sites) {
            if (n_cand > max_sites) return fail("%lld sites but room for %lld", n_cand, max_sites);
            if (n_cand) memcpy(sites, hs, (size_t)n_cand * sizeof(lfb200_site_t));
        }
    } else if (view) {
        *view = ctx->h_sites;
    }
    if (dbg) t3 = now();
    if (dbg)
        fprintf(stderr, "[lfb200] sites: wait %.0f us, fix-ups %.0f us (%u), p-values + copy %.0f us (%lld sites)\n",
                t1 - t0, t2 - t1, ctx->cur.n_cols > 0 ? ctx->h_counters->n_fix : 0u, t3 - t2, n_cand);
    sm.n_sites = n_cand;
    sm.bonf_subst_final = final_bonf(conf, ctx->cur.n_cols > 0 ? ctx->h_counters->bonf_start_used : conf->bonf_subst, sm.n_tested);
    conf->bonf_subst = sm.bonf_subst_final;
    conf->num_snv_tests += 3 * sm.n_tested;       // lofreq_call.c:801
    sm.num_snv_tests = conf->num_snv_tests;
    if (summary) *summary = sm;
    return 0;
}

extern "C" int lfb200_sites_device(lfb200_ctx *ctx, lfb200_conf_t *conf, void *stream, lfb200_site_t *sites,
                                   long long max_sites, lfb200_summary_t *summary)
{
    return sites_sync(ctx, conf, stream, sites, max_sites, nullptr, summary);
}

extern "C" int lfb200_sites_view(lfb200_ctx *ctx, lfb200_conf_t *conf, void *stream, const lfb200_site_t **sites,
                                 lfb200_summary_t *summary)
{
    if (!sites) return fail("null argument");
    return sites_sync(ctx, conf, stream, nullptr, 0, sites, summary);
}

extern "C" int lfb200_sites_buffer(lfb200_ctx *ctx, const lfb200_site_t **sites)
{
    if (!ctx || !sites) return fail("null argument");
    *sites = ctx->h_sites;
    return 0;
}

extern "C" int lfb200_set_site_pvalues(lfb200_ctx *ctx, int on)
{
    if (!ctx) return fail("no context");
    ctx->site_pvalues = on != 0;
    return 0;
}

extern "C" void lfb200_site_fill_pvalues(lfb200_site_t *sites, long long n)
{
    for (long long i = 0; i < n; ++i) fill_pvalues(sites[i]);
    errno = 0;
    feclearexcept(FE_ALL_EXCEPT);
}

// The same work on a per-context finisher thread, so that the caller's thread can go on launching the next batch
// (on another context) while the long double images of this batch's p-values are computed.
void lfb200_ctx::finisher_loop()
{
    for (;;) {
        {
            std::unique_lock<std::mutex> lk(fin_m);
            fin_cv.wait(lk, [this] { return fin_state == 1 || fin_stop; });
            if (fin_stop) return;
        }
        const int rc = sites_sync(this, fin_conf, fin_stream, fin_sites, fin_max, nullptr, &fin_summary);
        {
            std::lock_guard<std::mutex> lk(fin_m);
            fin_rc = rc;
            if (rc) snprintf(fin_err, sizeof(fin_err), "%s", g_err);
            fin_state = 2;
        }
        fin_done.notify_all();
    }
}

extern "C" int lfb200_sites_begin(lfb200_ctx *ctx, lfb200_conf_t *conf, void *stream, lfb200_site_t *sites, long long max_sites)
{
    if (!ctx || !ctx->have_batch) return fail("no screened batch");
    std::unique_lock<std::mutex> lk(ctx->fin_m);
    if (ctx->fin_state != 0) return fail("a sites request is already pending on this context");
    ctx->fin_conf = conf;
    ctx->fin_stream = stream;
    ctx->fin_sites = sites;
    ctx->fin_max = max_sites;
    if (!ctx->site_pvalues) {
        // nothing is left for the host but a wait and a copy: done by lfb200_sites_end itself, no thread involved
        ctx->fin_state = 3;
        return 0;
    }
    if (!ctx->fin_thread.joinable()) ctx->fin_thread = std::thread([ctx] { ctx->finisher_loop(); });
    ctx->fin_state = 1;
    lk.unlock();
    ctx->fin_cv.notify_all();
    return 0;
}

extern "C" int lfb200_sites_end(lfb200_ctx *ctx, lfb200_summary_t *summary)
{
    if (!ctx) return fail("no context");
    std::unique_lock<std::mutex> lk(ctx->fin_m);
    if (ctx->fin_state == 0) return fail("no sites request pending on this context");
    if (ctx->fin_state == 3) {
        ctx->fin_state = 0;
        lk.unlock();
        return sites_sync(ctx, ctx->fin_conf, ctx->fin_stream, ctx->fin_sites, ctx->fin_max, nullptr, summary);
    }
    ctx->fin_done.wait(lk, [ctx] { return ctx->fin_state == 2; });
    ctx->fin_state = 0;
    if (summary) *summary = ctx->fin_summary;
    if (ctx->fin_rc) return fail("%s", ctx->fin_err);
    return 0;
}

extern "C" int lfb200_set_profiling(lfb200_ctx *ctx, int on)
{
    if (!ctx) return fail("no context");
    if (on && !ctx->ev[0])
        for (int i = 0; i < 6; ++i) CU(cudaEventCreate(&ctx->ev[i]));
    ctx->profiling = on != 0;
    return 0;
}

extern "C" int lfb200_get_profile(lfb200_ctx *ctx, float *ms4)
{
    if (!ctx || !ctx->profiling) return fail("profiling is off");
    CU(cudaEventSynchronize(ctx->ev[5]));
    CU(cudaEventElapsedTime(&ms4[0], ctx->ev[0], ctx->ev[1]));   // k_front + k_scan_tiles
    CU(cudaEventElapsedTime(&ms4[1], ctx->ev[1], ctx->ev[2]));   // (nothing any more)
    CU(cudaEventElapsedTime(&ms4[2], ctx->ev[3], ctx->ev[4]));   // k_prune2 (alone, first, when profiling)
    CU(cudaEventElapsedTime(&ms4[3], ctx->ev[4], ctx->ev[5]));   // k_dp<*>, k_mid, k_xl, fallbacks
    return 0;
}

extern "C" int lfb200_copy_counts_device(lfb200_ctx *ctx, void *stream, int *dst_dev)
{
    if (!ctx || !ctx->have_batch) return fail("no screened batch");
    if (ctx->cur.n_cols)
        CU(cudaMemcpyAsync(dst_dev, ctx->ws.cnt6, (size_t)ctx->cur.n_cols * 6 * sizeof(int), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return 0;
}

// ------------------------------------------------------------------------------------------------
// BAQ HMM: kpa_ext_glocal (kprobaln_ext.h:38-40) for a batch of reads
// ------------------------------------------------------------------------------------------------
extern "C" int lfb200_kpa_glocal_batch(lfb200_ctx *ctx, long long n, const unsigned char *ref, const long long *ref_off,
                                       const unsigned char *query, const long long *qry_off, const unsigned char *qual, float d, float e,
                                       int bw, int *state, unsigned char *q)
{
    if (!ctx) return fail("no context");
    if (n < 0 || !ref_off || !qry_off || (n > 0 && (!ref || !query || !state || !q))) return fail("lfb200_kpa_glocal_batch: bad arguments");
    if (n == 0) return 0;
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const size_t tot_r = (size_t)ref_off[n], tot_q = (size_t)qry_off[n];
    // band and length of every read (kprobaln_ext.c:100-103): the scratch of a launch is sized for its widest / longest read
    int lmax = 0, bwmax = 0;
    for (long long r = 0; r < n; ++r) {
        const long long lr = ref_off[r + 1] - ref_off[r], lq = qry_off[r + 1] - qry_off[r];
        if (lr < 0 || lq < 0 || lr > 100000 || lq > 100000) return fail("lfb200_kpa_glocal_batch: read %lld has length %lld / %lld", r, lr, lq);
        if (lr == 0 || lq == 0) continue;                      // the reference returns at once (kprobaln_ext.c:88): outputs untouched
        int b = (int)std::max(lr, lq);
        if (b > bw) b = bw;
        if (b < (int)std::llabs(lr - lq)) b = (int)std::llabs(lr - lq);
        bwmax = std::max(bwmax, b);
        lmax = std::max(lmax, (int)lq);
    }
    const void *dv;
    if (upload(ctx->k_ref, ref, tot_r, 16, st, &dv)) return 1;
    const unsigned char *d_ref = (const unsigned char *)dv;
    if (upload(ctx->k_roff, ref_off, (size_t)(n + 1) * 8, 0, st, &dv)) return 1;
    const long long *d_roff = (const long long *)dv;
    if (upload(ctx->k_qry, query, tot_q, 16, st, &dv)) return 1;
    const unsigned char *d_qry = (const unsigned char *)dv;
    if (upload(ctx->k_qoff, qry_off, (size_t)(n + 1) * 8, 0, st, &dv)) return 1;
    const long long *d_qoff = (const long long *)dv;
    const unsigned char *d_qual = nullptr;
    if (qual) {
        if (upload(ctx->k_qual, qual, tot_q, 16, st, &dv)) return 1;
        d_qual = (const unsigned char *)dv;
    }
    float q2p[256];
    for (int i = 0; i < 256; ++i) q2p[i] = (float)pow(10, -i / 10.);      // g_qual2prob, kprobaln_ext.c:118-120
    if (upload(ctx->k_q2p, q2p, sizeof(q2p), 0, st, &dv)) return 1;
    const float *d_q2p = (const float *)dv;
    if (ctx->k_state.ensure(std::max<size_t>(tot_q, 1) * 4) || ctx->k_q.ensure(std::max<size_t>(tot_q, 1))) return fail("out of device memory");
    // outputs of empty reads stay what the caller passed in
    CU(cudaMemcpyAsync(ctx->k_state.p, state, tot_q * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(ctx->k_q.p, q, tot_q, cudaMemcpyHostToDevice, st));
    const int fix_cap = 65536;
    if (ctx->k_fix.ensure(sizeof(unsigned) * 4 + (size_t)fix_cap * sizeof(KpaFix))) return fail("out of device memory");
    unsigned *d_nfix = (unsigned *)ctx->k_fix.p;
    KpaFix *d_fix = (KpaFix *)((char *)ctx->k_fix.p + 16);
    CU(cudaMemsetAsync(d_nfix, 0, 16, st));
    const int w3 = (2 * bwmax + 1) * 3 + 6;
    const size_t per_read = ((size_t)(lmax + 1) * w3 + 2 * (size_t)w3 + (size_t)lmax + 2) * sizeof(double);
    size_t free_b = 0, total_b = 0;
    CU(cudaMemGetInfo(&free_b, &total_b));
    const size_t budget = std::min<size_t>(free_b / 2, (size_t)24 << 30);
    long long chunk = (long long)std::max<size_t>(budget / std::max<size_t>(per_read, 1), 128);
    chunk = std::min<long long>(chunk, n);
    chunk = std::min<long long>(chunk, 1 << 20);
    const size_t chunk32 = ((size_t)chunk + 31) / 32 * 32;             // the scratch is laid out per warp of 32 reads
    const int rows = lmax + 1;
    if (ctx->k_f.ensure(chunk32 * rows * w3 * 8) || ctx->k_b.ensure(chunk32 * 2 * w3 * 8) || ctx->k_s.ensure(chunk32 * (rows + 1) * 8))
        return fail("out of device memory for the scratch of %lld reads (%zu bytes each)", chunk, per_read);
    for (long long r0 = 0; r0 < n; r0 += chunk) {
        const int cnt = (int)std::min<long long>(chunk, n - r0);
        launch_kpa_glocal(r0, cnt, d_ref, d_roff, d_qry, d_qoff, d_qual, d, e, bw, d_q2p, (double *)ctx->k_f.p, (double *)ctx->k_b.p,
                          (double *)ctx->k_s.p, w3, rows, (int *)ctx->k_state.p, (unsigned char *)ctx->k_q.p, d_fix, fix_cap, d_nfix, st);
    }
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(state, ctx->k_state.p, tot_q * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(q, ctx->k_q.p, tot_q, cudaMemcpyDeviceToHost, st));
    unsigned n_fix = 0;
    CU(cudaMemcpyAsync(&n_fix, d_nfix, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (n_fix > (unsigned)fix_cap) return fail("lfb200_kpa_glocal_batch: %u bases inside the rounding guard band (at most %d expected)", n_fix, fix_cap);
    if (n_fix) {
        // -4.343 ln(x) + .499 within 1e-9 of an integer: the truncation is decided with this libm, as the reference does
        std::vector<KpaFix> fx(n_fix);
        CU(cudaMemcpy(fx.data(), d_fix, (size_t)n_fix * sizeof(KpaFix), cudaMemcpyDeviceToHost));
        for (const KpaFix &f : fx) {
            const int k = (int)(-4.343 * log(f.x) + .499);
            q[f.base] = (unsigned char)(k > 100 ? 99 : k);
        }
    }
    return 0;
}

extern "C" long long lfb200_graph_replays(lfb200_ctx *ctx) { return ctx ? ctx->g_front.replays + ctx->g_test.replays : 0; }

extern "C" int lfb200_last_job_counts(lfb200_ctx *ctx, long long out[4])
{
    if (!ctx || !out) return fail("null argument");
    const Counters &c = *ctx->h_counters;          // copied back by the last sites_sync
    out[0] = out[1] = out[2] = out[3] = 0;
    for (int i = 0; i < DP_NL; ++i)
        out[0] += (i % DP_NBIN1) == DP_NBIN ? (long long)c.n_pjobs[i] : std::min<long long>(c.n_pjobs[i], ctx->ws.pcap);
    out[1] = c.n_jobs[CLS_FALLBACK] + c.n_jobs[3] + c.n_jobs[4] + c.n_jobs[5] + c.n_jobs[6];     // k_heavy_all
    out[2] = c.n_jobs[CLS_XL];
    out[3] = c.n_jobs[0];
    return 0;
}

extern "C" double lfb200_dfma_peak(lfb200_ctx *ctx, void *stream)
{
    if (!ctx) return 0.0;
    cudaSetDevice(ctx->device);
    return measure_dfma_per_second(ctx->ls.sms, (cudaStream_t)stream);
}

extern "C" int lfb200_device_results(lfb200_ctx *ctx, const int **alt_counts, const int **alt_raw_counts,
                                     const unsigned char **tested, const long long **bonf_used)
{
    if (!ctx || !ctx->have_batch) return fail("no screened batch");
    // cnt6 is [n][6]: callers index alt_counts[6c+i], alt_raw_counts = alt_counts + 3 with the same stride
    if (alt_counts) *alt_counts = ctx->ws.cnt6;
    if (alt_raw_counts) *alt_raw_counts = ctx->ws.cnt6 + 3;
    if (tested) *tested = ctx->ws.tested;
    if (bonf_used) {
        // the factors live as ranks + a start value on the device; the dense array is written when somebody asks for it
        if (!ctx->test_enqueued) return fail("bonf_used is available after lfb200_test_device");
        if (!ctx->bonf_used_valid) {
            CU(cudaSetDevice(ctx->device));
            CU(cudaEventSynchronize(ctx->ev_done));
            launch_bonf_used(ctx->ls, ctx->last_dc, ctx->ws, ctx->cur.n_cols, ctx->stream);
            CU(cudaStreamSynchronize(ctx->stream));
            ctx->bonf_used_valid = true;
        }
        *bonf_used = ctx->ws.bonf_used;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// host buffers in, sites out
// ------------------------------------------------------------------------------------------------
static int upload(DevBuf &buf, const void *src, size_t bytes, size_t pad, cudaStream_t st, const void **dev)
{
    if (!src) { *dev = nullptr; return 0; }
    if (buf.ensure(bytes + pad)) return fail("out of device memory (%zu bytes)", bytes + pad);
    if (bytes) CU(cudaMemcpyAsync(buf.p, src, bytes, cudaMemcpyHostToDevice, st));
    *dev = buf.p;
    return 0;
}

extern "C" int lfb200_set_host_planes(lfb200_ctx *ctx, int mode)
{
    if (!ctx) return fail("no context");
    if (mode != 0 && mode != 1) return fail("host plane mode must be 0 (copy) or 1 (read pinned planes in place)");
    ctx->host_planes = mode;
    return 0;
}

extern "C" int lfb200_call_columns(lfb200_ctx *ctx, lfb200_conf_t *conf, const lfb200_batch_t *hb,
                                   const lfb200_dense_out_t *dense, lfb200_site_t *sites, long long max_sites,
                                   lfb200_summary_t *summary)
{
    if (!ctx) return fail("no context");
    CU(cudaSetDevice(ctx->device));
    const long long n = hb->n_cols;
    if (n < 0) return fail("negative n_cols");
    cudaStream_t st = ctx->stream;
    lfb200_batch_t db = *hb;
    const size_t plane_bytes = n ? (size_t)hb->col_off[n] : 0;
    const void *d;
    if (upload(ctx->in_off, hb->col_off, (size_t)(n + 1) * 8, 0, st, &d)) return 1;
    db.col_off = (const long long *)d;
    if (upload(ctx->in_cnt, hb->nt_cnt, (size_t)n * 16, 0, st, &d)) return 1;
    db.nt_cnt = (const int *)d;
    if (upload(ctx->in_ref, hb->ref_base, (size_t)n, 0, st, &d)) return 1;
    db.ref_base = (const char *)d;
    if (upload(ctx->in_cov, hb->coverage, (size_t)n * 4, 0, st, &d)) return 1;
    db.coverage = (const int *)d;
    if (upload(ctx->in_nb, hb->num_bases, (size_t)n * 4, 0, st, &d)) return 1;
    db.num_bases = (const int *)d;
    // Quality planes: copied to the device, or — host_planes mode 1 and the plane lies in pinned (page-locked) host
    // memory — read in place over PCIe.  The kernels touch only the reads that decide (non-reference reads and the
    // first few reads of a tested column in k_front / k_prune2, whole columns only for candidate sites), a
    // fraction of the bytes a bulk copy moves.  Pinned memory is mapped page-wise, so the aligned 16-byte loads
    // that straddle the end of a plane stay inside the mapping.
    auto place = [&](DevBuf &buf, const unsigned char *h, const unsigned char **out) -> int {
        *out = nullptr;
        if (!h) return 0;
        if (ctx->host_planes == 1) {
            cudaPointerAttributes at;
            if (cudaPointerGetAttributes(&at, h) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer &&
                ((uintptr_t)at.devicePointer & 15) == 0) {
                *out = (const unsigned char *)at.devicePointer;
                return 0;
            }
            cudaGetLastError();
        }
        const void *dp;
        if (upload(buf, h, plane_bytes, 32, st, &dp)) return 1;
        *out = (const unsigned char *)dp;
        return 0;
    };
    if (place(ctx->in_bq, hb->bq, &db.bq)) return 1;
    // planes the flags switch off are not needed on the device
    if (place(ctx->in_mq, (conf->flag & LFB200_USE_MQ) ? hb->mq : nullptr, &db.mq)) return 1;
    if (place(ctx->in_baq, (conf->flag & LFB200_USE_BAQ) ? hb->baq : nullptr, &db.baq)) return 1;
    if (place(ctx->in_sq, (conf->flag & LFB200_USE_SQ) ? hb->sq : nullptr, &db.sq)) return 1;

    if (lfb200_screen_device(ctx, conf, &db, st)) return 1;
    if (lfb200_test_device(ctx, conf, st)) return 1;
    lfb200_summary_t sm;
    const lfb200_conf_t conf_in = *conf;
    if (lfb200_sites_device(ctx, conf, st, sites, max_sites, &sm)) return 1;

    if (dense) {
        const size_t nn = (size_t)n;
        if (dense->alt_counts || dense->alt_raw_counts) {
            std::vector<int> c6(nn * 6);
            if (nn) CU(cudaMemcpyAsync(c6.data(), ctx->ws.cnt6, nn * 24, cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            for (size_t c = 0; c < nn; ++c)
                for (int i = 0; i < 3; ++i) {
                    if (dense->alt_counts) dense->alt_counts[3 * c + i] = c6[6 * c + i];
                    if (dense->alt_raw_counts) dense->alt_raw_counts[3 * c + i] = c6[6 * c + 3 + i];
                }
        }
        if (dense->tested && nn) CU(cudaMemcpyAsync(dense->tested, ctx->ws.tested, nn, cudaMemcpyDeviceToHost, st));
        if (dense->bonf_used && nn) {
            launch_bonf_used(ctx->ls, ctx->last_dc, ctx->ws, n, st);
            ctx->bonf_used_valid = true;
            CU(cudaMemcpyAsync(dense->bonf_used, ctx->ws.bonf_used, nn * 8, cudaMemcpyDeviceToHost, st));
        }
        CU(cudaStreamSynchronize(st));
        for (size_t c = 0; c < nn * 3; ++c) {
            if (dense->lnp) dense->lnp[c] = 0.0;
            if (dense->status) dense->status[c] = LFB200_ST_LDBLMAX;
            if (dense->pvalues) dense->pvalues[c] = LDBL_MAX;
            if (dense->called) dense->called[c] = 0;
            if (dense->qual) dense->qual[c] = -1;
        }
        for (long long k = 0; k < sm.n_sites; ++k) {
            const lfb200_site_t &s = sites[k];
            for (int i = 0; i < 3; ++i) {
                const size_t at = (size_t)s.col * 3 + i;
                if (dense->lnp) dense->lnp[at] = s.lnp[i];
                if (dense->status) dense->status[at] = s.status[i];
                if (dense->pvalues) dense->pvalues[at] = s.pvalue[i];
                if (dense->called) dense->called[at] = s.called[i];
                if (dense->qual) dense->qual[at] = s.qual[i];
            }
        }
    }
    (void)conf_in;
    if (summary) *summary = sm;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// column builder: the per-column callback surface (plp.h:159-163) in front of the batched door
// ------------------------------------------------------------------------------------------------
struct lfb200_builder {
    lfb200_ctx *ctx;
    lfb200_conf_t *conf;
    long long batch_cols;
    lfb200_site_fn on_site;
    void *user;
    std::vector<long long> col_off, tags;
    std::vector<int> nt_cnt, coverage, num_bases;
    std::vector<char> ref;
    std::vector<unsigned char> bq, mq, baq, sq;
    bool any_mq = false, any_baq = false, any_sq = false;
    std::vector<lfb200_site_t> sites;
    // reporting tail: strand counts per column (fw A,C,G,T then rv A,C,G,T), variant callback
    std::vector<int> strands;
    bool any_strands = false, defer_flush = false;
    lfb200_variant_fn on_variant = nullptr;
    void *variant_user = nullptr;
};

extern "C" int lfb200_builder_create(lfb200_builder **out, lfb200_ctx *ctx, lfb200_conf_t *conf, long long batch_cols,
                                     lfb200_site_fn on_site, void *user)
{
    *out = nullptr;
    if (!ctx || !conf) return fail("builder needs a context and a conf");
    if (batch_cols <= 0) return fail("batch_cols must be positive");
    lfb200_builder *b = new lfb200_builder();
    b->ctx = ctx;
    b->conf = conf;
    b->batch_cols = batch_cols;
    b->on_site = on_site;
    b->user = user;
    b->col_off.push_back(0);
    *out = b;
    return 0;
}

static inline unsigned char qbyte(int q) { return q < 0 ? 255 : (q > 254 ? (q == 255 ? 255 : 254) : (unsigned char)q); }

extern "C" int lfb200_builder_add_column(lfb200_builder *b, long long tag, char ref_base, int coverage_plp, int num_bases,
                                         const int *const base_quals[4], const int *const map_quals[4],
                                         const int *const baq_quals[4], const int *const source_quals[4], const int n[4])
{
    if (!b) return fail("no builder");
    size_t total = 0;
    for (int g = 0; g < 4; ++g) {
        if (n[g] < 0 || (n[g] > 0 && !base_quals[g])) return fail("column %lld: bad base-quality group %d", tag, g);
        total += (size_t)n[g];
    }
    const size_t at = b->bq.size();
    const size_t pitch = (total + 15) & ~(size_t)15;          // 16-byte aligned columns: full-width loads on the device
    b->bq.resize(at + pitch, 0);
    b->mq.resize(at + pitch, 255);
    b->baq.resize(at + pitch, 255);
    b->sq.resize(at + pitch, 255);
    size_t w = at;
    for (int g = 0; g < 4; ++g) {
        const int *mqv = map_quals ? map_quals[g] : nullptr;
        const int *baqv = baq_quals ? baq_quals[g] : nullptr;
        const int *sqv = source_quals ? source_quals[g] : nullptr;
        for (int j = 0; j < n[g]; ++j, ++w) {
            const int q = base_quals[g][j];
            b->bq[w] = q < 0 ? 0 : (q > 255 ? 255 : (unsigned char)q);
            // mq 255 is the SAM "unknown" which the reference maps to -1 itself (snpcaller.c:451-453)
            if (mqv) { b->mq[w] = qbyte(mqv[j]); b->any_mq = true; }
            if (baqv) { b->baq[w] = qbyte(baqv[j]); b->any_baq = true; }
            if (sqv) { b->sq[w] = qbyte(sqv[j]); b->any_sq = true; }
        }
        b->nt_cnt.push_back(n[g]);
    }
    b->col_off.push_back((long long)(at + pitch));
    b->ref.push_back(ref_base);
    b->coverage.push_back(coverage_plp);
    b->num_bases.push_back(num_bases);
    b->tags.push_back(tag);
    b->strands.resize(b->ref.size() * 8, 0);
    if (b->defer_flush) return 0;
    if ((long long)b->ref.size() >= b->batch_cols) return lfb200_builder_flush(b);
    return 0;
}

extern "C" int lfb200_builder_add_column_strands(lfb200_builder *b, long long tag, char ref_base, int coverage_plp, int num_bases,
                                                 const int *const base_quals[4], const int *const map_quals[4],
                                                 const int *const baq_quals[4], const int *const source_quals[4], const int n[4],
                                                 const long *fw_counts, const long *rv_counts)
{
    if (!b) return fail("no builder");
    b->defer_flush = true;
    const int rc = lfb200_builder_add_column(b, tag, ref_base, coverage_plp, num_bases, base_quals, map_quals, baq_quals, source_quals, n);
    b->defer_flush = false;
    if (rc) return rc;
    if (fw_counts && rv_counts) {
        int *s8 = b->strands.data() + (b->ref.size() - 1) * 8;
        for (int g = 0; g < 4; ++g) {
            s8[g] = (int)fw_counts[g];
            s8[4 + g] = (int)rv_counts[g];
        }
        b->any_strands = true;
    }
    if ((long long)b->ref.size() >= b->batch_cols) return lfb200_builder_flush(b);
    return 0;
}

extern "C" int lfb200_builder_on_variant(lfb200_builder *b, lfb200_variant_fn fn, void *user)
{
    if (!b) return fail("no builder");
    b->on_variant = fn;
    b->variant_user = user;
    return 0;
}

extern "C" long long lfb200_builder_pending(const lfb200_builder *b) { return b ? (long long)b->ref.size() : 0; }

extern "C" int lfb200_builder_flush(lfb200_builder *b)
{
    if (!b) return fail("no builder");
    const long long n = (long long)b->ref.size();
    if (n == 0) return 0;
    lfb200_batch_t hb;
    hb.n_cols = n;
    hb.col_off = b->col_off.data();
    hb.nt_cnt = b->nt_cnt.data();
    hb.ref_base = b->ref.data();
    hb.coverage = b->coverage.data();
    hb.num_bases = b->num_bases.data();
    // tail padding so that the device may read whole 16-byte chunks
    for (auto *v : {&b->bq, &b->mq, &b->baq, &b->sq}) v->resize(v->size() + 32, 0);
    hb.bq = b->bq.data();
    hb.mq = b->any_mq ? b->mq.data() : nullptr;
    hb.baq = b->any_baq ? b->baq.data() : nullptr;
    hb.sq = b->any_sq ? b->sq.data() : nullptr;
    b->sites.resize((size_t)n);
    lfb200_summary_t sm;
    int rc = lfb200_call_columns(b->ctx, b->conf, &hb, nullptr, b->sites.data(), n, &sm);
    if (rc == 0 && b->on_site)
        for (long long i = 0; i < sm.n_sites; ++i) {
            const lfb200_site_t &s = b->sites[(size_t)i];
            b->on_site(&s, b->tags[(size_t)s.col], b->ref[(size_t)s.col], b->coverage[(size_t)s.col], b->user);
        }
    if (rc == 0 && b->on_variant) {
        // the loop of call_snvs over the alt bases of every site (lofreq_call.c:818-871), with report_var()'s additions
        std::vector<lfb200_variant_t> vars;
        std::vector<lfb200_dp4_t> tabs;
        for (long long i = 0; i < sm.n_sites; ++i) {
            const lfb200_site_t &s = b->sites[(size_t)i];
            const size_t c = (size_t)s.col;
            const char ref = b->ref[c];
            const int ref_i = ref == 'A' ? 0 : ref == 'C' ? 1 : ref == 'G' ? 2 : 3;
            const int *s8 = b->strands.data() + c * 8;
            for (int a = 0; a < 3; ++a) {
                if (!s.called[a]) continue;
                const int alt_i = a + (a >= ref_i);                     // A,C,G,T minus the reference base (snpcaller.c:391-397)
                lfb200_variant_t v;
                v.tag = b->tags[c];
                v.bonf = s.bonf;
                v.lnp = s.lnp[a];
                v.af = s.alt_raw_count[a] / (float)b->coverage[c];      // lofreq_call.c:835
                v.qual = s.qual[a];
                v.dp = b->coverage[c];
                v.sb = 0;
                v.hqa = s.alt_count[a];                                 // lofreq_call.c:860
                v.dp4.ref_fw = s8[ref_i]; v.dp4.ref_rv = s8[4 + ref_i];
                v.dp4.alt_fw = s8[alt_i]; v.dp4.alt_rv = s8[4 + alt_i];
                v.ref_base = ref;
                v.alt_base = "ACGT"[alt_i];
                vars.push_back(v);
                tabs.push_back(v.dp4);
            }
        }
        std::vector<int> sb(vars.size());
        rc = lfb200_sb_qual_batch(b->ctx, (long long)vars.size(), tabs.data(), sb.data());
        if (rc == 0)
            for (size_t k = 0; k < vars.size(); ++k) {
                vars[k].sb = sb[k];
                b->on_variant(&vars[k], b->variant_user);
            }
    }
    b->strands.clear();
    b->any_strands = false;
    b->col_off.assign(1, 0);
    b->tags.clear();
    b->nt_cnt.clear();
    b->coverage.clear();
    b->num_bases.clear();
    b->ref.clear();
    b->bq.clear();
    b->mq.clear();
    b->baq.clear();
    b->sq.clear();
    b->any_mq = b->any_baq = b->any_sq = false;
    return rc;
}

extern "C" void lfb200_builder_destroy(lfb200_builder *b) { delete b; }

// ------------------------------------------------------------------------------------------------
// link-compatible single problems
// ------------------------------------------------------------------------------------------------
extern "C" int lfb200_snpcaller_batch(lfb200_ctx *ctx, long long n, const double *err_probs, const long long *ep_off,
                                      const int *noncons_counts, const long long *bonf, double sig_level,
                                      long double *snp_pvalues, double *lnp, unsigned char *status)
{
    if (!ctx) return fail("no context");
    if (n <= 0) return 0;
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const size_t n_ep = (size_t)ep_off[n];
    const void *d;
    ProbBatch pb;
    pb.n = n;
    pb.sig = sig_level;
    if (upload(ctx->p_ep, err_probs, n_ep * 8, 8, st, &d)) return 1;
    pb.err_probs = (const double *)d;
    if (upload(ctx->p_off, ep_off, (size_t)(n + 1) * 8, 0, st, &d)) return 1;
    pb.ep_off = (const long long *)d;
    if (upload(ctx->p_cnt, noncons_counts, (size_t)n * 12, 0, st, &d)) return 1;
    pb.counts = (const int *)d;
    if (upload(ctx->p_bonf, bonf, (size_t)n * 8, 0, st, &d)) return 1;
    pb.bonf = (const long long *)d;
    if (ctx->p_out.ensure((size_t)n * sizeof(Cand))) return fail("out of device memory");
    launch_prob_jobs(ctx->ls.sms, pb, (Cand *)ctx->p_out.p, st);
    CU(cudaGetLastError());
    if (ctx->ensure_cand((size_t)n)) return fail("out of pinned host memory");
    CU(cudaMemcpyAsync(ctx->h_cand, ctx->p_out.p, (size_t)n * sizeof(Cand), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    for (long long i = 0; i < n; ++i) {
        const Cand &cd = ctx->h_cand[(size_t)i];
        if (cd.flags & CF_UNSUPPORTED) return fail("problem %lld: alt count above %d or above the number of reads", i, 16384);
        if (cd.flags & CF_RANGE) return fail("problem %lld: tail outside the representable range", i);
        lfb200_site_t s;
        finish_site(cd, sig_level, s);
        for (int a = 0; a < 3; ++a) {
            snp_pvalues[3 * i + a] = s.pvalue[a];
            if (lnp) lnp[3 * i + a] = s.lnp[a];
            if (status) status[3 * i + a] = s.status[a];
        }
    }
    return 0;
}

static lfb200_ctx *g_default_ctx = nullptr;

extern "C" int lfb200_snpcaller(long double *snp_pvalues, const double *err_probs, const int num_err_probs,
                                const int *noncons_counts, const long long int bonf_factor, const double sig_level,
                                const int approx_threshold_n)
{
    if (approx_threshold_n > 0) {
        // a reference build without GSL exits here (snpcaller.c:1118-1125); we refuse without killing the caller
        fprintf(stderr, "FATAL(lofreq_b200): --approx-threshold needs the GSL Poisson pre-test, which this path does not provide\n");
        return 1;
    }
    if (!g_default_ctx && lfb200_create(&g_default_ctx, 0)) {
        fprintf(stderr, "FATAL(lofreq_b200): %s\n", g_err);
        return 1;
    }
    const long long off[2] = {0, num_err_probs};
    const long long bonf[1] = {bonf_factor};
    if (lfb200_snpcaller_batch(g_default_ctx, 1, err_probs, off, noncons_counts, bonf, sig_level, snp_pvalues, nullptr, nullptr)) {
        fprintf(stderr, "FATAL(lofreq_b200): %s\n", g_err);
        return 1;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// poissbin() rows and plp_to_errprobs() (snpcaller.h:72-75, 93-96)
// ------------------------------------------------------------------------------------------------
extern "C" int lfb200_poissbin_batch(lfb200_ctx *ctx, long long n, const double *err_probs, const long long *ep_off,
                                     const int *num_failures, const long long *bonf, double sig, const long long *row_off,
                                     double *rows, long double *pvalues, int *n_end)
{
    if (!ctx) return fail("no context");
    if (n <= 0) return 0;
    if (!err_probs || !ep_off || !num_failures || !bonf || !row_off || !rows) return fail("null argument");
    for (long long i = 0; i < n; ++i) {
        if (num_failures[i] < 1) return fail("problem %lld: num_failures must be at least 1", i);
        if (row_off[i + 1] - row_off[i] < num_failures[i] + 1) return fail("problem %lld: row_off leaves no room for the row", i);
    }
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const size_t n_ep = (size_t)ep_off[n], n_rows = (size_t)row_off[n];
    const void *d;
    ProbBatch pb;
    pb.n = n;
    pb.sig = sig;
    pb.counts = nullptr;
    if (upload(ctx->p_ep, err_probs, n_ep * 8, 8, st, &d)) return 1;
    pb.err_probs = (const double *)d;
    if (upload(ctx->p_off, ep_off, (size_t)(n + 1) * 8, 0, st, &d)) return 1;
    pb.ep_off = (const long long *)d;
    if (upload(ctx->p_bonf, bonf, (size_t)n * 8, 0, st, &d)) return 1;
    pb.bonf = (const long long *)d;
    // num_failures | row_off in one buffer; rows | double buffers | n_end in another
    std::vector<long long> meta((size_t)n + 1 + ((size_t)n + 1) / 2 + 1);
    memcpy(meta.data(), row_off, (size_t)(n + 1) * 8);
    memcpy(meta.data() + n + 1, num_failures, (size_t)n * 4);
    if (upload(ctx->p_cnt, meta.data(), meta.size() * 8, 0, st, &d)) return 1;
    const long long *d_row_off = (const long long *)d;
    const int *d_k = (const int *)(d_row_off + n + 1);
    if (ctx->p_out.ensure(n_rows * 8 * 3 + (size_t)n * 4 + 64)) return fail("out of device memory");
    double *d_rows = (double *)ctx->p_out.p, *d_buf = d_rows + n_rows;
    int *d_nend = (int *)(d_buf + 2 * n_rows);
    launch_poissbin_rows(pb, d_k, d_row_off, d_buf, d_rows, d_nend, st);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(rows, d_rows, n_rows * 8, cudaMemcpyDeviceToHost, st));
    std::vector<int> h_end((size_t)n);
    CU(cudaMemcpyAsync(h_end.data(), d_nend, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    for (long long i = 0; i < n; ++i) {
        if (pvalues) pvalues[i] = expl_clamped(rows[row_off[i] + num_failures[i]], false);      // snpcaller.c:1047-1059
        if (n_end) n_end[i] = h_end[(size_t)i];
    }
    return 0;
}


extern "C" double *lfb200_poissbin(long double *pvalue, const double *err_probs, const int num_err_probs, const int num_failures,
                                   const long long int bonf, const double sig)
{
    if (pvalue) *pvalue = LDBL_MAX;                    // snpcaller.c:1030
    if (!g_default_ctx && lfb200_create(&g_default_ctx, 0)) {
        fprintf(stderr, "FATAL(lofreq_b200): %s\n", g_err);
        return nullptr;
    }
    if (num_failures < 1 || num_err_probs < 0) {
        fprintf(stderr, "FATAL(lofreq_b200): poissbin needs num_failures >= 1\n");
        return nullptr;
    }
    double *row = (double *)malloc(((size_t)num_failures + 1) * sizeof(double));
    if (!row) {
        fprintf(stderr, "FATAL: couldn't allocate memory at %s:%s():%d\n", __FILE__, __FUNCTION__, __LINE__);
        return nullptr;
    }
    const long long off[2] = {0, num_err_probs}, roff[2] = {0, (long long)num_failures + 1}, b[1] = {bonf};
    long double pv = LDBL_MAX;
    if (lfb200_poissbin_batch(g_default_ctx, 1, err_probs, off, &num_failures, b, sig, roff, row, &pv, nullptr)) {
        fprintf(stderr, "FATAL(lofreq_b200): %s\n", g_err);
        free(row);
        return nullptr;
    }
    if (pvalue) *pvalue = pv;
    return row;
}

extern "C" int lfb200_batch_errprobs(lfb200_ctx *ctx, const lfb200_conf_t *conf, const lfb200_batch_t *hb, double *err_probs,
                                     int *num_err_probs, int *alt_bases, int *alt_counts, int *alt_raw_counts)
{
    if (!ctx) return fail("no context");
    if (!hb || !conf || !err_probs || !num_err_probs) return fail("null argument");
    const long long n = hb->n_cols;
    if (n <= 0) return 0;
    if (!hb->bq) return fail("the bq plane is required");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    DevConf dc;
    if (make_devconf(conf, hb, dc)) return 1;
    const size_t total = (size_t)hb->col_off[n];
    lfb200_batch_t db = *hb;
    const void *d;
    if (upload(ctx->in_off, hb->col_off, (size_t)(n + 1) * 8, 0, st, &d)) return 1;
    db.col_off = (const long long *)d;
    if (upload(ctx->in_cnt, hb->nt_cnt, (size_t)n * 16, 0, st, &d)) return 1;
    db.nt_cnt = (const int *)d;
    if (upload(ctx->in_ref, hb->ref_base, (size_t)n, 0, st, &d)) return 1;
    db.ref_base = (const char *)d;
    db.coverage = nullptr;
    db.num_bases = nullptr;
    if (upload(ctx->in_bq, hb->bq, total, 32, st, &d)) return 1;
    db.bq = (const unsigned char *)d;
    if (upload(ctx->in_mq, hb->mq, total, 32, st, &d)) return 1;
    db.mq = (const unsigned char *)d;
    if (upload(ctx->in_baq, hb->baq, total, 32, st, &d)) return 1;
    db.baq = (const unsigned char *)d;
    if (upload(ctx->in_sq, hb->sq, total, 32, st, &d)) return 1;
    db.sq = (const unsigned char *)d;
    DevBatch dbatch;
    to_devbatch(&db, dbatch);
    if (ctx->p_ep.ensure(total * 8 + 64)) return fail("out of device memory");
    if (ctx->p_out.ensure((size_t)n * 40 + 64)) return fail("out of device memory");
    double *d_ep = (double *)ctx->p_ep.p;
    int *d_n = (int *)ctx->p_out.p, *d_c9 = d_n + n;
    CU(cudaMemsetAsync(d_n, 0, (size_t)n * 40, st));
    launch_errprobs(dc, dbatch, ctx->d_lut, d_ep, d_n, d_c9, st);
    CU(cudaGetLastError());
    std::vector<int> h((size_t)n * 10);
    CU(cudaMemcpyAsync(err_probs, d_ep, total * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(h.data(), d_n, (size_t)n * 40, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    for (long long c = 0; c < n; ++c) {
        const char r = hb->ref_base[c];
        const bool acgt = r == 'A' || r == 'C' || r == 'G' || r == 'T';
        num_err_probs[c] = acgt ? h[(size_t)c] : 0;
        const int *o = h.data() + n + 9 * c;
        for (int k = 0; k < 3; ++k) {
            if (alt_bases) alt_bases[3 * c + k] = acgt ? o[k] : 0;
            if (alt_counts) alt_counts[3 * c + k] = acgt ? o[3 + k] : 0;
            if (alt_raw_counts) alt_raw_counts[3 * c + k] = acgt ? o[6 + k] : 0;
        }
    }
    return 0;
}

extern "C" void lfb200_plp_to_errprobs(double **err_probs, int *num_err_probs, int *alt_bases, int *alt_counts, int *alt_raw_counts,
                                       const lfb200_plp_col_t *p, const lfb200_conf_t *conf)
{
    *err_probs = nullptr;
    *num_err_probs = 0;
    if (!g_default_ctx && lfb200_create(&g_default_ctx, 0)) {
        fprintf(stderr, "FATAL(lofreq_b200): %s\n", g_err);
        return;
    }
    size_t total = 0;
    for (int g = 0; g < 4; ++g) total += (size_t)std::max(p->n[g], 0);
    const size_t room = std::max<size_t>(std::max<size_t>(total, (size_t)std::max(p->coverage_plp, 0)), 1);
    double *ep = (double *)malloc(room * sizeof(double));                // the reference: coverage_plp doubles (snpcaller.c:352)
    if (!ep) {
        fprintf(stderr, "FATAL: couldn't allocate memory at %s:%s():%d\n", __FILE__, __FUNCTION__, __LINE__);
        return;
    }
    std::vector<unsigned char> bq(total + 32, 0), mq(total + 32, 255), baq(total + 32, 255), sq(total + 32, 255);
    bool any_mq = false, any_baq = false, any_sq = false;
    size_t w = 0;
    int nt_cnt[4];
    for (int g = 0; g < 4; ++g) {
        nt_cnt[g] = std::max(p->n[g], 0);
        for (int j = 0; j < nt_cnt[g]; ++j, ++w) {
            const int q = p->base_quals[g][j];
            bq[w] = q < 0 ? 0 : (q > 255 ? 255 : (unsigned char)q);
            if (p->map_quals[g]) { mq[w] = qbyte(p->map_quals[g][j]); any_mq = true; }
            if (p->baq_quals[g]) { baq[w] = qbyte(p->baq_quals[g][j]); any_baq = true; }
            if (p->source_quals[g]) { sq[w] = qbyte(p->source_quals[g][j]); any_sq = true; }
        }
    }
    const long long col_off[2] = {0, (long long)total};
    lfb200_batch_t hb;
    memset(&hb, 0, sizeof(hb));
    hb.n_cols = 1;
    hb.col_off = col_off;
    hb.nt_cnt = nt_cnt;
    hb.ref_base = &p->ref_base;
    hb.bq = bq.data();
    hb.mq = any_mq ? mq.data() : nullptr;
    hb.baq = any_baq ? baq.data() : nullptr;
    hb.sq = any_sq ? sq.data() : nullptr;
    if (lfb200_batch_errprobs(g_default_ctx, conf, &hb, ep, num_err_probs, alt_bases, alt_counts, alt_raw_counts)) {
        fprintf(stderr, "FATAL(lofreq_b200): %s\n", g_err);
        free(ep);
        *num_err_probs = 0;
        return;
    }
    *err_probs = ep;
}

// ------------------------------------------------------------------------------------------------
// indel tests: plp_to_ins_errprobs / plp_to_del_errprobs (snpcaller.c:501-623) + snpcaller with (count, 0, 0)
// ------------------------------------------------------------------------------------------------
extern "C" int lfb200_indel_tests(lfb200_ctx *ctx, const lfb200_conf_t *conf, long long n, const long long *read_off,
                                  const unsigned char *iq, const unsigned char *mq, const unsigned char *aq, const unsigned char *sq,
                                  const int *event_count, const long long *bonf_indel, long double *pvalues, double *lnp,
                                  unsigned char *status, unsigned char *called, int *qual)
{
    if (!ctx) return fail("no context");
    if (n <= 0) return 0;
    if (!conf || !read_off || !iq || !event_count || !bonf_indel || !pvalues) return fail("null argument");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const size_t total = (size_t)read_off[n];
    // a test is a pseudo-column: "reference" group = the other reads, first alt group = the reads of the event
    std::vector<int> nt((size_t)n * 4, 0), cnt3((size_t)n * 3, 0);
    std::vector<char> ref((size_t)n, 'A');
    for (long long t = 0; t < n; ++t) {
        const long long reads = read_off[t + 1] - read_off[t];
        if (reads < 0 || event_count[t] < 0 || event_count[t] > reads) return fail("test %lld: bad read range / event count", t);
        nt[(size_t)4 * t] = (int)(reads - event_count[t]);
        nt[(size_t)4 * t + 1] = event_count[t];
        cnt3[(size_t)3 * t] = event_count[t];
    }
    DevConf dc;
    memset(&dc, 0, sizeof(dc));                        // no bq / jq filters, no overrides: every read counts (snpcaller.c:516-560)
    dc.use_mq = (conf->flag & LFB200_USE_MQ) && mq;
    dc.use_baq = (conf->flag & LFB200_USE_IDAQ) && aq;
    dc.use_sq = (conf->flag & LFB200_USE_SQ) && sq;
    dc.skip_jp = dc.skip_alt_jp = INFINITY;
    dc.sig = (double)conf->sig;
    dc.bonf_start = 1;
    DevBatch db;
    memset(&db, 0, sizeof(db));
    db.n_cols = n;
    const void *d;
    if (upload(ctx->in_off, read_off, (size_t)(n + 1) * 8, 0, st, &d)) return 1;
    db.col_off = (const long long *)d;
    if (upload(ctx->in_cnt, nt.data(), (size_t)n * 16, 0, st, &d)) return 1;
    db.nt_cnt = (const int *)d;
    if (upload(ctx->in_ref, ref.data(), (size_t)n, 0, st, &d)) return 1;
    db.ref_base = (const char *)d;
    if (upload(ctx->in_bq, iq, total, 32, st, &d)) return 1;
    db.bq = (const unsigned char *)d;
    if (upload(ctx->in_mq, dc.use_mq ? mq : nullptr, total, 32, st, &d)) return 1;
    db.mq = (const unsigned char *)d;
    if (upload(ctx->in_baq, dc.use_baq ? aq : nullptr, total, 32, st, &d)) return 1;
    db.baq = (const unsigned char *)d;
    if (upload(ctx->in_sq, dc.use_sq ? sq : nullptr, total, 32, st, &d)) return 1;
    db.sq = (const unsigned char *)d;
    if (ctx->p_ep.ensure(total * 8 + 64)) return fail("out of device memory");
    if (ctx->p_off.ensure((size_t)n * 40 + 64)) return fail("out of device memory");
    double *d_ep = (double *)ctx->p_ep.p;
    int *d_n = (int *)ctx->p_off.p, *d_c9 = d_n + n;
    launch_errprobs(dc, db, ctx->d_lut, d_ep, d_n, d_c9, st);      // nothing is filtered: column t fills [read_off[t], read_off[t+1])
    ProbBatch pb;
    pb.n = n;
    pb.sig = dc.sig;
    pb.err_probs = d_ep;
    pb.ep_off = db.col_off;
    if (upload(ctx->p_cnt, cnt3.data(), (size_t)n * 12, 0, st, &d)) return 1;
    pb.counts = (const int *)d;
    if (upload(ctx->p_bonf, bonf_indel, (size_t)n * 8, 0, st, &d)) return 1;
    pb.bonf = (const long long *)d;
    if (ctx->p_out.ensure((size_t)n * sizeof(Cand))) return fail("out of device memory");
    launch_prob_jobs(ctx->ls.sms, pb, (Cand *)ctx->p_out.p, st);
    CU(cudaGetLastError());
    if (ctx->ensure_cand((size_t)n)) return fail("out of pinned host memory");
    CU(cudaMemcpyAsync(ctx->h_cand, ctx->p_out.p, (size_t)n * sizeof(Cand), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    for (long long t = 0; t < n; ++t) {
        const Cand &cd = ctx->h_cand[(size_t)t];
        if (cd.flags & CF_UNSUPPORTED) return fail("test %lld: event count above %d", t, 16384);
        if (cd.flags & CF_RANGE) return fail("test %lld: tail outside the representable range", t);
        lfb200_site_t s;
        finish_site(cd, dc.sig, s);
        pvalues[t] = s.pvalue[0];
        if (lnp) lnp[t] = s.lnp[0];
        if (status) status[t] = s.status[0];
        if (called) called[t] = s.called[0];
        if (qual) qual[t] = s.qual[0];
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// binom() (binom.c:52-93): binomial CDF / survival function, batched on the device
// ------------------------------------------------------------------------------------------------
extern "C" int lfb200_binom_batch(lfb200_ctx *ctx, long long n, const int *num_trials, const int *num_success,
                                  const double *prob_success, double *p, double *q, int *status)
{
    if (!ctx) return fail("no context");
    if (n <= 0) return 0;
    if (!num_trials || !num_success || !prob_success || !status) return fail("null argument");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const void *d_n, *d_s, *d_pr;
    if (upload(ctx->p_cnt, num_trials, (size_t)n * 4, 0, st, &d_n)) return 1;
    if (upload(ctx->p_off, num_success, (size_t)n * 4, 0, st, &d_s)) return 1;
    if (upload(ctx->p_ep, prob_success, (size_t)n * 8, 0, st, &d_pr)) return 1;
    // outputs: cum | ccum | status in one buffer
    if (ctx->p_out.ensure((size_t)n * 20)) return fail("out of device memory");
    double *d_cum = (double *)ctx->p_out.p, *d_ccum = d_cum + n;
    int *d_status = (int *)(d_ccum + n);
    if (launch_binom(n, (const int *)d_n, (const int *)d_s, (const double *)d_pr, d_cum, d_ccum, d_status, st))
        return fail("could not upload the Stirling table");
    CU(cudaGetLastError());
    std::vector<double> h((size_t)n * 2);
    CU(cudaMemcpyAsync(h.data(), d_cum, (size_t)n * 16, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(status, d_status, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    for (long long i = 0; i < n; ++i) {
        if (status[i]) continue;         // the reference leaves *p / *q untouched when cdfbin refuses the arguments
        if (p) p[i] = h[(size_t)i];
        if (q) q[i] = h[(size_t)(n + i)];
    }
    return 0;
}

// same signature and return value as the reference's binom(): cdfbin's status, 0 = ok
extern "C" int lfb200_binom(double *p, double *q, int num_trials, int num_success, double prob_success)
{
    if (!g_default_ctx && lfb200_create(&g_default_ctx, 0)) {
        fprintf(stderr, "FATAL(lofreq_b200): %s\n", g_err);
        return 1;
    }
    int status = 1;                      // "error by default" (binom.c:56)
    if (lfb200_binom_batch(g_default_ctx, 1, &num_trials, &num_success, &prob_success, p, q, &status)) {
        fprintf(stderr, "FATAL(lofreq_b200): %s\n", g_err);
        return 1;
    }
    return status;
}

// ------------------------------------------------------------------------------------------------
// reporting tail: strand-bias quality on the device, INFO / record text on the host
// ------------------------------------------------------------------------------------------------
extern "C" int lfb200_sb_qual_batch(lfb200_ctx *ctx, long long n, const lfb200_dp4_t *dp4, int *sb_qual)
{
    if (!ctx) return fail("no context");
    if (n <= 0) return 0;
    if (!dp4 || !sb_qual) return fail("null argument");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const void *d_tab;
    if (upload(ctx->p_cnt, dp4, (size_t)n * sizeof(lfb200_dp4_t), 0, st, &d_tab)) return 1;
    if (ctx->p_out.ensure((size_t)n * 5 + 64)) return fail("out of device memory");
    int *d_sb = (int *)ctx->p_out.p;
    unsigned char *d_unsure = (unsigned char *)(d_sb + n);
    launch_sb_qual(ctx->ls.sms, (const int *)d_tab, n, d_sb, d_unsure, st);
    CU(cudaGetLastError());
    std::vector<unsigned char> unsure((size_t)n);
    CU(cudaMemcpyAsync(sb_qual, d_sb, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(unsure.data(), d_unsure, (size_t)n, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    for (long long i = 0; i < n; ++i)
        if (unsure[(size_t)i]) sb_qual[i] = sb_qual_host(&dp4[i].ref_fw);      // inside the guard band: glibc, like the reference
    return 0;
}

extern "C" int lfb200_format_snv_info(char *buf, unsigned long size, int dp, float af, int sb, const lfb200_dp4_t *dp4, int hqa)
{
    if (!buf || !dp4) return -1;
    const int k = snprintf(buf, size, "DP=%d;AF=%f;SB=%d;DP4=%d,%d,%d,%d;HQA=%d", dp, af, sb, dp4->ref_fw, dp4->ref_rv,
                           dp4->alt_fw, dp4->alt_rv, hqa);                      // vcf.c:615-624
    return (k < 0 || (unsigned long)k >= size) ? -1 : k;
}

extern "C" int lfb200_format_indel_info(char *buf, unsigned long size, int dp, float af, int sb, const lfb200_dp4_t *dp4, int hrun)
{
    if (!buf || !dp4) return -1;
    const int k = snprintf(buf, size, "DP=%d;AF=%f;SB=%d;DP4=%d,%d,%d,%d;INDEL;HRUN=%d", dp, af, sb, dp4->ref_fw, dp4->ref_rv,
                           dp4->alt_fw, dp4->alt_rv, hrun);                    // vcf.c:615-622
    return (k < 0 || (unsigned long)k >= size) ? -1 : k;
}

extern "C" int lfb200_format_snv_record(char *buf, unsigned long size, const char *chrom, long pos0, char ref, char alt, int qual,
                                        const char *info)
{
    if (!buf) return -1;
    int k;
    if (qual > -1)                                                              // vcf.c:472-495
        k = snprintf(buf, size, "%s\t%ld\t.\t%c\t%c\t%d\t.\t%s\n", chrom ? chrom : ".", pos0 + 1, ref, alt, qual, info ? info : ".");
    else
        k = snprintf(buf, size, "%s\t%ld\t.\t%c\t%c\t.\t.\t%s\n", chrom ? chrom : ".", pos0 + 1, ref, alt, info ? info : ".");
    return (k < 0 || (unsigned long)k >= size) ? -1 : k;
}

// ------------------------------------------------------------------------------------------------
// synthetic columns
// ------------------------------------------------------------------------------------------------
extern "C" int lfb200_synth_depths(int workload, long long c0, long long n_cols, int *depth_dev, void *stream)
{
    if (workload < 2 || workload > 5) return fail("workload must be 2..5");
    launch_synth_depths(workload, c0, n_cols, depth_dev, (cudaStream_t)stream);
    CU(cudaGetLastError());
    return 0;
}

extern "C" int lfb200_synth_columns(int workload, long long c0, long long n_cols, const long long *col_off_dev,
                                    int *nt_cnt_dev, char *ref_base_dev, unsigned char *bq_dev, unsigned char *mq_dev,
                                    unsigned char *baq_dev, void *stream)
{
    if (workload < 2 || workload > 5) return fail("workload must be 2..5");
    launch_synth_columns(workload, c0, n_cols, col_off_dev, nt_cnt_dev, ref_base_dev, bq_dev, mq_dev, baq_dev,
                         (cudaStream_t)stream);
    CU(cudaGetLastError());
    return 0;
}
