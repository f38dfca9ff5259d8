// Region shards on several GPUs: the one exchange of the path — every shard's tested-column count, so that each
// shard continues the running Bonferroni factor of the shards before it (lofreq_call.c:794-800 across shards; the
// reference's call-parallel instead restarts per region and sums the counts afterwards,
// lofreq2_call_pparallel.py:131-161).  Per batch that is a mailbox in shared host memory polled by one warp on the
// kernels' stream (mailbox.cu: no collective, a shard waits only for the shards before it; LFB200_EXCHANGE_NCCL=1
// selects the earlier ncclAllGather per batch instead), which leaves this shard's starting factor in device memory
// where the test phase reads it (k_set_start) — no host round trip and no framework call in the path.  NCCL does the final gather of
// the per-region counts (lfb200_comm_gathered): 2 x int64 per rank over NVLink.
// NCCL is dlopen'ed (libnccl.so.2, the copy already loaded by the process if there is one) so that single-GPU hosts
// need no libnccl.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>

#include <cuda_runtime.h>

#include "../../include/lofreq_b200.h"
#include "internal.h"

using namespace lfb;

const unsigned long long *lfb_ctx_ntested_dev(lfb200_ctx *ctx);
int lfb_ctx_device(lfb200_ctx *ctx);
void **lfb_ctx_comm_slot(lfb200_ctx *ctx);
int lfb_fail(const char *msg);

namespace {

typedef struct { char internal[128]; } nccl_uid_t;          // ncclUniqueId, nccl.h:37-38
typedef void *nccl_comm_t;
enum { NCCL_INT64 = 4 };                                     // ncclDataType_t: ncclInt64

struct NcclApi {
    void *h = nullptr;
    int (*GetUniqueId)(nccl_uid_t *) = nullptr;
    int (*CommInitRank)(nccl_comm_t *, int, nccl_uid_t, int) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*CommDestroy)(nccl_comm_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool ok = false;
};

NcclApi &nccl()
{
    static NcclApi a;
    if (a.h || a.ok) return a;
    a.h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!a.h) a.h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!a.h) return a;
    a.GetUniqueId = (int (*)(nccl_uid_t *))dlsym(a.h, "ncclGetUniqueId");
    a.CommInitRank = (int (*)(nccl_comm_t *, int, nccl_uid_t, int))dlsym(a.h, "ncclCommInitRank");
    a.AllGather = (int (*)(const void *, void *, size_t, int, nccl_comm_t, cudaStream_t))dlsym(a.h, "ncclAllGather");
    a.CommDestroy = (int (*)(nccl_comm_t))dlsym(a.h, "ncclCommDestroy");
    a.GetErrorString = (const char *(*)(int))dlsym(a.h, "ncclGetErrorString");
    a.ok = a.GetUniqueId && a.CommInitRank && a.AllGather && a.CommDestroy && a.GetErrorString;
    return a;
}

struct CommState {
    nccl_comm_t comm = nullptr;
    int world = 1, rank = 0;
    long long *d_mine = nullptr;      // [2]: tested columns of this batch, sites of the previous batch
    long long *d_all = nullptr;       // [world][2]
    long long *d_tested = nullptr;    // [world]
    long long *d_start = nullptr;     // [1]
    long long *h_all = nullptr;       // pinned copy of d_all
    // mailbox in shared host memory (mailbox.cu): slots [world][MAIL_DEPTH], then ack [world]
    void *shm = nullptr;
    size_t shm_bytes = 0;
    MailSlot *d_slots = nullptr;      // device view of the mapping
    unsigned long long *d_ack = nullptr;
    unsigned long long seq = 0;       // exchanges issued so far on this communicator
    int *d_err = nullptr, *h_err = nullptr;   // mapped pinned flag: a peer did not post within the timeout
};

// every rank maps the same POSIX shm segment, named after the communicator's unique id, and registers it with CUDA
int open_mailbox(CommState *cs, const unsigned char id[128], char name[64])
{
    unsigned long long h1 = 1469598103934665603ull, h2 = 1099511628211ull;
    for (int i = 0; i < 128; ++i) {
        h1 = (h1 ^ id[i]) * 1099511628211ull;
        h2 = (h2 + id[i]) * 6364136223846793005ull + 1442695040888963407ull;
    }
    snprintf(name, 64, "/lfb200_%016llx%016llx", h1, h2);
    cs->shm_bytes = ((size_t)cs->world * MAIL_DEPTH * sizeof(MailSlot) + (size_t)cs->world * 8 + 4095) & ~(size_t)4095;
    const int fd = shm_open(name, O_CREAT | O_RDWR, 0600);
    if (fd < 0) return 1;
    if (ftruncate(fd, (off_t)cs->shm_bytes) != 0) { close(fd); return 1; }
    cs->shm = mmap(nullptr, cs->shm_bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (cs->shm == MAP_FAILED) { cs->shm = nullptr; return 1; }
    if (cudaHostRegister(cs->shm, cs->shm_bytes, cudaHostRegisterMapped | cudaHostRegisterPortable) != cudaSuccess) {
        cudaGetLastError();
        munmap(cs->shm, cs->shm_bytes);
        cs->shm = nullptr;
        return 1;
    }
    void *dp = nullptr;
    if (cudaHostGetDevicePointer(&dp, cs->shm, 0) != cudaSuccess) return 1;
    cs->d_slots = (MailSlot *)dp;
    cs->d_ack = (unsigned long long *)((char *)dp + (size_t)cs->world * MAIL_DEPTH * sizeof(MailSlot));
    if (cudaHostAlloc(&cs->h_err, sizeof(int), cudaHostAllocMapped) != cudaSuccess) return 1;
    *cs->h_err = 0;
    if (cudaHostGetDevicePointer((void **)&cs->d_err, cs->h_err, 0) != cudaSuccess) return 1;
    return 0;
}

int nccl_fail(const char *what, int rc)
{
    char buf[256];
    snprintf(buf, sizeof(buf), "NCCL %s failed: %s", what, nccl().GetErrorString ? nccl().GetErrorString(rc) : "?");
    return lfb_fail(buf);
}

}  // namespace

extern "C" int lfb200_comm_unique_id(unsigned char id[128])
{
    if (!nccl().ok) return lfb_fail("libnccl.so.2 could not be loaded");
    nccl_uid_t u;
    const int rc = nccl().GetUniqueId(&u);
    if (rc) return nccl_fail("ncclGetUniqueId", rc);
    memcpy(id, u.internal, 128);
    return 0;
}

extern "C" int lfb200_comm_init(lfb200_ctx *ctx, int world, int rank, const unsigned char id[128])
{
    void **slot = lfb_ctx_comm_slot(ctx);
    if (!slot) return lfb_fail("no context");
    if (*slot) return lfb_fail("communicator already initialised on this context");
    if (world < 1 || rank < 0 || rank >= world) return lfb_fail("bad world/rank");
    if (!nccl().ok) return lfb_fail("libnccl.so.2 could not be loaded");
    if (cudaSetDevice(lfb_ctx_device(ctx)) != cudaSuccess) return lfb_fail("cudaSetDevice failed");
    CommState *cs = new CommState();
    cs->world = world;
    cs->rank = rank;
    // the mailbox first: ncclCommInitRank returns once every rank has entered it, i.e. has mapped the segment, so rank 0
    // can unlink the name right after (nothing is left behind in /dev/shm, whatever happens later)
    char shm_name[64];
    const bool want_mail = !getenv("LFB200_EXCHANGE_NCCL");
    const bool have_mail = want_mail && open_mailbox(cs, id, shm_name) == 0;
    nccl_uid_t u;
    memcpy(u.internal, id, 128);
    const int rc = nccl().CommInitRank(&cs->comm, world, u, rank);
    if (want_mail && rank == 0) shm_unlink(shm_name);
    if (rc) { delete cs; return nccl_fail("ncclCommInitRank", rc); }
    if (want_mail && !have_mail) { delete cs; return lfb_fail("could not map the count mailbox (POSIX shm + cudaHostRegister)"); }
    if (cudaMalloc(&cs->d_mine, 16) != cudaSuccess || cudaMalloc(&cs->d_all, 16 * (size_t)world) != cudaSuccess ||
        cudaMalloc(&cs->d_tested, 8 * (size_t)world) != cudaSuccess || cudaMalloc(&cs->d_start, 8) != cudaSuccess ||
        cudaMallocHost(&cs->h_all, 16 * (size_t)world) != cudaSuccess)
        return lfb_fail("out of memory for the count exchange");
    cudaMemset(cs->d_mine, 0, 16);
    cudaMemset(cs->d_all, 0, 16 * (size_t)world);
    *slot = cs;
    return 0;
}

void lfb_comm_release(lfb200_ctx *ctx)
{
    void **slot = lfb_ctx_comm_slot(ctx);
    if (!slot || !*slot) return;
    CommState *cs = (CommState *)*slot;
    if (cs->comm && nccl().ok) nccl().CommDestroy(cs->comm);
    cudaFree(cs->d_mine);
    cudaFree(cs->d_all);
    cudaFree(cs->d_tested);
    cudaFree(cs->d_start);
    cudaFreeHost(cs->h_all);
    if (cs->h_err) cudaFreeHost(cs->h_err);
    if (cs->shm) {
        cudaHostUnregister(cs->shm);
        munmap(cs->shm, cs->shm_bytes);
    }
    delete cs;
    *slot = nullptr;
}

// non-zero when an exchange of this context timed out (checked by lfb200_sites_* after the stream has been waited for)
int lfb_comm_error(lfb200_ctx *ctx)
{
    void **slot = lfb_ctx_comm_slot(ctx);
    if (!slot || !*slot) return 0;
    CommState *cs = (CommState *)*slot;
    return cs->h_err ? *(volatile int *)cs->h_err : 0;
}

extern "C" int lfb200_comm_exchange(lfb200_ctx *ctx, void *stream, long long bonf_subst, long long sites_prev_batch,
                                    const long long **bonf_start_dev)
{
    void **slot = lfb_ctx_comm_slot(ctx);
    if (!slot || !*slot) return lfb_fail("lfb200_comm_init has not been called on this context");
    CommState *cs = (CommState *)*slot;
    const unsigned long long *nt = lfb_ctx_ntested_dev(ctx);
    if (!nt) return lfb_fail("no screened batch");
    cudaStream_t st = (cudaStream_t)stream;
    if (cs->d_slots) {
        // mailbox: one warp posts this shard's counts and waits for the shards before it (mailbox.cu)
        if (*cs->h_err) return lfb_fail("count exchange: a shard did not post its counts within the timeout");
        launch_mail_exchange(cs->d_slots, cs->d_ack, cs->world, cs->rank, ++cs->seq, nt, sites_prev_batch, bonf_subst, cs->d_mine,
                             cs->d_start, cs->d_err, st);
        if (cudaGetLastError() != cudaSuccess) return lfb_fail("launch failed");
        if (bonf_start_dev) *bonf_start_dev = cs->d_start;
        return 0;
    }
    // mine = [tested of the batch just screened (device -> device), sites of the batch finished before (host value)]
    if (cudaMemcpyAsync(cs->d_mine, nt, 8, cudaMemcpyDeviceToDevice, st) != cudaSuccess) return lfb_fail("memcpy failed");
    launch_set_i64(cs->d_mine + 1, sites_prev_batch, st);
    const int rc = nccl().AllGather(cs->d_mine, cs->d_all, 2, NCCL_INT64, cs->comm, st);
    if (rc) return nccl_fail("ncclAllGather", rc);
    launch_bonf_start_strided(cs->d_all, 2, cs->rank, bonf_subst, cs->d_start, st);
    if (cudaGetLastError() != cudaSuccess) return lfb_fail("launch failed");
    if (bonf_start_dev) *bonf_start_dev = cs->d_start;
    return 0;
}

extern "C" int lfb200_comm_gathered(lfb200_ctx *ctx, void *stream, long long *tested_all, long long *sites_prev_all)
{
    void **slot = lfb_ctx_comm_slot(ctx);
    if (!slot || !*slot) return lfb_fail("lfb200_comm_init has not been called on this context");
    CommState *cs = (CommState *)*slot;
    cudaStream_t st = (cudaStream_t)stream;
    if (cs->d_slots) {
        // the final gather of the per-region counts: the one NCCL collective of the path
        const int rc = nccl().AllGather(cs->d_mine, cs->d_all, 2, NCCL_INT64, cs->comm, st);
        if (rc) return nccl_fail("ncclAllGather", rc);
    }
    if (cudaMemcpyAsync(cs->h_all, cs->d_all, 16 * (size_t)cs->world, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess)
        return lfb_fail("copy of the gathered counts failed");
    if (cs->h_err && *(volatile int *)cs->h_err) return lfb_fail("count exchange: a shard did not post its counts within the timeout");
    for (int r = 0; r < cs->world; ++r) {
        if (tested_all) tested_all[r] = cs->h_all[2 * r];
        if (sites_prev_all) sites_prev_all[r] = cs->h_all[2 * r + 1];
    }
    return 0;
}
