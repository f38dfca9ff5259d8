// Region shards on several GPUs: the one exchange of the path — every shard's tested-column count, so that each
// shard continues the running Bonferroni factor of the shards before it (lofreq_call.c:794-800 across shards; the
// reference's call-parallel instead restarts per region and sums the counts afterwards,
// lofreq2_call_pparallel.py:131-161) — as ONE NCCL all_gather per batch, issued directly on the kernels' stream:
// 2 x int64 per rank over NVLink (tested columns of this batch, sites of the previous batch = the per-region
// variant-count gather), no host round trip and no framework call in the path.  A one-thread kernel then turns the
// gathered counts into this shard's starting factor in device memory, where k_finalize reads it.
// NCCL is dlopen'ed (libnccl.so.2, the copy already loaded by the process if there is one) so that single-GPU hosts
// need no libnccl.
#include <cstdio>
#include <cstring>
#include <dlfcn.h>

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
};

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
    nccl_uid_t u;
    memcpy(u.internal, id, 128);
    const int rc = nccl().CommInitRank(&cs->comm, world, u, rank);
    if (rc) { delete cs; return nccl_fail("ncclCommInitRank", rc); }
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
    delete cs;
    *slot = nullptr;
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
    if (cudaMemcpyAsync(cs->h_all, cs->d_all, 16 * (size_t)cs->world, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess)
        return lfb_fail("copy of the gathered counts failed");
    for (int r = 0; r < cs->world; ++r) {
        if (tested_all) tested_all[r] = cs->h_all[2 * r];
        if (sites_prev_all) sites_prev_all[r] = cs->h_all[2 * r + 1];
    }
    return 0;
}
