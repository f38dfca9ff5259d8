// The per-batch count exchange between region shards without a collective: a mailbox in shared host memory
// (POSIX shm, registered with every process's CUDA context), written and polled by one warp.
//
// Why not ncclAllGather for this: the payload is 16 bytes per rank, but a NCCL kernel needs its own CTA with tens of
// KB of shared memory, so on a GPU filled by the persistent O(depth*K) kernels it waits for a slot, and it couples
// all ranks in lock step; measured at 2 GPUs it cost 0.1 ms of a 0.3 ms step.  A shard only needs the counts of the
// shards BEFORE it (the running Bonferroni factor continues from them, lofreq_call.c:794-800), so rank 0 never waits,
// rank r waits for r-1 ... 0, and the wait is a few system-scope loads by one lane per lower rank.
// NCCL stays for what BASELINE.json names it for: the final gather of the per-region counts (lfb200_comm_gathered).
#include <cuda_runtime.h>
#include "internal.h"

namespace lfb {

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ unsigned long long now_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

constexpr unsigned long long MAIL_TIMEOUT_NS = 20ull * 1000 * 1000 * 1000;

// One warp.  Lane 0 posts {tested columns of the batch just screened, sites of the batch before}; lanes r < rank wait
// for rank r's post of the same sequence number and add up its tested count; lane 0 turns the sum into this shard's
// starting factor and acknowledges, so that the lower ranks may reuse the slot MAIL_DEPTH exchanges later.
__global__ void k_mail_exchange(MailSlot *slots, unsigned long long *ack, int world, int rank, unsigned long long seq,
                                const unsigned long long *n_tested_dev, long long sites_prev, long long bonf_subst,
                                long long *d_mine, long long *d_start, int *d_err)
{
    const int lane = threadIdx.x;
    const unsigned long long t0 = now_ns();
    bool late = false;
    if (lane == 0) {
        const long long tested = (long long)*n_tested_dev;
        d_mine[0] = tested;
        d_mine[1] = sites_prev;
        if (seq > MAIL_DEPTH)                           // the slot about to be reused must have been read by every higher rank
            for (int j = rank + 1; j < world && !late; ++j)
                while (ld_acquire_sys(&ack[j]) + MAIL_DEPTH < seq)
                    if (now_ns() - t0 > MAIL_TIMEOUT_NS) { late = true; break; }
        MailSlot *s = &slots[(size_t)rank * MAIL_DEPTH + seq % MAIL_DEPTH];
        s->tested = tested;
        s->sites = sites_prev;
        st_release_sys(&s->seq, seq);
    }
    long long before = 0;
    for (int r = lane; r < rank; r += 32) {
        const MailSlot *s = &slots[(size_t)r * MAIL_DEPTH + seq % MAIL_DEPTH];
        while (ld_acquire_sys(&s->seq) != seq)
            if (now_ns() - t0 > MAIL_TIMEOUT_NS) { late = true; break; }
        before += *reinterpret_cast<const volatile long long *>(&s->tested);
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) before += __shfl_xor_sync(0xffffffffu, before, m);
    late = __any_sync(0xffffffffu, late);
    if (lane == 0) {
        // lofreq_call.c:794-800: after j tested columns the factor is 3j when it started at 1, else start + 3j
        *d_start = before > 0 ? (bonf_subst == 1 ? 0 : bonf_subst) + 3 * before : bonf_subst;
        if (late) {
            // a peer did not post in time: the partial sum would be a factor that is too small (false positives), so the
            // batch gets a factor no tail can pass, and the host turns the flag into an error at lfb200_sites_* / comm_gathered
            *d_start = 0x3fffffffffffffffll;
            *d_err = 1;
        }
        st_release_sys(&ack[rank], seq);
    }
}

void launch_mail_exchange(MailSlot *slots, unsigned long long *ack, int world, int rank, unsigned long long seq,
                          const unsigned long long *n_tested_dev, long long sites_prev, long long bonf_subst, long long *d_mine,
                          long long *d_start, int *d_err, cudaStream_t st)
{
    k_mail_exchange<<<1, 32, 0, st>>>(slots, ack, world, rank, seq, n_tested_dev, sites_prev, bonf_subst, d_mine, d_start, d_err);
}

}  // namespace lfb
