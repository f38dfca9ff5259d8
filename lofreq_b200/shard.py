"""Region sharding across GPUs (one process per GPU, torch.distributed).

The reference's only parallel mode shards the genome into regions, runs `lofreq call` per region
and sums the per-region test counts afterwards (lofreq2_call_pparallel.py:131-161, 594-613); each
region restarts its running Bonferroni factor.  Here the shards exchange their tested-column counts
between the screen and the test phase, so every shard continues the running factor exactly where
the previous shard ends and the multi-GPU result equals the single-process `lofreq call`.

The only data-path communication is 2 x int64 per rank and batch (tested columns, site counts).  On
the GPUs it is the library's own exchange (ShardComm -> lfb200_comm_*: a shared-memory mailbox polled
by one warp, NCCL for the final gather); the torch.distributed helpers below (gather_counts) work
with the nccl backend (GPU tensors) and with gloo (CPU tensors, used by the CPU tests)."""


def shard_range(n_cols, rank, world):
    """contiguous columns [lo, hi) of this rank: regions in genomic order, sizes differ by at most one"""
    base, rem = divmod(n_cols, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_counts(value, device=None):
    """all_gather of one int64 per rank -> python list (world_size entries)"""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [int(value)]
    mine = torch.tensor([int(value)], dtype=torch.int64, device=device)
    out = torch.zeros(dist.get_world_size(), dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(out, mine)
    return [int(x) for x in out.tolist()]


def bonf_start_for_rank(tested_counts, rank, bonf_subst=1, bonf_dynamic=1):
    """conf->bonf_subst this rank must start from (lofreq_call.c:794-800): after j tested columns the
    factor is 3j when it started at 1, else start + 3j; a shard with nothing tested before it starts
    from the caller's value."""
    if not bonf_dynamic:
        return bonf_subst
    before = sum(tested_counts[:rank])
    if before == 0:
        return bonf_subst
    return (0 if bonf_subst == 1 else bonf_subst) + 3 * before


def final_counters(tested_counts, bonf_subst=1, bonf_dynamic=1, num_snv_tests=0):
    """(bonf_subst, num_snv_tests) after all shards: what a single `lofreq call` would report
    ("Number of substitution tests performed", lofreq_call.c:1562)"""
    total = sum(tested_counts)
    bonf = bonf_subst
    if bonf_dynamic and total:
        bonf = (0 if bonf_subst == 1 else bonf_subst) + 3 * total
    return bonf, num_snv_tests + 3 * total


class ShardComm:
    """The library's own exchange (include/lofreq_b200.h: lfb200_comm_*), one communicator per context.
    torch.distributed is used once, to hand rank 0's unique id to the other ranks."""

    def __init__(self, caller, device):
        import ctypes as C
        import torch
        import torch.distributed as dist
        from . import capi
        self.lib, self.caller = capi.load(), caller
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        uid = torch.zeros(128, dtype=torch.uint8)
        if self.rank == 0:
            buf = (C.c_ubyte * 128)()
            capi.check(self.lib.lfb200_comm_unique_id(buf))
            uid = torch.tensor(list(buf), dtype=torch.uint8)
        uid = uid.to(device)
        dist.broadcast(uid, 0)
        raw = bytes(uid.cpu().tolist())
        self._uid = (C.c_ubyte * 128).from_buffer_copy(raw)
        capi.check(self.lib.lfb200_comm_init(caller._ctx, self.world, self.rank, self._uid))
        self.start_ptr = C.c_void_p()

    def exchange(self, stream_ptr, sites_prev_batch=0, bonf_subst=1):
        """after screen: gathers {tested, sites_prev_batch} of every shard; returns the device pointer of this
        shard's starting Bonferroni factor (for lfb200_test_device_from).  Asynchronous."""
        import ctypes as C
        from . import capi
        capi.check(self.lib.lfb200_comm_exchange(self.caller._ctx, stream_ptr, int(bonf_subst), int(sites_prev_batch),
                                                  C.byref(self.start_ptr)))
        return self.start_ptr

    def gathered(self, stream_ptr):
        """(tested counts, sites of the previous batch) of every shard from the last exchange; synchronises"""
        import ctypes as C
        from . import capi
        a = (C.c_longlong * self.world)()
        b = (C.c_longlong * self.world)()
        capi.check(self.lib.lfb200_comm_gathered(self.caller._ctx, stream_ptr, a, b))
        return list(a), list(b)
