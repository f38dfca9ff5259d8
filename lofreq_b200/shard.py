"""Region sharding across GPUs (one process per GPU, torch.distributed).

The reference's only parallel mode shards the genome into regions, runs `lofreq call` per region
and sums the per-region test counts afterwards (lofreq2_call_pparallel.py:131-161, 594-613); each
region restarts its running Bonferroni factor.  Here the shards exchange their tested-column counts
between the screen and the test phase, so every shard continues the running factor exactly where
the previous shard ends and the multi-GPU result equals the single-process `lofreq call`.

The only data-path communication is 1 x int64 per rank (all_gather), twice: tested columns, then
site counts.  Works with the nccl backend (GPU tensors) and with gloo (CPU tensors, used by the
CPU tests)."""


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


class DeviceCountExchange:
    """The same exchange kept on the device (NCCL on the kernels' stream, no host synchronisation):
    every rank contributes its tested-column count, `start` ends up holding the running Bonferroni factor
    this rank's test phase must start from (1 when nothing was tested before it).  A second int64 rides along
    in the same all_gather: the site count of the batch this context finished before (the "final per-region
    variant-count gather"), so one collective per batch carries both."""

    def __init__(self, device, bonf_subst=1):
        import torch
        import torch.distributed as dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.mine = torch.zeros(2, dtype=torch.int64, device=device)        # [tested (this batch), sites (previous batch)]
        self.all = torch.zeros((self.world, 2), dtype=torch.int64, device=device)
        self.tested_all = torch.zeros(self.world, dtype=torch.int64, device=device)
        self.start = torch.full((1,), bonf_subst, dtype=torch.int64, device=device)
        self.bonf_subst = bonf_subst

    def exchange(self, stream_ptr=None):
        """call with the torch current stream = the stream the count was copied on (stream_ptr = its handle)"""
        import ctypes as C
        import torch.distributed as dist
        from . import capi
        dist.all_gather_into_tensor(self.all.view(-1), self.mine)
        self.tested_all.copy_(self.all[:, 0])
        capi.check(capi.load().lfb200_bonf_start_device(stream_ptr, C.c_void_p(self.tested_all.data_ptr()), self.rank,
                                                         self.bonf_subst, C.c_void_p(self.start.data_ptr())))
        return self.start

    def site_counts(self):
        """site counts of the previous batch of every rank, as gathered by the last exchange (synchronises)"""
        return [int(x) for x in self.all[:, 1].tolist()]
