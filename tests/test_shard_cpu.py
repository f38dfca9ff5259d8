"""world_size-2 gloo test of the region-shard logic (lofreq_b200/shard.py): the shards exchange their
tested-column counts, continue the running Bonferroni factor, and the concatenated result equals the
single-process run.  The per-shard compute is stood in for by the CPU oracle (no GPU here); on the GPU
the same functions drive bench.py and tests/test_parity_gpu.py::test_two_shards_equal_one."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_cols, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from lofreq_b200 import shard
    from oracle import synth_np
    from oracle.pyoracle import Oracle, default_conf
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    lo, hi = shard.shard_range(n_cols, rank, world)
    b = synth_np.generate("C2", lo, hi - lo)
    orc = Oracle("port")
    # phase 1 ("screen"): how many columns of my shard are tested
    mine = orc.call_columns(b, default_conf())
    counts = shard.gather_counts(int(mine["tested"].sum()))
    # phase 2 ("test") with the running factor continued from the shards before me
    conf = default_conf(bonf_subst=shard.bonf_start_for_rank(counts, rank))
    out = orc.call_columns(b, conf)
    sites = shard.gather_counts(int(out["called"].any(axis=1).sum()))
    q.put((rank, lo, hi, out["bonf_used"], out["called"], out["qual"], counts, sites))
    dist.barrier()
    dist.destroy_process_group()


def test_two_region_shards_equal_single_process():
    from lofreq_b200 import shard
    from oracle import synth_np
    from oracle.pyoracle import Oracle, default_conf
    n_cols, world, port = 3001, 2, 29000 + os.getpid() % 2000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_cols, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = Oracle("port").call_columns(synth_np.generate("C2", 0, n_cols), default_conf())
    bonf = np.concatenate([r[3] for r in res])
    called = np.concatenate([r[4] for r in res])
    qual = np.concatenate([r[5] for r in res])
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == n_cols
    assert np.array_equal(bonf, want["bonf_used"])
    assert np.array_equal(called, want["called"]) and np.array_equal(qual, want["qual"])
    counts = res[0][6]
    assert counts == res[1][6] and sum(counts) == int(want["tested"].sum())
    assert shard.final_counters(counts) == (want["bonf_subst"], want["num_snv_tests"])
    assert sum(res[0][7]) == int(want["called"].any(axis=1).sum())


def test_shard_helpers():
    from lofreq_b200 import shard
    assert [shard.shard_range(10, r, 3) for r in range(3)] == [(0, 4), (4, 7), (7, 10)]
    assert shard.bonf_start_for_rank([5, 0, 7], 0) == 1
    assert shard.bonf_start_for_rank([5, 0, 7], 1) == 15
    assert shard.bonf_start_for_rank([5, 0, 7], 2) == 15
    assert shard.bonf_start_for_rank([0, 4], 1) == 1
    assert shard.bonf_start_for_rank([5, 2], 1, bonf_subst=100) == 115
    assert shard.bonf_start_for_rank([5, 2], 1, bonf_subst=100, bonf_dynamic=0) == 100
    assert shard.final_counters([5, 0, 7]) == (36, 36)
    assert shard.final_counters([0, 0]) == (1, 0)
    assert shard.gather_counts(3) == [3]
