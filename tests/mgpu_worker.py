"""TEST INFRASTRUCTURE — one rank of tests/test_multi_gpu.py (launched by torch.distributed.run, one process per GPU).

Drives region shards exactly like bench.py drives one rank per GPU:
    lfb200_screen_device -> lfb200_comm_exchange (mailbox.cu: k_mail_exchange) -> lfb200_test_device_from -> sites
for N_BATCH batches (more than MAIL_DEPTH = 64, so that mailbox slots are reused and the acknowledgements are hit),
two contexts alternating so that exchanges of consecutive batches are in flight together.  Batch k of rank r holds
columns [(k * world + r) * n, +n) of workload C2; the single-process answer for batch k is one call over the
world * n columns of all ranks in rank order with a fresh conf (lofreq_call.c:794-801 continued across shards).
Rank 0 computes that answer on its own GPU afterwards and compares every site of every rank with it."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import lofreq_b200
    from lofreq_b200 import capi, shard, synth

    n = int(os.environ.get("MGPU_COLS", "20000"))
    n_batch = int(os.environ.get("MGPU_BATCHES", "72"))
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    NC = 2
    callers = [lofreq_b200.Caller(local) for _ in range(NC)]
    lib = callers[0].lib
    comms = [shard.ShardComm(callers[i], dev) for i in range(NC)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(NC)]
    sts = [C.c_void_p(s.cuda_stream) for s in streams]
    fields = ("col", "bonf", "alt_count", "alt_raw_count", "qual", "status", "called", "lnp")
    mine = []           # per batch: (sites, bonf_subst_final, n_tested, n_sites)
    gathered = []
    data = [synth.generate_device("C2", (k * world + rank) * n, n, device=str(dev)) for k in range(n_batch)]
    torch.cuda.synchronize()
    confs = [None] * NC
    prev_sites = [0] * NC

    def screen(k):
        i = k % NC
        cf = lofreq_b200.varcall_conf()
        capi.check(lib.lfb200_screen_device(callers[i]._ctx, C.byref(cf), C.byref(callers[i].device_batch(data[k])), sts[i]))
        comms[i].exchange(sts[i], sites_prev_batch=prev_sites[i])
        confs[i] = cf

    def test_and_sites(k):
        i = k % NC
        capi.check(lib.lfb200_test_device_from(callers[i]._ctx, C.byref(confs[i]), sts[i], comms[i].start_ptr))
        s, sm = callers[i].sites(confs[i], n, stream=sts[i])
        prev_sites[i] = int(sm.n_sites)
        mine.append((s.copy(), int(sm.bonf_subst_final), int(sm.n_tested), int(sm.n_sites)))
        gathered.append(comms[i].gathered(sts[i]))      # counts of every shard from this context's last exchange

    screen(0)
    for k in range(n_batch):
        if k + 1 < n_batch:
            screen(k + 1)          # the exchange of batch k+1 is posted before batch k has been tested
        test_and_sites(k)
    torch.cuda.synchronize()

    # every rank: its starting factor must be the exclusive prefix of the gathered tested counts
    for k, ((s, bonf_final, n_tested, n_sites), (tested_all, _)) in enumerate(zip(mine, gathered)):
        assert tested_all[rank] == n_tested, (k, tested_all, n_tested)
        assert bonf_final == 3 * sum(tested_all[: rank + 1]), (k, rank, bonf_final, tested_all)
        assert shard.final_counters(tested_all[: rank + 1])[0] == bonf_final

    # collect at rank 0 and compare with the single-process run
    payload = [[(s.tobytes(), bf, nt, ns) for (s, bf, nt, ns) in mine]]
    allres = [None] * world
    dist.gather_object(payload[0], allres if rank == 0 else None, dst=0)
    ok = True
    if rank == 0:
        dt = capi._site_dtype()
        one = lofreq_b200.Caller(local)
        n_cmp = 0
        for k in range(n_batch):
            whole = synth.generate_device("C2", k * world * n, world * n, device=str(dev))
            cf = lofreq_b200.varcall_conf()
            one.screen(one.device_batch(whole), cf)
            one.test(cf)
            want, want_sm = one.sites(cf, world * n)
            got = []
            for r in range(world):
                raw, bf, nt, ns = allres[r][k]
                s = np.frombuffer(raw, dtype=dt).copy()
                s["col"] += r * n
                got.append(s)
            got = np.concatenate(got)
            assert len(got) == len(want), (k, len(got), len(want))
            for f in fields:
                assert np.array_equal(got[f], want[f]), (k, f)
            assert allres[world - 1][k][1] == want_sm.bonf_subst_final, k
            assert 3 * sum(allres[r][k][2] for r in range(world)) == want_sm.num_snv_tests, k
            n_cmp += len(want)
        print("multi-GPU parity ok: %d ranks x %d batches x %d columns, %d sites compared (bonf of every site, site set, "
              "final counters == single-GPU run)" % (world, n_batch, n, n_cmp))
        one.close()
    dist.barrier()
    for c in callers:
        c.close()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
