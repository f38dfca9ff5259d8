#!/usr/bin/env python
"""bench.py — pileup-columns/sec of the per-column SNV test (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--cols M] [--workload C2]

A step = one pass of the hot path over one batch of synthetic pileup columns (workload C2 of
BASELINE.json: 1M columns, depth 500, uniform Q30, MQ60; 1 % variant columns).  Per rank the batch is
resident in HBM for `value`; `e2e` goes through the host entry point of the C ABI with pinned host
buffers (H2D of the planes + D2H of the sites inside the timed region).

`--impl reference` times the reference's own CPU implementation (oracle/_ref/libsnpref.so: the
unmodified snpcaller.c driven like call_snvs; falls back to the C restatement when it was not built)
on all host cores over disjoint column ranges, the way lofreq2_call_pparallel.py shards regions.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "pileup_columns_per_sec"
UNIT = "columns/s"
DEPTH = {"C2": 500, "C3": 2000, "C4": 300}


_OUT_FD = None


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries (NCCL prints its version line there) must not add to it:
    fd 1 is pointed at stderr for the whole run and the line goes out through a duplicate of the original."""
    global _OUT_FD
    if _OUT_FD is None:
        sys.stdout.flush()
        _OUT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _OUT_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_OUT_FD, data)


def workload_desc(wl, cols):
    return {"C2": "C2: %d synthetic pileup columns, depth 500, uniform Q30, MQ60, 1%% variant sites" % cols,
            "C3": "C3: %d synthetic columns, depth 2000, Q20-40" % cols,
            "C4": "C4: %d synthetic columns, depth 300, mixed Q" % cols,
            "C5": "C5: %d synthetic columns, depth 50-10000 log-uniform, Q20-40" % cols}[wl]


# ------------------------------------------------------------------------------------------------
# CPU side: the reference (or the port) over column ranges, one process per core
# ------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    kind, wl, c0, n = args
    from oracle import synth_np
    from oracle.pyoracle import Oracle, default_conf
    orc = Oracle(kind)
    conf = default_conf()
    dt, tested, chunk = 0.0, 0, 10_000          # generate in chunks (numpy generator memory), time only the test
    for s0 in range(0, n, chunk):
        b = synth_np.generate(wl, c0 + s0, min(chunk, n - s0))
        t0 = time.perf_counter()
        out = orc.call_columns(b, conf)
        dt += time.perf_counter() - t0
        conf["bonf_subst"], conf["num_snv_tests"] = out["bonf_subst"], out["num_snv_tests"]   # running Bonferroni
        tested += int(out["tested"].sum())
    return dt, tested


def cpu_kind():
    from oracle.pyoracle import have_reference
    return "reference" if have_reference() else "port"


def cpu_run(wl, cols_per_proc, procs, c0=0, pool=None):
    """returns (columns/s over all procs, seconds of the slowest worker, kind)"""
    kind = cpu_kind()
    jobs = [(kind, wl, c0 + i * cols_per_proc, cols_per_proc) for i in range(procs)]
    if procs == 1 or pool is None:
        res = [_cpu_worker(j) for j in jobs]
        slow = sum(r[0] for r in res)
    else:
        res = pool.map(_cpu_worker, jobs, chunksize=1)
        slow = max(r[0] for r in res)
    return cols_per_proc * procs / slow, slow, kind


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def mark(self):
        """lines from here on belong to the timed region (the ones before: warm-up, also under load, kept as a fallback)"""
        self.mark_at = len(self.lines)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        lines = self.lines[getattr(self, "mark_at", 0):]
        if len(lines) < 2:                      # a very short timed region: the warm-up lines were taken under the same load
            lines = self.lines[max(0, len(self.lines) - 8):]
        for ln in lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            with open(p) as f:
                return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


STREAM_KERNELS = ("k_front", "k_prune2")
HEAVY_KERNELS = ("k_mid", "k_dp", "k_xl", "k_heavy")               # k_dp<0..2>, k_xl, k_heavy_all, k_heavy_xl
KERNELS_PER_STEP = 13      # (N > 1: + the count exchange and k_set_start) k_front, k_scan_tiles, k_prune2, k_mid, k_xl, k_dp<0>, k_dp<1>, k_dp<2>, k_heavy_xl, k_heavy_all, k_scan_blocks, k_rank_cands, k_emit_sites


def ncu_traffic(names, wl="C2"):
    """dram__bytes_read.sum + dram__bytes_write.sum per step of the named kernels, from the committed `ncu --set full`
    capture of one step of this workload (profiles/r2_step_kernels_<workload>.json, written by tools/ncu_extract.py in
    the same gpurun as the round's bench: tools/gpu_profile_round.sh); None when the capture is not there."""
    p = os.path.join(ROOT, "profiles", "r2_step_kernels_%s.json" % wl)
    try:
        with open(p) as f:
            ks = json.load(f)
    except Exception:
        return None
    tot, hit = 0.0, False
    for k in ks:
        nm = k.get("kernel", "")
        if any(nm.startswith(n) or ("lfb::" + n) in nm or (" " + n) in nm for n in names):
            tot += k.get("dram__bytes_read.sum [bytes]", 0.0) + k.get("dram__bytes_write.sum [bytes]", 0.0)
            hit = True
    return tot if hit else None


def algorithmic_bytes(sum_depth, n_cols, planes=2):
    """planes the configuration merges (bq, mq[, baq]) over the column's reads +
    per-column metadata in (col_off 8, nt_cnt 16, ref_base 1) + per-column results out (alt counts and
    raw counts 24, tested 1, bonf 8)."""
    return sum_depth * planes + (25 + 33) * n_cols


# ------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    procs = max(1, cores)
    # bounded sample: ~1.5 s of CPU work per process and step
    cols_per_proc = args.ref_cols_per_proc
    import multiprocessing as mp
    pool = mp.get_context("spawn").Pool(procs) if procs > 1 else None
    for _ in range(args.warmup):
        cpu_run(args.workload, max(cols_per_proc // 8, 256), procs, pool=pool)
    vals, secs = [], []
    for k in range(args.steps):
        v, s, kind = cpu_run(args.workload, cols_per_proc, procs, c0=k * cols_per_proc * procs, pool=pool)
        vals.append(v); secs.append(s)
    if pool:
        pool.close()
    value = sum(vals) / len(vals)
    sample = "%d columns per step (%d per process x %d processes), %s" % (cols_per_proc * procs, cols_per_proc, procs,
                                                                          workload_desc(args.workload, args.cols))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * sum(secs) / len(secs), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_desc(args.workload, args.cols) + " per GPU (region shard = rank*cols)",
                       "cols_per_gpu": args.cols, "depth": DEPTH.get(args.workload), "cpu_path": kind_desc(kind)},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def kind_desc(kind):
    return ("unmodified reference snpcaller.c/utils.c compiled into oracle/_ref/libsnpref.so, driven like call_snvs"
            if kind == "reference" else "C restatement oracle/snv_oracle.c (reference not built here)")


def run_ours(args):
    import numpy as np
    import torch
    import lofreq_b200
    from lofreq_b200 import capi, shard, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    wl, n = args.workload, args.cols
    # two contexts = two workspaces: the host finishing (D2H of the sites + long double arithmetic) of
    # batch k overlaps the kernels of batch k+1, the way a caller streaming many batches would run it
    # contexts = batches in flight; a third one (screen and count exchange two batches ahead) measured slower at 1 and 2 GPUs
    NC = args.contexts if args.contexts else 2
    callers = [lofreq_b200.Caller(local) for _ in range(NC)]
    caller = callers[0]
    lib = caller.lib
    streams = [torch.cuda.Stream(device=dev) for _ in range(NC)]
    sts = [C.c_void_p(s_.cuda_stream) for s_ in streams]

    # this rank's region shard: columns [rank*n, (rank+1)*n)
    t = synth.generate_device(wl, rank * n, n, with_baq=args.baq, device=str(dev))
    torch.cuda.synchronize()
    db = caller.device_batch(t)
    max_sites = n
    sites_buf = (capi.Site * max_sites)()          # e2e leg (lfb200_call_columns copies the sites out)
    # `value` leg: the sites stay in each context's pinned buffer (lfb200_sites_buffer), where k_emit_sites wrote them in
    # column order with status / called / QUAL decided on the device; the x87 long double images of the p-values are
    # not requested (lfb200_set_site_pvalues 0) — a caller that writes VCF needs QUAL, not the long double
    for c_ in callers:
        capi.check(lib.lfb200_set_site_pvalues(c_._ctx, 0))
    sms = [capi.Summary() for _ in range(NC)]
    sm = sms[0]
    confs = [None] * NC
    exchanges = [shard.ShardComm(callers[i], dev) for i in range(NC)] if world > 1 else None
    prev_sites = [0] * NC

    def screen(i):
        """gates + alt counts of the next batch on context i, and (N > 1) the exchange of the tested-column counts:
        1 x int64 per rank over NCCL on this stream, no host round trip.  Issued one batch ahead so that the
        collective has a whole batch of kernels to hide behind."""
        cf = lofreq_b200.varcall_conf()
        ctx = callers[i]._ctx
        capi.check(lib.lfb200_screen_device(ctx, C.byref(cf), C.byref(db), sts[i]))
        if world > 1:
            exchanges[i].exchange(sts[i], sites_prev_batch=prev_sites[i])     # ncclAllGather on this stream
        confs[i] = cf

    def test(i):
        """running Bonferroni (continued from the shards before this one), early-exit prune, O(depth*K) kernels"""
        ctx = callers[i]._ctx
        if world > 1:
            capi.check(lib.lfb200_test_device_from(ctx, C.byref(confs[i]), sts[i], exchanges[i].start_ptr))
        else:
            capi.check(lib.lfb200_test_device(ctx, C.byref(confs[i]), sts[i]))

    def launch(i):
        screen(i)
        test(i)

    def finish_begin(i):
        """D2H of the sites of context i + host finishing, on the context's own finisher thread"""
        capi.check(lib.lfb200_sites_begin(callers[i]._ctx, C.byref(confs[i]), sts[i], None, 0))

    def finish_end(i):
        capi.check(lib.lfb200_sites_end(callers[i]._ctx, C.byref(sms[i])))
        prev_sites[i] = int(sms[i].n_sites)   # the per-region variant-count gather rides in this context's next exchange
        return sms[i]

    def finish(i):
        finish_begin(i)
        return finish_end(i)

    host_t = {"screen": 0.0, "test": 0.0, "finish_begin": 0.0, "finish_end": 0.0}
    if os.environ.get("LFB200_HOST_TIMING"):       # where the launching thread spends its time (diagnostics only)
        def timed(name, fn):
            def wrapped(i):
                t0 = time.perf_counter()
                r = fn(i)
                host_t[name] += time.perf_counter() - t0
                return r
            return wrapped
        screen, test, finish_begin, finish_end = (timed("screen", screen), timed("test", test),
                                                  timed("finish_begin", finish_begin), timed("finish_end", finish_end))

    def run_steps(k_steps):
        """batch k on context k % NC.  The screen (and, N > 1, the count exchange) of a batch is issued NC - 1 batches
        ahead of its test, and the sites of batch k are finished on a host thread while the launching thread goes on"""
        lead = NC - 1
        for j in range(min(lead, k_steps)):
            screen(j % NC)
        done = 0                                   # batches whose finishing has been waited for
        for k in range(k_steps):
            test(k % NC)
            finish_begin(k % NC)
            nxt = k + lead
            if nxt < k_steps:
                while done <= nxt - NC:            # context nxt % NC must be done with batch nxt - NC
                    finish_end(done % NC)
                    done += 1
                screen(nxt % NC)
        while done < k_steps:
            finish_end(done % NC)
            done += 1

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- value: resident inputs -------------------------------------------------------------
    # the clock sampler starts before the warm-up: with 8 busy GPUs nvidia-smi needs tens of milliseconds for its first
    # line, and the timed region of a short run is not much longer
    sampler = ClockSampler(local)
    sampler.start()
    run_steps(max(args.warmup, 3))

    def agree(flag):
        """the same answer on every rank (the batches of a step are exchanged between the ranks: all run the same count)"""
        if world == 1:
            return bool(flag)
        t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return bool(int(t.item()))

    def under_load_until(enough, limit_s):
        """untimed steps that keep the GPU under the same load until the sampler has what `enough` asks for"""
        t_end = time.perf_counter() + limit_s
        while agree(sampler.proc is not None and not enough() and time.perf_counter() < t_end):
            run_steps(20)

    under_load_until(lambda: len(sampler.lines) >= 1, 5.0)      # nvidia-smi needs a while for its first line
    barrier()
    sampler.mark()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(streams[0])
    for k_ in host_t:
        host_t[k_] = 0.0
    t0 = time.perf_counter()
    run_steps(args.steps)
    e1.record(streams[0])          # issued after the last batch's host finishing returned
    barrier()
    wall = time.perf_counter() - t0
    dev_ms = e0.elapsed_time(e1)
    # a timed region shorter than the sampling period: the same load goes on (untimed) until the sampler has two lines of it
    under_load_until(lambda: len(sampler.lines) - sampler.mark_at >= 2, 3.0)
    if os.environ.get("LFB200_HOST_TIMING"):
        print("host seconds in the launching thread (timed steps):", host_t, "wall of timed steps", wall, file=sys.stderr)
    clocks = sampler.stop()
    sm = sms[(args.steps - 1) % NC]
    n_sites, n_tested, n_heavy = sm.n_sites, sm.n_tested, sm.n_heavy
    final_sites = [int(n_sites)]
    if world > 1:
        # the last batch's site counts: one more (tiny) gather outside the timed region
        final_sites = shard.gather_counts(int(n_sites), device=dev)
        # every run checks the exchange it just timed: the factor this shard's last batch started from must be the
        # exclusive prefix of the tested counts of all shards (lofreq_call.c:794-800 continued across shards)
        li = (args.steps - 1) % NC
        tested_all, _ = exchanges[li].gathered(sts[li])
        assert tested_all[rank] == int(n_tested), ("tested count", tested_all, int(n_tested))
        assert int(sm.bonf_subst_final) == 3 * sum(tested_all[: rank + 1]), ("running Bonferroni across shards", rank,
                                                                            int(sm.bonf_subst_final), tested_all)
    # CUDA events on the launching stream bracket the K steps (the closing event is recorded after the last
    # batch's host finishing has returned, so that host work is inside the region too); max over ranks
    el = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(el, op=dist.ReduceOp.MAX)
    ms_step = float(el.item()) / args.steps
    value = world * n / (ms_step * 1e-3)

    # ---- kernel times for the roofline: a few serialized steps on one context (no overlap between batches),
    #      CUDA events on the launching stream around each kernel group (lfb200_set_profiling)
    capi.check(lib.lfb200_set_profiling(callers[0]._ctx, 1))
    nprof = 5
    prof = np.zeros((nprof, 4), np.float32)
    for k in range(nprof):
        launch(0)
        finish(0)
        row = (C.c_float * 4)()
        capi.check(lib.lfb200_get_profile(callers[0]._ctx, row))
        prof[k] = list(row)
    capi.check(lib.lfb200_set_profiling(callers[0]._ctx, 0))
    capi.check(lib.lfb200_set_site_pvalues(caller._ctx, 1))     # the e2e leg runs the host entry point with its defaults

    # ---- e2e: host buffers through lfb200_call_columns ----------------------------------------
    total = t["total_bytes"]
    n_planes = 3 if args.baq else 2
    planes = [("bq", t["bq"][: total + 16]), ("mq", t["mq"][: total + 16])] + ([("baq", t["baq"][: total + 16])] if args.baq else [])
    hb_t = {k: torch.empty(v.shape, dtype=v.dtype, pin_memory=True).copy_(v) for k, v in
            [("col_off", t["col_off"]), ("nt_cnt", t["nt_cnt"]), ("ref_base", t["ref_base"])] + planes}
    torch.cuda.synchronize()
    hb = capi.Batch(n, hb_t["col_off"].data_ptr(), hb_t["nt_cnt"].data_ptr(), hb_t["ref_base"].data_ptr(), None,
                    hb_t["bq"].data_ptr(), hb_t["mq"].data_ptr(), hb_t["baq"].data_ptr() if args.baq else None, None, None)
    h2d = 8 * (n + 1) + 16 * n + n + n_planes * total

    def step_e2e():
        cf = lofreq_b200.varcall_conf()
        capi.check(lib.lfb200_call_columns(caller._ctx, C.byref(cf), C.byref(hb), None, sites_buf, max_sites, C.byref(sm)))
        return sm
    e2e_steps = max(1, min(args.steps, 5))

    def time_e2e(mode):
        """mode 1 (the default of lfb200_call_columns): planes that lie in pinned host memory are read in place over PCIe
        by the kernels — only the reads that decide a column cross the bus — and only the per-column metadata is copied;
        mode 0: every plane is copied to the device first"""
        capi.check(lib.lfb200_set_host_planes(caller._ctx, mode))
        for _ in range(2):
            step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            step_e2e()
        barrier()
        el = torch.tensor([(time.perf_counter() - t0) / e2e_steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(el, op=dist.ReduceOp.MAX)
        return float(el.item()), int(sm.n_sites), int(sm.num_snv_tests)

    if args.no_e2e:          # profiling runs only (ncu launch lists): never a bench line
        e2e_copy_s = e2e_s = float("nan")
    else:
        e2e_copy_s, sites_copy, tests_copy = time_e2e(0)
        e2e_s, sites_map, tests_map = time_e2e(1)
        capi.check(lib.lfb200_set_host_planes(caller._ctx, 1))
        assert (sites_copy, tests_copy) == (sites_map, tests_map), "in-place and copied planes disagree"
    e2e_value = world * n / e2e_s                 # headline: the entry point with its defaults on pinned host buffers
    d2h = 128 + int(sm.n_sites) * 80
    h2d_meta = 8 * (n + 1) + 16 * n + n

    # ---- rooflines -------------------------------------------------------------------------------
    # HBM side (k_front + k_prune2 stream the quality planes): algorithmic bytes per column =
    # every quality byte the configuration merges + metadata in + results out, whether or not the
    # early-exit prune lets the kernels skip them (ncu `traffic` shows what is actually moved).
    # fp64 side (k_dp / k_mid / k_xl, the O(depth*K) recurrence): algorithmic flops per column = 3 x cells + 12 x depth
    # (2 mul + 1 add per cell of the reference's recurrence, 12 flops per read for the merge),
    # cells = sum_n (min(n, K-1) + 1) + (depth - K).
    peak, peak_src = measured_peaks()
    ph = prof.mean(axis=0) * 1e-3
    t_stream = float(ph[0] + ph[2])
    bytes_algo = algorithmic_bytes(int(t["depths"].sum().item()), n, planes=n_planes)
    hbm = {"bound": "hbm", "kernel": "k_front + k_prune2", "achieved": bytes_algo / t_stream / 1e9,
           "peak": peak, "unit": "GB/s", "frac": bytes_algo / t_stream / 1e9 / peak, "traffic": ncu_traffic(STREAM_KERNELS, wl), "peak_source": peak_src,
           "note": "algorithmic bytes = every quality byte the configuration merges (SURVEY.md 8d); the early-exit prune reads far fewer "
                   "(see traffic), so a fraction above 1 means bytes skipped, not bandwidth above peak",
           "algorithmic_bytes_per_launch": bytes_algo, "kernel_ms": t_stream * 1e3}
    # heavy columns of the last batch: K and depth from the device-resident results
    cnt6 = torch.empty((n, 6), dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    capi.check(lib.lfb200_copy_counts_device(callers[0]._ctx, sts[0], C.c_void_p(cnt6.data_ptr())))
    torch.cuda.synchronize()
    kmax = cnt6[:, :3].max(dim=1).values.to(torch.int64)
    depth_t = t["depths"].to(torch.int64)
    heavy = kmax > 8
    kh, nh = kmax[heavy].double(), depth_t[heavy].double()
    cells = (kh - 1) * (kh + 2) / 2 + (nh - kh + 1) * kh + (nh - kh)
    flops_algo = float((3 * cells + 12 * nh).sum().item())
    dfma = lib.lfb200_dfma_peak(callers[0]._ctx, sts[0])
    fp64_peak = 2.0 * dfma / 1e12
    t_heavy = float(ph[3])
    fp64 = {"bound": "fp64", "kernel": "k_dp<0..2> (8 < K <= 2048, several columns per warp, bytes staged by bulk TMA) beside k_mid and k_xl (K > 2048, wavefront of 8 warps per column)", "achieved": flops_algo / t_heavy / 1e12,
            "peak": fp64_peak, "unit": "TFLOP/s", "frac": flops_algo / t_heavy / 1e12 / fp64_peak if fp64_peak else None,
            "traffic": ncu_traffic(HEAVY_KERNELS, wl), "peak_source": "DFMA microbenchmark run in this process (lfb200_dfma_peak); MEASURED_PEAKS.json has no fp64 entry",
            "algorithmic_flops_per_launch": flops_algo, "kernel_ms": t_heavy * 1e3, "columns": int(heavy.sum().item())}
    if hbm["traffic"] is not None:
        hbm["traffic_frac"] = hbm["traffic"] / t_stream / 1e9 / peak      # measured DRAM bytes / time / peak: the real pressure
    # the dominant kernel of the step is k_dp<0> (profiles/r2_summary.md), so the fp64 side is the headline roofline
    # and the stream side rides along as `other`
    roofline = dict(fp64)
    roofline["phase_ms"] = {"k_front": float(ph[0] * 1e3), "k_prune2": float(ph[2] * 1e3),
                            "heavy": float(ph[3] * 1e3)}
    roofline["other"] = hbm

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu:
            procs = 1
            v, s, kind = cpu_run(wl, args.cpu_sample, procs)
            cpu = {"value": v, "unit": UNIT, "cores": procs, "kind": kind,
                   "sample": "first %d columns of the same workload, %.1f s, single thread (the reference is "
                             "single-threaded; all-core number: --impl reference)" % (args.cpu_sample, s)}
            # the literal `lofreq call` of the metric: the reference's own prebuilt 2.1.4 binary on a synthetic BAM of the same
            # column model (BAM decode + pileup + test), without BAQ like this workload's planes
            try:
                from oracle import cli_baseline
                if cli_baseline.have_cli() and wl in DEPTH:
                    cpu["cli"] = cli_baseline.run_cli(20000 if DEPTH[wl] <= 500 else 4000, depth=DEPTH[wl], baq=False)
            except Exception as e:          # the CLI leg is a reported extra: never fails the bench
                cpu["cli"] = {"unavailable": str(e)[:200]}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_desc(wl, n) + " per GPU (region shard = rank*cols)",
                           "cols_per_gpu": n, "depth": DEPTH.get(wl), "l2": "inputs larger than L2 (%.2f GB per step)" % (2 * total / 1e9),
                           "planes": "bq+mq" + ("+baq" if args.baq else " (no BAQ plane: SURVEY 8d core benchmark; --baq adds it: 4.87 G col/s in profiles/r2_bench_C2baq.json)"),
                           "value_region": "k_front (gates, counts, running Bonferroni, prune) + O(depth*K) kernels + per-site decision (status / called / QUAL) on the device + sites in column order into pinned host memory + host wait, inputs resident in HBM; long double p-value images not requested (lfb200_set_site_pvalues 0); %d contexts = batches in flight" % NC + "",
                           "tested_columns": int(n_tested), "sites": int(n_sites), "heavy_columns": int(n_heavy),
                           "sites_all_ranks": final_sites,
                           "bonf_subst_final_this_rank": int(sm.bonf_subst_final)},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_meta, "d2h_bytes_per_step": d2h,
                        "ms_per_step": e2e_s * 1e3, "steps": e2e_steps, "copies_declared": True,
                        "h2d_bytes_mapped_per_step": n_planes * total,
                        "h2d_mode": "lfb200_call_columns with its defaults on pinned host buffers (host plane mode 1): the per-column "
                                    "metadata is copied (h2d_bytes_per_step); the quality planes (h2d_bytes_mapped_per_step) lie in "
                                    "pinned memory and are read in place over PCIe by the kernels, so only the reads that decide a "
                                    "column cross the bus; sites come back through pinned memory; identical results in both modes "
                                    "(asserted every run)",
                        "copy_mode": {"value": world * n / e2e_copy_s, "ms_per_step": e2e_copy_s * 1e3, "h2d_bytes_per_step": h2d,
                                      "h2d_mode": "host plane mode 0 (lfb200_set_host_planes): every plane is copied to the device "
                                                  "first; PCIe-bound (the round-1 headline)"}},
                "gpu_launches": (KERNELS_PER_STEP + (2 if world > 1 else 0)) * args.steps, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
                "wall_ms_per_step": wall * 1e3 / args.steps}
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    for c_ in callers:
        c_.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2", choices=["C2", "C3", "C4", "C5"])
    ap.add_argument("--cols", type=int, default=1_000_000, help="columns per GPU and step (batch size)")
    ap.add_argument("--baq", action="store_true", help="add a BAQ plane (general 4-way merge) to the workload")
    ap.add_argument("--cpu-sample", type=int, default=200_000, help="columns of the single-thread cpu_baseline sample")
    ap.add_argument("--ref-cols-per-proc", type=int, default=30_000)
    ap.add_argument("--contexts", type=int, default=0, help="contexts (batches in flight) per GPU; default 2")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer e2e measurement (profiling runs)")
    args = ap.parse_args()
    claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
