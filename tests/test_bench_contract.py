"""bench.py's reference arm (the CPU implementation of the path on the host cores) prints exactly one JSON line on
stdout with the keys the driver reads; the GPU arm is covered on the GPU box."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--ref-cols-per-proc", "500"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "pileup_columns_per_sec" and d["unit"] == "columns/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["config"]["workload"].startswith("C2: 1000000 synthetic pileup columns, depth 500")


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "1", "--ref-cols-per-proc", "500"], capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_committed_bench_lines_carry_the_contract_keys():
    """the GPU arm's lines committed under profiles/ (written on the GPU box by the round's evidence scripts): every key the
    driver and the judge read is there, the roofline is internally consistent, the traffic comes from the committed capture"""
    prof = os.path.join(ROOT, "profiles")
    for wl in ("C2", "C3", "C5", "C2baq"):
        d = json.load(open(os.path.join(prof, "r2_bench_%s.json" % wl)))
        for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                  "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
            assert k in d, (wl, k)
        assert d["metric"] == "pileup_columns_per_sec" and d["unit"] == "columns/s" and d["dtype"] == "f64" and d["vs_baseline"] is None
        assert d["gpu_launches"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
        assert 0 < d["e2e"]["value"] < d["value"]
        r = d["roofline"]
        assert r["bound"] in ("fp64", "hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
        assert d["clocks"]["sm_mhz"] and not d["clocks"]["reasons"]
        assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
        # value = columns of the step / time of the step
        assert abs(d["value"] - d["n_gpus"] * d["config"]["cols_per_gpu"] / (d["ms_per_step"] * 1e-3)) / d["value"] < 1e-6


def test_traffic_comes_from_the_committed_capture():
    sys.path.insert(0, ROOT)
    import bench
    t = bench.ncu_traffic(["k_dp<0>"], "C2")
    assert t and t > 1e6
    t = bench.ncu_traffic(["k_front", "k_prune2"], "C2")
    assert t and t > 5e7
