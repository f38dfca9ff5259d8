#!/bin/bash
# refresh of the round's C2 evidence with the final code (bench line, launch list, --set full extracts) + C3 / C5 / C2+BAQ bench lines
mkdir -p gpurun_out
STEPS=500 CPUS=20000 bash tools/gpu_profile_round.sh "C2"
LFB200_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_C2.csv python bench.py --workload C2 --steps 2 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r2_launches_C2.csv > gpurun_out/r2_launches_C2.txt 2>&1; head -16 gpurun_out/r2_launches_C2.txt
for wl in C3 C5; do timeout 900 python bench.py --workload $wl --steps 100 --cpu-sample 20000 > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; echo "$wl rc=$?"; done
timeout 900 python bench.py --workload C2 --baq --steps 200 --cpu-sample 20000 > gpurun_out/bench_C2baq.json 2> gpurun_out/bench_C2baq.err; echo "C2baq rc=$?"
python - <<'PY'
import json
for wl in ["C2","C3","C5","C2baq"]:
    try:
        d=json.load(open("gpurun_out/bench_%s.json" % wl)); r=d["roofline"]
        print(wl, "value %.4g ms %.4f e2e %.4g frac %.3f traffic %s other_frac %.3f clocks %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], r["frac"], r.get("traffic"), r["other"]["frac"], d["clocks"]))
    except Exception as e:
        print(wl, "no line", e)
PY
