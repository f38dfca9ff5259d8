#!/bin/bash
# like gpu_quick.sh with tight timeouts (a hung kernel must not eat the GPU budget) + the other workloads
mkdir -p gpurun_out
( timeout 200 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for w in C2:1000000 C3:200000 C5:200000; do
  wl=${w%%:*}; n=${w##*:}
  timeout 120 python bench.py --workload $wl --cols $n --steps 50 --no-e2e --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(\"$wl\", round(d[\"value\"]/1e6,1), \"Mcol/s\", round(d[\"ms_per_step\"],3), {k:round(v,3) for k,v in d[\"roofline\"][\"phase_ms\"].items()}, d[\"config\"][\"sites\"], d[\"config\"][\"heavy_columns\"])"
done
