#!/bin/bash
# quick GPU pass: parity tests (verbose timing of the slowest) + smoke + a short bench (+ optional ncu launch list: NCU=1)
mkdir -p gpurun_out
nproc > gpurun_out/gpu.txt; nvidia-smi -L >> gpurun_out/gpu.txt
( time timeout 1500 python -m pytest tests -m gpu -q -s --durations=8 ${PYTEST_ARGS:-} ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "vs (reference|port)|multi-GPU|passed|failed|error|rc=|s call|FAILED" gpurun_out/pytest_gpu.log | tail -30
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -2 gpurun_out/smoke.log
LFB200_HOST_TIMING=1 timeout 600 python bench.py --steps ${STEPS:-100} --no-cpu > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_quick.json'))
print(d['value'], d['ms_per_step'], d['roofline']['phase_ms'], d['config']['sites'], d['config']['heavy_columns'], d['config']['tested_columns'])
print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'copy_mode', d['e2e']['copy_mode']['value'], 'frac', d['roofline']['frac'])
PY
tail -3 gpurun_out/bench_quick.err
if [ -n "${NCU:-}" ]; then
  LFB200_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_bench.log 2>&1
  echo "launch list rc=$?"
  python tools/launch_summary.py gpurun_out/launches.csv | tail -25
fi
if [ -n "${NCUFULL:-}" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"${NCUFULL}" -s ${NCUSKIP:-6} -c ${NCUCOUNT:-4} -f -o gpurun_out/prof_r2 python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_full.log 2>&1
  echo "full rc=$?"; ls -la gpurun_out/prof_r2.ncu-rep
fi
