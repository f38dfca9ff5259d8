#!/bin/bash
# quick GPU pass: parity tests + a short bench, optionally the A/B run without k_packed
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 100 --no-cpu > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_quick.json'))
print(d['value'], d['ms_per_step'], d['roofline']['phase_ms'], d['config']['sites'], d['config']['heavy_columns'], d['config']['tested_columns'])
PY
tail -3 gpurun_out/bench_quick.err
if [ "${1:-}" = "ab" ]; then
LFB200_NO_PACKED=1 timeout 600 python bench.py --steps 20 --no-cpu > gpurun_out/bench_nopk.json 2>/dev/null
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_nopk.json'))
print('nopacked', d['value'], d['ms_per_step'], d['roofline']['phase_ms'], d['config']['sites'], d['config']['heavy_columns'])
PY
fi
