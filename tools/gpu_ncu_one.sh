#!/bin/bash
# one --set full capture of the kernels matching $1 (regex) in a bench run, reduced on the box: metrics JSON + per-line summaries
# usage: gpu_ncu_one.sh REGEX TAG [skip] [count] ; extra bench args in $BENCH_ARGS, line summaries for kernels in $LINES
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$1" -s ${3:-6} -c ${4:-2} -f -o /tmp/prof_$2 python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e ${BENCH_ARGS:-} > gpurun_out/ncu_full_$2.log 2>&1
echo "ncu rc=$?"
python tools/ncu_extract.py /tmp/prof_$2.ncu-rep gpurun_out/ncu_$2.json
for k in ${LINES:-}; do python tools/ncu_lines.py /tmp/prof_$2.ncu-rep $k 30 > gpurun_out/lines_$2_$k.txt 2>&1; done
rm -f /tmp/prof_$2.ncu-rep
