#!/bin/bash
# one --set full capture of the kernels matching $1 (regex), launch skip $2, count $3, output name $4; extra bench args in $BENCH_ARGS
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$1" -s ${2:-3} -c ${3:-1} -f -o gpurun_out/prof_${4:-k} python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e ${BENCH_ARGS:-} > gpurun_out/ncu_full.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/*.ncu-rep
