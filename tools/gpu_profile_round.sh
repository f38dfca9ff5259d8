#!/bin/bash
# the round's ncu evidence: launch list of a short bench run + one --set full capture of the 15 kernels of one step
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_bench.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_screen|k_scan_blocks|k_finalize|k_prune2|k_mid|k_pk_prep|k_packed|k_heavy" -s 45 -c 15 -f -o gpurun_out/prof_step python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"; ls -la gpurun_out/prof_step.ncu-rep
