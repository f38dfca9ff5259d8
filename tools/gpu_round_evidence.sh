#!/bin/bash
# everything profiles/ holds for the round, in one gpurun: launch lists (C2, C3, C5), bench lines, --set full extracts, the
# reference arm, and a contexts sweep
mkdir -p gpurun_out
bash tools/gpu_profile_round.sh "C2 C3 C5"
for wl in C2 C3 C5; do
  LFB200_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_$wl.csv python bench.py --workload $wl --steps 2 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1
  python tools/launch_summary.py gpurun_out/r2_launches_$wl.csv > gpurun_out/r2_launches_$wl.txt 2>&1
  head -16 gpurun_out/r2_launches_$wl.txt
done
for nc in 2 3 4; do timeout 300 python bench.py --steps 150 --no-cpu --no-e2e --contexts $nc 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('contexts $nc', '%.4g col/s' % d['value'], d['ms_per_step'])"; done
