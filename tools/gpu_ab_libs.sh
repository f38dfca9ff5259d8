#!/bin/bash
# A/B of library builds (tools/build_variant.py) on one box: bash tools/gpu_ab_libs.sh "base f3 f5" [workload]
mkdir -p gpurun_out
wl=${2:-C2}
for rep in $(seq 1 ${REPS:-2}); do
for v in $1; do
  lib=""; [ "$v" != base ] && lib="LFB200_LIB=$PWD/lofreq_b200/lib/var_$v.so"
  env $lib timeout 300 python bench.py --workload $wl --steps ${STEPS:-100} --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$v', '%.4g col/s' % d['value'], 'ms/step %.4f' % d['ms_per_step'], {k:round(v,4) for k,v in d['roofline']['phase_ms'].items()}, 'sites', d['config']['sites'])"
done; done
