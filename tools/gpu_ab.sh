#!/bin/bash
# A/B of launch geometry through environment knobs: each line = one short bench run
mkdir -p gpurun_out
run() { env "$@" timeout 300 python bench.py --steps 60 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$*', '%.3g col/s' % d['value'], 'ms/step %.4f' % d['ms_per_step'], {k:round(v,4) for k,v in d['roofline']['phase_ms'].items()})"; }
run A=1
run LFB200_DP0_CTAS=4
run LFB200_DP0_CTAS=4 LFB200_FRONT_CTAS=2
run LFB200_DP0_CTAS=3 LFB200_FRONT_CTAS=2
run LFB200_DP0_CTAS=5 LFB200_FRONT_CTAS=2
run LFB200_FRONT_CTAS=2
run LFB200_DP0_CTAS=4 LFB200_FRONT_CTAS=1
