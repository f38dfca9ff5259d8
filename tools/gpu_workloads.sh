#!/bin/bash
# bench lines + ncu launch lists for the non-default workloads (C3, C5, C4, C2 with BAQ); outputs under gpurun_out/
mkdir -p gpurun_out
for wl in ${1:-C3 C5}; do
  extra=""; [ "$wl" = "C2baq" ] && { wl=C2; extra="--baq"; tag=C2baq; } || tag=$wl
  timeout 900 python bench.py --workload $wl $extra --steps ${STEPS:-20} --cpu-sample ${CPUS:-20000} > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "$tag bench rc=$?"
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_$tag.json'))
    r=d['roofline']
    print('$tag value %.4g col/s  ms/step %.4f  phases %s  frac %.3f  e2e %.4g  sites %d heavy %d cpu %s' % (d['value'], d['ms_per_step'], {k:round(v,4) for k,v in r['phase_ms'].items()}, r['frac'], d['e2e']['value'], d['config']['sites'], d['config']['heavy_columns'], (d.get('cpu_baseline') or {}).get('value')))
except Exception as e:
    print('$tag: no line', e)
PY
  tail -2 gpurun_out/bench_$tag.err
  if [ -n "${NCU:-}" ]; then
    LFB200_NO_GRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --workload $wl $extra --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_$tag.log 2>&1
    python tools/launch_summary.py gpurun_out/launches_$tag.csv 2>/dev/null | head -16
  fi
done
