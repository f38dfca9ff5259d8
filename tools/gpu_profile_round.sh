#!/bin/bash
# the round's evidence in ONE gpurun: per workload the bench line and one --set full capture of the 13 kernels of one step,
# reduced ON THE BOX (the .ncu-rep files are too big to travel) to profiles-ready JSON (tools/ncu_extract.py; bench.py reads
# it for roofline.traffic) and per-source-line summaries of the dominant kernels (tools/ncu_lines.py)
mkdir -p gpurun_out
KREG='^k_(front|scan_tiles|prune2|mid|dp|xl|heavy_xl|heavy_all|scan_blocks|rank_cands|emit_sites)'
for wl in ${1:-C2 C3 C5}; do
  timeout 900 python bench.py --workload $wl --steps ${STEPS:-200} --cpu-sample ${CPUS:-20000} > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; echo "$wl bench rc=$?"
  LFB200_NO_GRAPH=1 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$KREG" -s 26 -c 13 -f -o /tmp/prof_step_$wl python bench.py --workload $wl --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_full_$wl.log 2>&1
  echo "$wl ncu rc=$?"
  python tools/ncu_extract.py /tmp/prof_step_$wl.ncu-rep gpurun_out/r2_step_kernels_$wl.json > gpurun_out/r2_step_kernels_$wl.txt 2>&1
  for k in k_dp k_front k_prune2 k_xl; do python tools/ncu_lines.py /tmp/prof_step_$wl.ncu-rep $k 25 > gpurun_out/r2_lines_${wl}_$k.txt 2>&1; done
  rm -f /tmp/prof_step_$wl.ncu-rep
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_ref.json | cut -c1-200
du -sh gpurun_out
