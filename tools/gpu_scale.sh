#!/bin/bash
# multi-GPU pass on an N-GPU box: the 2-rank parity test, then bench at the given rank counts with the host-side timing
# breakdown (LFB200_HOST_TIMING).  usage: gpu_scale.sh "2 4 8"
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt; nproc >> gpurun_out/gpus.txt
( timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q -s ) > gpurun_out/pytest_mgpu.log 2>&1; echo "mgpu pytest rc=$?"
grep -E "multi-GPU|passed|failed|skipped" gpurun_out/pytest_mgpu.log | tail -5
for n in $1; do
  if [ "$n" = "1" ]; then
    LFB200_HOST_TIMING=1 timeout 600 python bench.py --gpus 1 --steps 100 --no-cpu > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err
  else
    LFB200_HOST_TIMING=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $n --steps 100 --no-cpu > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  fi
  echo "N=$n rc=$?"
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/scale_n$n.json'))
    print('N=$n value %.3g ms/step %.4f e2e %.3g wall_ms/step %.4f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['wall_ms_per_step']))
except Exception as e:
    print('N=$n no line', e)
PY
  grep "host seconds" gpurun_out/scale_n$n.err | head -8
done
