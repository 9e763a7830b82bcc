#!/bin/bash
# navigation-step bench at the GPU counts listed in $1 (one box), launched the way the driver launches it
N=${1:-8}
mkdir -p gpurun_out
run() {
  n=$1
  if [ "$n" = "1" ]; then
    timeout 600 python bench.py --gpus 1 --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2958$n bench.py --gpus $n --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  fi
  echo "N=$n exit=$?"
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/scale_n$n.json').read().splitlines() if l.startswith('{')][-1])
    print($n, {k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'])
except Exception as e: print('no json', e)
PY
}
# usage: gpu_scale.sh "1 8" (a list of GPU counts, all on the box gpurun --gpus <max> provides)
for n in $N; do run $n; done
