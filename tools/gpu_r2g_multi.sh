#!/bin/bash
# round 2, run G (8 GPUs of one box): scaling of the navigation step (episodes sharded, no data-path collective) and the
# pretraining gradient step (flat NCCL all-reduce + clip + AdamW)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
for n in 1 2 4 8; do
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/r2g_bench_n1.json 2> gpurun_out/r2g_bench_n1.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/r2g_bench_n$n.json 2> gpurun_out/r2g_bench_n$n.err
  fi
  echo "bench n=$n exit=$?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2g_bench_n$n.json').read().strip().splitlines()[-1])
    print($n, d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'serial', d['e2e']['serial_value'])
except Exception as e:
    print('parse failed', e)
PY
done
for n in 2 4 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) tools/bench_gradstep.py > gpurun_out/r2g_gradstep_n$n.json 2> gpurun_out/r2g_gradstep_n$n.err
  echo "gradstep n=$n exit=$?"; tail -1 gpurun_out/r2g_gradstep_n$n.json
done
