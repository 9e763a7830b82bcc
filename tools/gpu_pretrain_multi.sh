#!/bin/bash
# pretraining step at N GPUs (N = $1): bench line under gpurun_out/
N=${1:-8}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 bench.py --workload pretrain --gpus $N --steps 20 --warmup 4 > gpurun_out/r2o_pretrain_n$N.json 2> gpurun_out/r2o_pretrain_n$N.err; echo "exit=$?"
tail -2 gpurun_out/r2o_pretrain_n$N.err | cut -c1-300
python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2o_pretrain_n$N.json').read().splitlines() if l.startswith('{')][-1])
    print($N, {k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['e2e']['ms_per_step'], d['phases'])
except Exception as e: print('no json', e)
PY
if [ "$N" = "8" ]; then
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29572 bench.py --workload pretrain --gpus 2 --steps 20 --warmup 4 > gpurun_out/r2o_pretrain_n2.json 2> gpurun_out/r2o_pretrain_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29573 bench.py --workload pretrain --gpus 4 --steps 20 --warmup 4 > gpurun_out/r2o_pretrain_n4.json 2> gpurun_out/r2o_pretrain_n4.err
timeout 600 python bench.py --workload pretrain --gpus 1 --steps 20 --warmup 4 > gpurun_out/r2o_pretrain_n1.json 2> gpurun_out/r2o_pretrain_n1.err
for n in 1 2 4; do python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2o_pretrain_n$n.json').read().splitlines() if l.startswith('{')][-1])
    print($n, {k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['e2e']['ms_per_step'], d['phases'])
except Exception as e: print('no json', e)
PY
done
fi
