#!/bin/bash
# pretraining step: gradient-parity tests, then the bench with the eager attention cores and with torch's fused attention
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider -s > gpurun_out/train_tests.log 2>&1; echo "train tests exit=$?"
grep -E "passed|failed" gpurun_out/train_tests.log | tail -1; grep -E "^(FAILED|ERROR)|^E  |rel" gridmm 2>/dev/null; grep -iE "rel|err" gpurun_out/train_tests.log | head -8 | cut -c1-250
for v in 0 1; do
  GRIDMM_TRAIN_SDPA=$v timeout 600 python bench.py --workload pretrain --steps 20 --warmup 4 > gpurun_out/train_sdpa$v.json 2> gpurun_out/train_sdpa$v.err; echo "SDPA=$v exit=$?"
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/train_sdpa$v.json').read().splitlines() if l.startswith('{')][-1])
    print($v, {k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d.get('phases'))
except Exception as e: print('no json', e)
PY
done
