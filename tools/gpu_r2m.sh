#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q --timeout 600 --timeout-method=thread -p no:cacheprovider -s > gpurun_out/r2m_tests.log 2>&1; echo "tests exit=$?"
grep -E "passed|failed" gpurun_out/r2m_tests.log | tail -2; grep -E "^(FAILED|ERROR)|^E  " gpurun_out/r2m_tests.log | head -20 | cut -c1-400
grep -E "gradient parity" gpurun_out/r2m_tests.log | cut -c1-400
timeout 900 python bench.py --workload pretrain --steps 20 --warmup 4 > gpurun_out/r2m_pretrain.json 2> gpurun_out/r2m_pretrain.err; echo "bench exit=$?"
tail -3 gpurun_out/r2m_pretrain.err; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2m_pretrain.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['e2e']['ms_per_step'], d['phases'], d['clocks'])
except Exception as e: print('no json', e)
PY
timeout 600 python tools/prof_pretrain.py 2>&1 | grep -v "^---" | cut -c1-60,118-250 | tail -50
