#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "attention" --timeout 300 --timeout-method=thread -p no:cacheprovider > gpurun_out/r2q_tests.log 2>&1; echo "kernel tests exit=$?"
grep -E "passed|failed" gpurun_out/r2q_tests.log | tail -2; grep -E "^(FAILED|ERROR)|^E  " gpurun_out/r2q_tests.log | head -20 | cut -c1-300
timeout 600 python tools/microbench2.py attn 2>&1 | grep -E "ATTN|---" 
timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err; echo "bench exit=$?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2q_bench.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['e2e'].get('serial_ms_per_step'))
    print({k:v for k,v in d['kernel_ms_per_step'].items() if 'attention' in k})
except Exception as e: print('no json', e)
PY
