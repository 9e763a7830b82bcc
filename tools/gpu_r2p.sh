#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "attention" --timeout 300 --timeout-method=thread -p no:cacheprovider > gpurun_out/r2p_tests.log 2>&1; echo "kernel tests exit=$?"
grep -E "passed|failed" gpurun_out/r2p_tests.log | tail -2; grep -E "^(FAILED|ERROR)|^E  " gpurun_out/r2p_tests.log | head -20 | cut -c1-300
timeout 1200 python -m pytest tests/test_gpu_nav.py -m gpu -q -x --timeout 600 --timeout-method=thread -p no:cacheprovider > gpurun_out/r2p_nav.log 2>&1; echo "nav tests exit=$?"
grep -E "passed|failed" gpurun_out/r2p_nav.log | tail -2; grep -E "^(FAILED|ERROR)|^E  " gpurun_out/r2p_nav.log | head -20 | cut -c1-300
timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; echo "bench exit=$?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2p_bench.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['e2e'].get('serial_ms_per_step'))
    print(d.get('kernel_breakdown') or d.get('breakdown'))
except Exception as e: print('no json', e)
PY
