#!/bin/bash
mkdir -p gpurun_out
tag=${1:-q2}
timeout 900 python -m pytest tests/test_gpu_nav.py tests/test_gpu_kernels.py -m gpu -q -x --timeout 300 --timeout-method=thread -p no:cacheprovider > gpurun_out/${tag}_tests.log 2>&1; echo "nav+kernel tests exit=$?"
grep -E "passed|failed" gpurun_out/${tag}_tests.log | tail -1; grep -E "^(FAILED|ERROR)|^E  " gpurun_out/${tag}_tests.log | head -20 | cut -c1-300
timeout 900 python bench.py --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench exit=$?"; tail -3 gpurun_out/${tag}_bench.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${tag}_bench.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], d['e2e'].get('serial_ms_per_step'))
    print('roofline', d['roofline']['frac'], d['roofline']['ms'], 'pool', {k:d['roofline_pool'][k] for k in ('frac','ms','ms_back_to_back','achieved')})
    print(d['kernel_ms_per_step'])
except Exception as e: print('no json', e)
PY
