#!/bin/bash
mkdir -p gpurun_out
for v in 0 1; do
  GRIDMM_PDL=$v timeout 600 python bench.py --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/pdl${v}_bench.json 2> gpurun_out/pdl${v}_bench.err; echo "PDL=$v bench exit=$?"; tail -2 gpurun_out/pdl${v}_bench.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/pdl${v}_bench.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], d['e2e'].get('serial_ms_per_step'))
except Exception as e: print('no json', e)
PY
done
GRIDMM_PDL=1 timeout 900 python -m pytest tests/test_gpu_nav.py -m gpu -q -x --timeout 300 --timeout-method=thread -p no:cacheprovider > gpurun_out/pdl1_nav.log 2>&1; echo "nav tests with PDL exit=$?"; grep -E "passed|failed" gpurun_out/pdl1_nav.log | tail -1
