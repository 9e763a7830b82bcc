#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_nav.py -m gpu -q -k "feature_db or gmap130 or batch1 or batch64 or lazy or active or staged" --timeout 300 --timeout-method=thread -p no:cacheprovider -s > gpurun_out/r2h_tests.log 2>&1; echo "tests exit=$?"
grep -E "passed|failed" gpurun_out/r2h_tests.log | tail -2; grep -E "^(FAILED|ERROR)|^E  " gpurun_out/r2h_tests.log | head -20
timeout 900 python bench.py --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo "bench exit=$?"; tail -3 gpurun_out/r2h_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2h_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'])
PY
timeout 600 python tools/microbench2.py gemm > gpurun_out/r2h_microbench_gemm.txt 2>&1; tail -12 gpurun_out/r2h_microbench_gemm.txt
