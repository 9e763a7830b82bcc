#!/bin/bash
# round 2, run C: packed map sequence -- kernel tests, nav tests, bench, host time
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider -s > gpurun_out/r2c_tests.log 2>&1; echo "gpu tests exit=$?"
grep -E "passed|failed" gpurun_out/r2c_tests.log | tail -3
grep -E "^(FAILED|ERROR)|^E  " gpurun_out/r2c_tests.log | head -40
grep -E "trained-scale|packed vs padded" gpurun_out/r2c_tests.log | cut -c1-700
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c_smoke.log 2>&1; echo "smoke exit=$?"; tail -1 gpurun_out/r2c_smoke.log
timeout 600 python bench.py --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench exit=$?"; tail -3 gpurun_out/r2c_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2c_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], 'serial', d['e2e']['serial_value'], d['e2e']['serial_ms_per_step'])
print('gemm', d['roofline']['achieved'], d['roofline']['frac'], d['roofline'].get('frac_valid_rows'), 'pool', d['roofline_pool']['achieved'], d['roofline_pool']['frac']); print(d['kernel_ms_per_step'])
PY
timeout 300 python tools/host_time.py > gpurun_out/r2c_host_time.txt 2>&1; head -4 gpurun_out/r2c_host_time.txt
