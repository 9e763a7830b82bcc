#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -k "linear" -q --timeout 100 --timeout-method=thread -p no:cacheprovider > gpurun_out/test_lin4.log 2>&1; echo "linear exit=$?"; tail -3 gpurun_out/test_lin4.log
timeout 400 python -m pytest tests/test_gpu_nav.py -q -s --timeout 150 --timeout-method=thread -p no:cacheprovider > gpurun_out/test_nav4.log 2>&1; echo "nav exit=$?"; tail -3 gpurun_out/test_nav4.log
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench3.json 2> gpurun_out/bench3.err; echo "bench exit=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench3.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'])
print(d['roofline']); print(d['roofline_pool']); print(d['kernel_ms_per_step'])
PY
tail -5 gpurun_out/bench3.err
