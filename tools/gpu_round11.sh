#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -k "linear or pool" -q --timeout 100 --timeout-method=thread -p no:cacheprovider > gpurun_out/test_k11.log 2>&1; echo "kernels exit=$?"; tail -5 gpurun_out/test_k11.log
timeout 400 python -m pytest tests/test_gpu_nav.py -q --timeout 150 --timeout-method=thread -p no:cacheprovider > gpurun_out/test_nav11.log 2>&1; echo "nav exit=$?"; tail -3 gpurun_out/test_nav11.log
timeout 600 python tools/microbench.py > gpurun_out/microbench11.log 2>&1; echo "micro exit=$?"; tail -17 gpurun_out/microbench11.log
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench6.json 2> gpurun_out/bench6.err; echo "bench exit=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench6.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'])
print(d['roofline']['achieved'], d['roofline']['frac'], d['roofline_pool']['achieved'], d['roofline_pool']['frac']); print(d['kernel_ms_per_step'])
PY
tail -5 gpurun_out/bench6.err
