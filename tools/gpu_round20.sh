#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/microbench.py gemm pool > gpurun_out/microbench20.log 2>&1; echo "micro exit=$?"; grep -E "^GEMM|^---|^pool|^mode 0" gpurun_out/microbench20.log | cut -c1-140
timeout 900 python -m pytest tests -m gpu -q -x --timeout 150 --timeout-method=thread -p no:cacheprovider > gpurun_out/test_all20.log 2>&1; echo "all gpu tests exit=$?"; tail -3 gpurun_out/test_all20.log
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench10.json 2> gpurun_out/bench10.err; echo "bench exit=$?"; python - <<PY
import json
d=json.loads(open('gpurun_out/bench10.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'])
print(d['roofline']['achieved'], d['roofline']['frac'], d['roofline_pool']['achieved'], d['roofline_pool']['frac']); print(d['kernel_ms_per_step'])
PY
tail -5 gpurun_out/bench10.err
