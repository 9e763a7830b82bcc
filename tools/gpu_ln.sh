#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 400 python -m pytest "$@" -q -x --timeout 100 --timeout-method=thread -p no:cacheprovider > gpurun_out/test_$name.log 2>&1; echo "$name exit=$?"; tail -3 gpurun_out/test_$name.log; grep -E "^E " gpurun_out/test_$name.log | head -5; }
run lnf tests/test_gpu_kernels.py -k "linear_ln"
run nav tests/test_gpu_nav.py -s
grep -E "^B=" gpurun_out/test_nav.log | cut -c1-260
timeout 300 python tools/microbench2.py gemmln > gpurun_out/microbench40.log 2>&1; echo "micro exit=$?"; head -5 gpurun_out/microbench40.log | cut -c1-160
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench40.json 2> gpurun_out/bench40.err; echo "bench exit=$?"; python - <<PY
import json
d=json.loads(open('gpurun_out/bench40.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'])
print(d['kernel_ms_per_step'])
PY
