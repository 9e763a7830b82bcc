#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 python -m pytest "$@" -q -x --timeout 150 --timeout-method=thread -p no:cacheprovider > gpurun_out/test_$name.log 2>&1; echo "$name exit=$?"; tail -4 gpurun_out/test_$name.log; }
run nav tests/test_gpu_nav.py -s
grep -E "^B=" gpurun_out/test_nav.log | cut -c1-400
run kernels tests/test_gpu_kernels.py
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench28.json 2> gpurun_out/bench28.err; echo "bench exit=$?"; python - <<PY
import json
d=json.loads(open('gpurun_out/bench28.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'])
print(d['roofline']['achieved'], d['roofline']['frac'], d['roofline_pool']['achieved'], d['roofline_pool']['frac']); print(d['kernel_ms_per_step'])
PY
tail -5 gpurun_out/bench28.err
