#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 python -m pytest "$@" -q -x --timeout 150 --timeout-method=thread -p no:cacheprovider > gpurun_out/test_$name.log 2>&1; echo "$name exit=$?"; tail -4 gpurun_out/test_$name.log; }
run lnfused tests/test_gpu_kernels.py -k "linear_ln or attention"
grep -E "assert|Error" gpurun_out/test_lnfused.log | head -5
timeout 300 python tools/microbench2.py gemmln > gpurun_out/microbench31.log 2>&1; echo "micro exit=$?"; cat gpurun_out/microbench31.log | cut -c1-200
run nav tests/test_gpu_nav.py
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench31.json 2> gpurun_out/bench31.err; echo "bench exit=$?"; python - <<PY
import json
d=json.loads(open('gpurun_out/bench31.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'])
print(d['kernel_ms_per_step'])
PY
tail -5 gpurun_out/bench31.err
