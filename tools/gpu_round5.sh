#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -k "pool" -q --timeout 100 --timeout-method=thread -p no:cacheprovider > gpurun_out/test_pool5.log 2>&1; echo "pool exit=$?"; tail -5 gpurun_out/test_pool5.log
timeout 400 python -m pytest tests/test_gpu_nav.py -q --timeout 150 --timeout-method=thread -p no:cacheprovider > gpurun_out/test_nav5.log 2>&1; echo "nav exit=$?"; tail -3 gpurun_out/test_nav5.log
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench4.json 2> gpurun_out/bench4.err; echo "bench exit=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench4.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'])
print(d['roofline']); print(d['roofline_pool']); print(d['kernel_ms_per_step'])
PY
tail -5 gpurun_out/bench4.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"pool_kernel|gemm_f16" -s 44 -c 8 -o gpurun_out/prof_r1 python tools/prof_kernels.py > gpurun_out/ncu_prof.log 2>&1; echo "ncu exit=$?"; tail -3 gpurun_out/ncu_prof.log
ls -la gpurun_out/*.ncu-rep
