#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --timeout 150 --timeout-method=thread -p no:cacheprovider > gpurun_out/test_all13.log 2>&1; echo "all gpu tests exit=$?"; tail -4 gpurun_out/test_all13.log
timeout 600 python tools/microbench.py gemm > gpurun_out/microbench13.log 2>&1; echo "micro exit=$?"; grep -E "^GEMM" gpurun_out/microbench13.log | cut -c1-200
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench8.json 2> gpurun_out/bench8.err; echo "bench exit=$?"; python - <<PY
import json
d=json.loads(open('gpurun_out/bench8.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'])
print(d['roofline']['achieved'], d['roofline']['frac'], d['roofline_pool']['achieved'], d['roofline_pool']['frac']); print(d['kernel_ms_per_step'])
PY
tail -5 gpurun_out/bench8.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attn_kernel|layernorm_kernel" -s 29 -c 10 -o gpurun_out/prof_attn python tools/prof_pool.py > gpurun_out/ncu_attn.log 2>&1; echo "ncu exit=$?"
