#!/bin/bash
# round 2, run D: 384-wide pair tiles + two-chain fusion encoder: kernel tests, nav tests, bench A/B (chains 1 vs 2)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider -s > gpurun_out/r2d_tests.log 2>&1; echo "gpu tests exit=$?"
grep -E "passed|failed" gpurun_out/r2d_tests.log | tail -3
grep -E "^(FAILED|ERROR)|^E  " gpurun_out/r2d_tests.log | head -40
grep -E "packed vs padded" gpurun_out/r2d_tests.log | cut -c1-500
for ch in 2 1 3; do
GRIDMM_FUSION_CHAINS=$ch timeout 600 python bench.py --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/r2d_bench_ch$ch.json 2> gpurun_out/r2d_bench_ch$ch.err; echo "bench chains=$ch exit=$?"; tail -2 gpurun_out/r2d_bench_ch$ch.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2d_bench_ch$ch.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], 'serial', d['e2e']['serial_value'], d['e2e']['serial_ms_per_step'])
print('gemm', d['roofline']['achieved'], d['roofline']['frac'], 'pool', d['roofline_pool']['achieved'], d['roofline_pool']['frac']); print(d['kernel_ms_per_step'])
PY
done
GRIDMM_GEMM_384=0 timeout 600 python bench.py --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/r2d_bench_no384.json 2> gpurun_out/r2d_bench_no384.err; python -c "
import json
d=json.loads(open('gpurun_out/r2d_bench_no384.json').read().strip().splitlines()[-1]); print('no384', d['value'], d['ms_per_step'])"
