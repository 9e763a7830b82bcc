#!/bin/bash
# round 2, run F: single fp16 softmax weights in the pooling sums: accuracy (nav oracle tests) and speed; AdamW test; gradient step
mkdir -p gpurun_out
GRIDMM_POOL_SPLIT=0 timeout 900 python -m pytest tests/test_gpu_nav.py tests/test_gpu_kernels.py -m gpu -q -k "oracle or golden or pool or adamw" --timeout 300 --timeout-method=thread -p no:cacheprovider -s > gpurun_out/r2f_tests_single.log 2>&1; echo "single-weight tests exit=$?"
grep -E "passed|failed" gpurun_out/r2f_tests_single.log | tail -2; grep -E "^(FAILED|ERROR)|^E  " gpurun_out/r2f_tests_single.log | head -20
grep -E "^\.?B=" gpurun_out/r2f_tests_single.log | cut -c1-420
GRIDMM_POOL_SPLIT=0 timeout 300 python tools/microbench2.py pool > gpurun_out/r2f_pool_single.txt 2>&1; tail -2 gpurun_out/r2f_pool_single.txt | cut -c1-900
GRIDMM_POOL_SPLIT=0 timeout 600 python bench.py --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/r2f_bench_single.json 2> gpurun_out/r2f_bench_single.err; python -c "
import json
d=json.loads(open('gpurun_out/r2f_bench_single.json').read().strip().splitlines()[-1]); print('single', d['value'], d['ms_per_step'], d['roofline_pool']['frac'], d['roofline_pool']['ms'])"
timeout 600 python bench.py --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; python -c "
import json
d=json.loads(open('gpurun_out/r2f_bench.json').read().strip().splitlines()[-1]); print('split', d['value'], d['ms_per_step'], d['roofline_pool']['frac'], d['roofline_pool']['ms'])"
timeout 300 python tools/bench_gradstep.py > gpurun_out/r2f_gradstep_1gpu.json 2> gpurun_out/r2f_gradstep.err; tail -1 gpurun_out/r2f_gradstep_1gpu.json; tail -3 gpurun_out/r2f_gradstep.err
