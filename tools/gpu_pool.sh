#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 420 python -m pytest "$@" -q -x --timeout 150 --timeout-method=thread -p no:cacheprovider > gpurun_out/test_$name.log 2>&1; echo "$name exit=$?"; tail -4 gpurun_out/test_$name.log; }
run pool tests/test_gpu_kernels.py -k "pool"
timeout 300 python tools/microbench2.py pool > gpurun_out/microbench24.log 2>&1; echo "micro exit=$?"; cat gpurun_out/microbench24.log | cut -c1-1500
run nav tests/test_gpu_nav.py

