#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 python -m pytest "$@" -q -x --timeout 150 --timeout-method=thread -p no:cacheprovider > gpurun_out/test_$name.log 2>&1; echo "$name exit=$?"; tail -4 gpurun_out/test_$name.log; }
run pool tests/test_gpu_kernels.py -k "pool"
grep -E "assert|Error" gpurun_out/test_pool.log | head -8
