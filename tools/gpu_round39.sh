#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 300 python -m pytest "$@" -q -x --timeout 100 --timeout-method=thread -p no:cacheprovider > gpurun_out/test_$name.log 2>&1; echo "$name exit=$?"; tail -3 gpurun_out/test_$name.log; grep -E "^E " gpurun_out/test_$name.log | head -5; }
run fus tests/test_gpu_nav.py -k "average_fusion or staged"
run lnf tests/test_gpu_kernels.py -k "linear_ln"
timeout 300 python tools/microbench2.py gemmln > gpurun_out/microbench39.log 2>&1; echo "micro exit=$?"; cat gpurun_out/microbench39.log | cut -c1-160
