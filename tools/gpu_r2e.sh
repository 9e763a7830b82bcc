#!/bin/bash
# round 2, run E: pooling kernel with the weighted sums on tcgen05 -- kernel tests first (guarded by a short timeout: a wrong
# barrier protocol would hang), then timing of both modes
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "pool" --timeout 120 --timeout-method=thread -p no:cacheprovider -x > gpurun_out/r2e_pool_tests.log 2>&1; echo "pool tests exit=$?"
tail -15 gpurun_out/r2e_pool_tests.log | cut -c1-300
timeout 300 python tools/microbench2.py pool > gpurun_out/r2e_pool_tc.txt 2>&1; tail -4 gpurun_out/r2e_pool_tc.txt | cut -c1-900
GRIDMM_POOL_HMMA=1 timeout 300 python tools/microbench2.py pool > gpurun_out/r2e_pool_hmma.txt 2>&1; tail -4 gpurun_out/r2e_pool_hmma.txt | cut -c1-900
