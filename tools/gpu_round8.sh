#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -k "pool" -q --timeout 100 --timeout-method=thread -p no:cacheprovider > gpurun_out/test_k8.log 2>&1; echo "pool tests exit=$?"; tail -4 gpurun_out/test_k8.log
timeout 600 python tools/microbench.py pool > gpurun_out/microbench8.log 2>&1; echo "micro exit=$?"; tail -4 gpurun_out/microbench8.log
