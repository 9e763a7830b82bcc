#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -k "linear" -q --timeout 100 --timeout-method=thread -p no:cacheprovider > gpurun_out/test_k18.log 2>&1; echo "linear exit=$?"; tail -8 gpurun_out/test_k18.log
timeout 300 python tools/microbench.py gemm > gpurun_out/microbench18.log 2>&1; echo "micro exit=$?"; grep -E "^GEMM|^---" gpurun_out/microbench18.log | cut -c1-175
