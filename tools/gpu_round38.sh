#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_nav.py -k "average_fusion or staged" -q -x --timeout 150 --timeout-method=thread -p no:cacheprovider > gpurun_out/test_38.log 2>&1; echo "exit=$?"; tail -3 gpurun_out/test_38.log; grep -E "^E " gpurun_out/test_38.log | head
