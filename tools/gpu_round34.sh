#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 150 --timeout-method=thread -p no:cacheprovider > gpurun_out/test_all34.log 2>&1; echo "all gpu tests exit=$?"; tail -3 gpurun_out/test_all34.log; grep -E "^E " gpurun_out/test_all34.log | head
