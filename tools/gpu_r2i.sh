#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_train.py tests/test_gpu_nav.py -m gpu -q -k "train or gradient or linear_fn or gmap130 or feature_db" --timeout 600 --timeout-method=thread -p no:cacheprovider -s > gpurun_out/r2i_tests.log 2>&1; echo "tests exit=$?"
grep -E "passed|failed" gpurun_out/r2i_tests.log | tail -2; grep -E "^(FAILED|ERROR)|^E  " gpurun_out/r2i_tests.log | head -30
grep -E "gradient parity" gpurun_out/r2i_tests.log | cut -c1-900
