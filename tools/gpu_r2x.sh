#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "pool" --timeout 120 --timeout-method=thread -p no:cacheprovider > gpurun_out/r2x_tests.log 2>&1; echo "pool tests exit=$?"
grep -E "passed|failed" gpurun_out/r2x_tests.log | tail -2; grep -E "^(FAILED|ERROR)|^E  " gpurun_out/r2x_tests.log | head -20 | cut -c1-300
for e in 0 5 3 4; do
  echo "=== exp $e (0: shipped, 5: one tile per MMA, 3: half the contraction, 4: a twelfth)"
  POOL_EXP=$e timeout 300 python tools/pool_probe.py x$e 2>&1 | grep "^events, 8\|^valid rows\|^prod_tot\|^setup"
  python tools/pool_trace_summary.py x$e | head -2
done
