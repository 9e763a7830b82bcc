#!/bin/bash
# same-box A/B of the pooling kernel: _ab/ holds a build of the last commit (git archive HEAD | tar -x -C _ab; build there), the
# working tree is the candidate.  Boxes differ by several us, so only same-call comparisons count.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "pool" --timeout 120 --timeout-method=thread -p no:cacheprovider > gpurun_out/ab_tests.log 2>&1; echo "pool tests exit=$?"
grep -E "passed|failed" gpurun_out/ab_tests.log | tail -2; grep -E "^(FAILED|ERROR)|^E  " gpurun_out/ab_tests.log | head -20 | cut -c1-300
for r in 1 2; do
  echo "=== last commit, run $r"; (cd _ab && timeout 300 python tools/pool_probe.py ab_old 2>&1 | grep "^events, 8\|^valid rows" | cut -c1-400)
  echo "=== working tree, run $r"; timeout 300 python tools/pool_probe.py ab_new 2>&1 | grep "^events, 8\|^valid rows\|^prod_tot" | cut -c1-500
done
python tools/pool_trace_summary.py ab_new 2>/dev/null | head -2
