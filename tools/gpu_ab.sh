#!/bin/bash
# same-box A/B of the pooling kernel: _ab/ holds a build of the previous commit
mkdir -p gpurun_out
for r in 1 2; do
  echo "=== HEAD~ (committed) run $r"; (cd _ab && timeout 300 python tools/pool_probe.py ab_old 2>&1 | grep "^events, 8\|^valid rows\|^pool_tot\|^prod_tot" | cut -c1-400)
  echo "=== working tree run $r"; timeout 300 python tools/pool_probe.py ab_new 2>&1 | grep "^events, 8\|^valid rows\|^prod_tot\|^setup" | cut -c1-400
  echo "=== working tree, one tile per MMA, run $r"; POOL_EXP=5 timeout 300 python tools/pool_probe.py ab_new5 2>&1 | grep "^events, 8\|^valid rows" | cut -c1-400
done
