#!/bin/bash
# pool kernel with the work plan (cells cut between CTAs): kernel tests, probe, nav parity
mkdir -p gpurun_out
tag=${1:-r2r}
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "pool" --timeout 120 --timeout-method=thread -p no:cacheprovider > gpurun_out/${tag}_tests.log 2>&1; echo "pool tests exit=$?"
grep -E "passed|failed" gpurun_out/${tag}_tests.log | tail -2; grep -E "^(FAILED|ERROR)|^E  " gpurun_out/${tag}_tests.log | head -20 | cut -c1-300
timeout 300 python tools/pool_probe.py ${tag} 2>&1 | tail -12
