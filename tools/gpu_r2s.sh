#!/bin/bash
# pool kernel A/B: single fp16 weights (default now) vs value + residual; kernel tests; nav parity under the default
mkdir -p gpurun_out
tag=${1:-r2s}
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "pool" --timeout 120 --timeout-method=thread -p no:cacheprovider > gpurun_out/${tag}_tests.log 2>&1; echo "pool tests exit=$?"
grep -E "passed|failed" gpurun_out/${tag}_tests.log | tail -2; grep -E "^(FAILED|ERROR)|^E  " gpurun_out/${tag}_tests.log | head -20 | cut -c1-300
timeout 300 python tools/pool_probe.py ${tag} 2>&1 | grep -v "^  tile\|^CTA\|^plan\|^events, 1"
timeout 900 python -m pytest tests/test_gpu_nav.py -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider -s > gpurun_out/${tag}_nav.log 2>&1; echo "nav tests exit=$?"
grep -E "passed|failed" gpurun_out/${tag}_nav.log | tail -2; grep -E "^(FAILED|ERROR)|^E  " gpurun_out/${tag}_nav.log | head -20 | cut -c1-300
grep -i "logit" gpurun_out/${tag}_nav.log | head -20 | cut -c1-200
