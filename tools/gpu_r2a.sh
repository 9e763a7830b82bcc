#!/bin/bash
# round 2, run A: the whole GPU suite (new parity cases included), smoke, a bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider -s > gpurun_out/r2a_tests.log 2>&1; echo "gpu tests exit=$?"
grep -E "passed|failed" gpurun_out/r2a_tests.log | tail -3
grep -E "^(FAILED|ERROR)|^E  " gpurun_out/r2a_tests.log | head -40
grep -E "errors|error |^B=|CE B=16|trained-scale" gpurun_out/r2a_tests.log | cut -c1-600 | head -40
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.log 2>&1; echo "smoke exit=$?"; tail -1 gpurun_out/r2a_smoke.log
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench exit=$?"; tail -c 1500 gpurun_out/r2a_bench.json
