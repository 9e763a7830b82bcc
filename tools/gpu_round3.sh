#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_nav.py -q -s --timeout 150 --timeout-method=thread -p no:cacheprovider > gpurun_out/test_nav3.log 2>&1; echo "nav exit=$?"; grep -E "^(B=|r2r|reverie|\.B=|\.r)" gpurun_out/test_nav3.log | cut -c1-400; tail -3 gpurun_out/test_nav3.log
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench2.json 2> gpurun_out/bench2.err; echo "bench exit=$?"; tail -c 2500 gpurun_out/bench2.json; tail -5 gpurun_out/bench2.err
