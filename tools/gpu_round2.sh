#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_nav.py -q --timeout 150 --timeout-method=thread -p no:cacheprovider > gpurun_out/test_nav2.log 2>&1; echo "nav exit=$?"
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?"; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench1.json 2> gpurun_out/bench1.err; echo "bench exit=$?"; tail -c 3000 gpurun_out/bench1.json; tail -5 gpurun_out/bench1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu exit=$?"
