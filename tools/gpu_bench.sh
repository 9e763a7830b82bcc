#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 200 --warmup 5 > gpurun_out/bench37.json 2> gpurun_out/bench37.err; echo "bench exit=$?"; tail -3 gpurun_out/bench37.err
