#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/microbench.py > gpurun_out/microbench7.log 2>&1; echo "micro exit=$?"; cat gpurun_out/microbench7.log | tail -20
