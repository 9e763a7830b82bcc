#!/bin/bash
# Run the GPU test groups in separate processes so that one hung kernel cannot take the others down.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; timeout 420 python -m pytest "$@" -q --timeout 150 --timeout-method=thread -p no:cacheprovider > gpurun_out/test_$name.log 2>&1; echo "$name exit=$?" >> gpurun_out/summary.txt; tail -3 gpurun_out/test_$name.log; }
rm -f gpurun_out/summary.txt
run linear tests/test_gpu_kernels.py -k "linear"
run attn tests/test_gpu_kernels.py -k "attention or layernorm or pos_embed"
run grid tests/test_gpu_kernels.py -k "grid_update"
run pool tests/test_gpu_kernels.py -k "pool"
run nav tests/test_gpu_nav.py -s
cat gpurun_out/summary.txt
