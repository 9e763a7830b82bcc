#!/bin/bash
# full status round: GPU tests, bench (with cpu baseline), microbench, ncu launch list, ncu --set full of pool+gemm
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
ls -la MEASURED_PEAKS.json 2>/dev/null && cat MEASURED_PEAKS.json
timeout 900 python -m pytest tests -m gpu -q -x --timeout 150 --timeout-method=thread -p no:cacheprovider > gpurun_out/test_all21.log 2>&1; echo "all gpu tests exit=$?"; tail -4 gpurun_out/test_all21.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke21.log 2>&1; echo "smoke exit=$?"; tail -2 gpurun_out/smoke21.log
timeout 600 python bench.py --steps 100 --warmup 5 > gpurun_out/bench21.json 2> gpurun_out/bench21.err; echo "bench exit=$?"; cat gpurun_out/bench21.json | cut -c1-3000; tail -5 gpurun_out/bench21.err
timeout 300 python tools/microbench.py gemm pool > gpurun_out/microbench21.log 2>&1; echo "micro exit=$?"; grep -E "^GEMM|^---|^pool|^mode" gpurun_out/microbench21.log | cut -c1-260
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches_r1b.csv python tools/prof_pool.py > gpurun_out/ncu_l21.log 2>&1; echo "ncu launches exit=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pool_kernel|gemm_f16|attn_kernel" -s 108 -c 54 -o gpurun_out/prof_r1b -f python tools/prof_pool.py > gpurun_out/ncu_f21.log 2>&1; echo "ncu full exit=$?"; tail -3 gpurun_out/ncu_f21.log
ls -la gpurun_out
