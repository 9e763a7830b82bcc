#!/bin/bash
# graph-timed microbenchmarks + ncu launch list + ncu --set full (reports exported to CSV on the box; only small reps travel back)
mkdir -p gpurun_out
timeout 600 python tools/microbench2.py > gpurun_out/microbench22.log 2>&1; echo "micro exit=$?"; cat gpurun_out/microbench22.log | cut -c1-1200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches_r1b.csv python tools/prof_pool.py > gpurun_out/ncu_l22.log 2>&1; echo "ncu launches exit=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"pool_kernel" -s 2 -c 1 -o gpurun_out/prof_r1b_pool -f python tools/prof_pool.py > gpurun_out/ncu_f22a.log 2>&1; echo "ncu pool exit=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attn_kernel" -s 22 -c 5 -o gpurun_out/prof_r1b_attn -f python tools/prof_pool.py > gpurun_out/ncu_f22b.log 2>&1; echo "ncu attn exit=$?"
timeout 900 ncu --set full --clock-control none -k regex:"gemm_f16" -s 84 -c 42 -o /tmp/prof_r1b_gemm -f python tools/prof_pool.py > gpurun_out/ncu_f22c.log 2>&1; echo "ncu gemm exit=$?"
ncu -i /tmp/prof_r1b_gemm.ncu-rep --page raw --csv > gpurun_out/prof_r1b_gemm_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none -k regex:"layernorm|grid_update|grid_assemble|embed_kernel" -s 44 -c 22 -o /tmp/prof_r1b_rows -f python tools/prof_pool.py > gpurun_out/ncu_f22d.log 2>&1; echo "ncu rows exit=$?"
ncu -i /tmp/prof_r1b_rows.ncu-rep --page raw --csv > gpurun_out/prof_r1b_rows_raw.csv 2>/dev/null
du -sh gpurun_out; ls -la gpurun_out
