#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none -k regex:"gemm_f16|gemm_ln|attn_|head_rows|nav_logits2|fusion_inputs|grid_update|embed_kernel|grid_assemble|layernorm" -s 123 -c 58 -o /tmp/prof_r1c_all -f python tools/prof_pool.py > gpurun_out/ncu_f35.log 2>&1; echo "ncu all exit=$?"
ncu -i /tmp/prof_r1c_all.ncu-rep --page raw --csv > gpurun_out/prof_r1c_all_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attn_tc_kernel" -s 6 -c 2 -o gpurun_out/prof_r1c_attn_tc -f python tools/prof_pool.py > gpurun_out/ncu_f35b.log 2>&1; echo "ncu attn exit=$?"
du -sh gpurun_out
