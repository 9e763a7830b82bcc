#!/bin/bash
# ncu --set full over every launch of one navigation step (third step of tools/prof_pool.py: 7 grid builds + 2 x 58 launches precede it)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --kernel-name-base demangled -k regex:"gmm::" -s 123 -c 58 -o /tmp/prof_r1d_all -f python tools/prof_pool.py > gpurun_out/ncu_step_final.log 2>&1; echo "ncu all exit=$?"
ncu -i /tmp/prof_r1d_all.ncu-rep --page raw --csv > gpurun_out/prof_r1d_all_raw.csv 2>/dev/null
du -sh gpurun_out; wc -l gpurun_out/prof_r1d_all_raw.csv
