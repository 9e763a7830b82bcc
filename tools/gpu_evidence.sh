#!/bin/bash
# final round-1 evidence: GPU tests, smoke, bench (with CPU baseline), reference arm, ncu launch list, ncu --set full of every kernel class
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 150 --timeout-method=thread -p no:cacheprovider > gpurun_out/test_all29.log 2>&1; echo "all gpu tests exit=$?"; tail -2 gpurun_out/test_all29.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke29.log 2>&1; echo "smoke exit=$?"; tail -1 gpurun_out/smoke29.log
timeout 600 python bench.py --steps 200 --warmup 5 > gpurun_out/bench29.json 2> gpurun_out/bench29.err; echo "bench exit=$?"; cut -c1-2500 gpurun_out/bench29.json; tail -3 gpurun_out/bench29.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench29_ref.json 2> gpurun_out/bench29_ref.err; echo "ref exit=$?"; cut -c1-600 gpurun_out/bench29_ref.json
timeout 600 python tools/microbench2.py > gpurun_out/microbench29.log 2>&1; echo "micro exit=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1c.csv python tools/prof_pool.py > gpurun_out/ncu_l29.log 2>&1; echo "ncu launches exit=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"pool_kernel" -s 2 -c 1 -o gpurun_out/prof_r1c_pool -f python tools/prof_pool.py > gpurun_out/ncu_f29a.log 2>&1; echo "ncu pool exit=$?"
timeout 900 ncu --set full --clock-control none --kernel-name-base demangled -k regex:"gmm::" -s 125 -c 59 -o /tmp/prof_r1c_all -f python tools/prof_pool.py > gpurun_out/ncu_f29b.log 2>&1; echo "ncu all exit=$?"
ncu -i /tmp/prof_r1c_all.ncu-rep --page raw --csv > gpurun_out/prof_r1c_all_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attn_tc_kernel" -s 6 -c 2 -o gpurun_out/prof_r1c_attn_tc -f python tools/prof_pool.py > gpurun_out/ncu_f29c.log 2>&1; echo "ncu attn exit=$?"
du -sh gpurun_out
