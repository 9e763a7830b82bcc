#!/bin/bash
# end-of-round evidence on the final code: GPU tests, smoke, bench (both arms), ncu launch list, microbenchmarks
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 150 --timeout-method=thread -p no:cacheprovider > gpurun_out/test_final.log 2>&1; echo "all gpu tests exit=$?"; tail -2 gpurun_out/test_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; echo "smoke exit=$?"; tail -1 gpurun_out/smoke_final.log
timeout 600 python bench.py --steps 200 --warmup 5 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench exit=$?"; cut -c1-2500 gpurun_out/bench_final.json; tail -3 gpurun_out/bench_final.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err; echo "ref exit=$?"; cut -c1-600 gpurun_out/bench_final_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1d.csv python tools/prof_pool.py > gpurun_out/ncu_lfinal.log 2>&1; echo "ncu launches exit=$?"
timeout 400 python tools/microbench2.py > gpurun_out/microbench_final.log 2>&1; echo "micro exit=$?"
du -sh gpurun_out
