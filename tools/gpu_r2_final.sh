#!/bin/bash
# round-2 evidence on the current code: GPU tests, smoke, bench (both arms + every workload), host time, microbenchmarks, pooling
# kernel probe (per-CTA timeline, stage trace), ncu launch list, ncu --set full of every launch of one step, sanitizer runs
mkdir -p gpurun_out
tag=${1:-r2}
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider -s > gpurun_out/${tag}_tests.log 2>&1; echo "gpu tests exit=$?"
grep -E "passed|failed" gpurun_out/${tag}_tests.log | tail -2; grep -E "^(FAILED|ERROR)|^E  " gpurun_out/${tag}_tests.log | head -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke exit=$?"; tail -1 gpurun_out/${tag}_smoke.log
timeout 900 python bench.py --steps 200 --warmup 5 --gpu-eager-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench exit=$?"; cut -c1-1200 gpurun_out/${tag}_bench.json; tail -3 gpurun_out/${tag}_bench.err
GRIDMM_PDL=0 timeout 600 python bench.py --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_bench_nopdl.json 2> gpurun_out/${tag}_bench_nopdl.err; echo "bench (PDL off) exit=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; echo "ref exit=$?"; cut -c1-500 gpurun_out/${tag}_bench_ref.json
for w in "reverie 8" "ce 8" "r2r 1" "r2r 15"; do set -- $w; timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --workload $1 --T $2 > gpurun_out/${tag}_bench_$1_T$2.json 2> gpurun_out/${tag}_bench_$1_T$2.err; echo "bench $1 T=$2 exit=$?"; done
timeout 600 python bench.py --workload pretrain --steps 20 --warmup 4 > gpurun_out/${tag}_bench_pretrain.json 2> gpurun_out/${tag}_bench_pretrain.err; echo "bench pretrain exit=$?"; cut -c1-400 gpurun_out/${tag}_bench_pretrain.json
timeout 300 python tools/host_time.py > gpurun_out/${tag}_host_time.txt 2>&1; head -4 gpurun_out/${tag}_host_time.txt
timeout 600 python tools/microbench2.py > gpurun_out/${tag}_microbench.txt 2>&1; echo "micro exit=$?"
timeout 300 python tools/pool_probe.py ${tag} > gpurun_out/${tag}_pool_probe.txt 2>&1; echo "pool probe exit=$?"; python tools/pool_trace_summary.py ${tag} >> gpurun_out/${tag}_pool_probe.txt 2>&1
for e in 2 4; do POOL_EXP=$e timeout 300 python tools/pool_probe.py ${tag}_e$e 2>&1 | grep "^experiment\|^events, 8\|^valid rows" > gpurun_out/${tag}_pool_exp$e.txt; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/${tag}_launches.csv python tools/prof_pool.py > gpurun_out/${tag}_ncu_l.log 2>&1; echo "ncu launches exit=$?"
# one full step = the launches between the last two grid_update kernels; count them from the launch list
python - <<PY
import csv
rows=[r for r in csv.DictReader([l for l in open('gpurun_out/${tag}_launches.csv') if not l.startswith('==')]) if r.get('Metric Name')=='gpu__time_duration.sum']
gm=[i for i,r in enumerate(rows) if 'gmm::' in r['Kernel Name']]
names=[rows[i]['Kernel Name'] for i in gm]
idx=[i for i,n in enumerate(names) if 'grid_update' in n]
print('gmm launches', len(names), 'grid_update at', idx[-3:])
open('gpurun_out/${tag}_skip.txt','w').write('%d %d' % (idx[-1], len(names)-idx[-1]))
PY
read skip cnt < gpurun_out/${tag}_skip.txt
timeout 1200 ncu --set full --clock-control none --kernel-name-base demangled -k regex:"gmm::" -s $skip -c $cnt -o /tmp/prof_${tag}_all -f python tools/prof_pool.py > gpurun_out/${tag}_ncu_f.log 2>&1; echo "ncu all exit=$?"
ncu -i /tmp/prof_${tag}_all.ncu-rep --page raw --csv > gpurun_out/${tag}_all_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"pool_kernel" -s 2 -c 1 -o gpurun_out/${tag}_pool -f python tools/prof_pool.py > gpurun_out/${tag}_ncu_p.log 2>&1; echo "ncu pool exit=$?"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_memcheck.log 2>&1; echo "memcheck exit=$?"; tail -3 gpurun_out/${tag}_memcheck.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "pool_handmade and 40_ctas" -p no:cacheprovider > gpurun_out/${tag}_memcheck_pool.log 2>&1; echo "memcheck pool pieces exit=$?"; tail -3 gpurun_out/${tag}_memcheck_pool.log
du -sh gpurun_out
