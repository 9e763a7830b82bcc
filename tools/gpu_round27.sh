#!/bin/bash
mkdir -p gpurun_out
for pdl in 0 1; do
GRIDMM_PDL=$pdl timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench27_$pdl.json 2> gpurun_out/bench27_$pdl.err; echo "bench pdl=$pdl exit=$?"; python - <<PY
import json
d=json.loads(open('gpurun_out/bench27_$pdl.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'])
PY
done
