#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench33.json 2> gpurun_out/bench33.err; echo "bench exit=$?"; python - <<PY
import json
d=json.loads(open('gpurun_out/bench33.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, {k:v for k,v in d['e2e'].items() if k!='mode'})
PY
tail -5 gpurun_out/bench33.err
