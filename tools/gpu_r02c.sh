#!/bin/bash
mkdir -p gpurun_out
echo "== overlap matrix"; timeout 900 python tools/overlap_time.py > gpurun_out/r02c_overlap.log 2>&1; echo "rc=$?"; cat gpurun_out/r02c_overlap.log
echo "== bench"; timeout 1500 python bench.py > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err; echo "rc=$?"; tail -5 gpurun_out/r02c_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02c_bench.json'))
for k in ('value','ms_per_step','kernel_ms_per_step','e2e','e2e_int8','e2e_f64','skato','parity','engine'):
    print(k, d.get(k))
print('cpu', {k:v for k,v in d.get('cpu_baseline',{}).items() if k!='reference_algorithm_wall'})
PY
