#!/bin/bash
mkdir -p gpurun_out
echo "== full suite"; timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02o_pytest.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/r02o_pytest.log
echo "== step timing (split statistics)"; timeout 300 python tools/overlap_time.py 2500 quick 2>&1 | grep "skato"
echo "== launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_fin|k_sweep_tc' -c 16 --csv --log-file gpurun_out/r02o_launches.csv python tools/overlap_time.py 2500 quick > /dev/null 2>&1; python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02o_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
for r in rows[1:]: print(r[ki].split('(')[0][:60], float(r[vi].replace(',',''))/1e6,'ms')
PY
echo "== configs[1]"; timeout 600 python bench.py --samples 100000 --variants 30 --genes 2000 --no-cpu --no-e2e --no-skato 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print(d['value'], d['kernel_ms_per_step'], d['roofline']['frac'])"
