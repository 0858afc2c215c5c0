#!/bin/bash
mkdir -p gpurun_out
echo "== full suite"; timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02g_pytest.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/r02g_pytest.log
echo "== skato default (TRUNC2 + pack)"; timeout 300 python tools/overlap_time.py 2500 quick 2>&1 | grep "skato 1"
echo "== launch list (skato step)"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02g_launches_skato.csv python tools/overlap_time.py 2500 quick > /dev/null 2>&1; echo "rc=$?"; python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02g_launches_skato.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
agg={}
for r in rows[1:]:
    agg.setdefault(r[ki].split('(')[0],[]).append(float(r[vi].replace(',','')))
for k,v in agg.items(): print(k, len(v), 'launches, last', v[-1]/1e6, 'ms')
PY
echo "== ncu packed qags"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_skato_qags_packed --launch-skip 2 --launch-count 1 -o gpurun_out/r02g_qags python tools/overlap_time.py 2500 quick > gpurun_out/r02g_ncu.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r02g_ncu.log
echo "== bench meta"; timeout 900 python bench.py --workload meta > gpurun_out/r02g_bench_meta.json 2> gpurun_out/r02g_bench_meta.err; echo "rc=$?"; tail -3 gpurun_out/r02g_bench_meta.err; cat gpurun_out/r02g_bench_meta.json
echo "== bench meta reference"; timeout 900 python bench.py --workload meta --impl reference --steps 2 --warmup 1 > gpurun_out/r02g_bench_meta_ref.json 2> gpurun_out/r02g_bench_meta_ref.err; echo "rc=$?"; cat gpurun_out/r02g_bench_meta_ref.json
