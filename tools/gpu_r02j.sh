#!/bin/bash
mkdir -p gpurun_out
echo "== ncu launch list (headline step only)"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-skato > gpurun_out/ncu_bench.log 2>&1; echo "rc=$?"; grep -c . gpurun_out/launches.csv
echo "== ncu pair sweep (meta)"; timeout 900 ncu --set full --clock-control none -k regex:'k_sweep_tc' -s 4 -c 1 -o gpurun_out/prof_meta -f python tools/meta_time.py > gpurun_out/ncu_meta.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/ncu_meta.log | cut -c1-200
