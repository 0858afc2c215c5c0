#!/bin/bash
mkdir -p gpurun_out
echo "== tc_debug"; timeout 90 python tests/diag/tc_debug.py 70000 50 3 > gpurun_out/tc_debug_big.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/tc_debug_big.log
echo "== pytest gpu"; timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/pytest_gpu.log
echo "== phases"; timeout 300 python tools/phases.py > gpurun_out/phases.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/phases.log
echo "== full bench (auto engine)"; timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "rc=$?"; cat gpurun_out/bench_full.json; tail -3 gpurun_out/bench_full.err
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$?"; cat gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err
echo "== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r01.csv python bench.py --genes 512 --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_bench.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_bench.log; grep -c . gpurun_out/launches_r01.csv
echo "== ncu full set on sweep + finalize"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_sweep_tc|k_finalize' -s 2 -c 2 -o gpurun_out/prof_r01 -f python bench.py --genes 512 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out/
