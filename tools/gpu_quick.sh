#!/bin/bash
mkdir -p gpurun_out
echo "== tc_debug"; timeout 90 python tests/diag/tc_debug.py 70000 50 3 > gpurun_out/tc_debug_big.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/tc_debug_big.log
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/pytest_gpu.log
echo "== phases"; timeout 300 python tools/phases.py > gpurun_out/phases.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/phases.log
echo "== full bench"; timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "rc=$?"; cut -c1-2600 gpurun_out/bench_full.json; tail -3 gpurun_out/bench_full.err
