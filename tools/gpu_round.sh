#!/bin/bash
# One GPU session: TC diagnostic, GPU tests, benches.  Everything under timeout; logs to gpurun_out/.
mkdir -p gpurun_out
echo "== tc_debug small"; timeout 90 python tests/diag/tc_debug.py 2048 50 3 > gpurun_out/tc_debug_small.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/tc_debug_small.log
echo "== tc_debug big"; timeout 90 python tests/diag/tc_debug.py 70000 50 3 > gpurun_out/tc_debug_big.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/tc_debug_big.log
echo "== pytest gpu"; timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_gpu.log
echo "== bench simt (small)"; timeout 600 python bench.py --engine 1 --genes 256 --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_simt.json 2> gpurun_out/bench_simt.err; echo "rc=$?"; cat gpurun_out/bench_simt.json; tail -3 gpurun_out/bench_simt.err
echo "== bench tc (small)"; timeout 600 python bench.py --engine 2 --genes 256 --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err; echo "rc=$?"; cat gpurun_out/bench_tc.json; tail -3 gpurun_out/bench_tc.err
