#!/bin/bash
# Bolt: GPU tests, bench line, ncu --set full of one launch of each panel-product kernel
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_bolt.py -m gpu -x -q --durations=5) > gpurun_out/pytest_bolt.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_bolt.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_bolt_xw2|k_bolt_xtv2' -s 8 -c 2 -o gpurun_out/prof_bolt -f python bench.py --workload bolt --steps 1 --warmup 1 --no-e2e > gpurun_out/ncu_bolt.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_bolt.log | cut -c1-200
ls -la gpurun_out/prof_bolt.ncu-rep
