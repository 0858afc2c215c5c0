#!/bin/bash
mkdir -p gpurun_out
echo "== parity tests"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_golden.py tests/test_gpu_aug.py -m gpu -q > gpurun_out/r02n_tests.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/r02n_tests.log
echo "== step timing"; timeout 300 python tools/overlap_time.py 2500 quick 2>&1 | grep "skato"
echo "== phases"; timeout 300 python tools/phases.py 2>&1 | tail -9
