#!/bin/bash
# r02r: FastLMM covariance band test + the full default bench line (with the missing-call e2e leg)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_lmm.py -x -q -m gpu > gpurun_out/r02r_lmm.log 2>&1; echo "lmm rc=$?"; tail -5 gpurun_out/r02r_lmm.log
timeout 600 python bench.py > gpurun_out/r02r_bench.json 2> gpurun_out/r02r_bench.err; echo "bench rc=$?"; cat gpurun_out/r02r_bench.json; tail -5 gpurun_out/r02r_bench.err
