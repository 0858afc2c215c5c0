#!/bin/bash
mkdir -p gpurun_out
echo "== aug tests"; timeout 600 python -m pytest tests/test_gpu_aug.py -m gpu -q > gpurun_out/r02k_aug.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/r02k_aug.log
echo "== imputed timing"; timeout 600 python tools/imputed_time.py > gpurun_out/r02k_imputed.log 2>&1; echo "rc=$?"; cat gpurun_out/r02k_imputed.log
