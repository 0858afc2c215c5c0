#!/bin/bash
mkdir -p gpurun_out
echo "== perm tests"; timeout 900 python -m pytest tests/test_gpu_perm.py -m gpu -q > gpurun_out/r02p_perm.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/r02p_perm.log
