#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_gpu.log
echo "== skato timing"; timeout 600 python tools/skato_time.py > gpurun_out/skato_time.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/skato_time.log
