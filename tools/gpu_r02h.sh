#!/bin/bash
mkdir -p gpurun_out
echo "== full suite"; timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02h_pytest.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/r02h_pytest.log
echo "== skato (single quadrature launch)"; timeout 300 python tools/overlap_time.py 2500 quick 2>&1 | grep "skato"
echo "== skato 600 genes"; timeout 300 python tools/skato_time.py 2>&1 | tail -6
