#!/bin/bash
mkdir -p gpurun_out
echo "== full suite (TRUNC2 + packed quadrature + parallel HWE + multi-device batcher)"; timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02f_pytest.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/r02f_pytest.log
echo "== skato default"; timeout 300 python tools/overlap_time.py 2500 quick 2>&1 | grep "skato 1"
echo "== skato RCP"; RVT_B200_LIB_VARIANT=$PWD/rvtests_b200/librvtests_b200_RCP.so timeout 300 python tools/overlap_time.py 2500 quick 2>&1 | grep "skato 1"
echo "== meta timing"; timeout 900 python tools/meta_time.py > gpurun_out/r02f_meta_time.log 2>&1; echo "rc=$?"; cat gpurun_out/r02f_meta_time.log
echo "== ncu packed qags"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_skato_qags_packed --launch-count 1 -o gpurun_out/r02f_qags python tools/overlap_time.py 1200 quick > gpurun_out/r02f_ncu.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r02f_ncu.log
