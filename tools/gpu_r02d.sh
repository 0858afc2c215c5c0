#!/bin/bash
mkdir -p gpurun_out
echo "== dropin + adapters tests"; timeout 900 python -m pytest tests/test_gpu_dropin.py tests/test_gpu_adapters.py tests/test_gpu_zz_fp64_skato.py -m gpu -q -x > gpurun_out/r02d_dropin.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/r02d_dropin.log
echo "== skato timing q12 (2500 genes)"; timeout 600 python tools/overlap_time.py 2500 quick > gpurun_out/r02d_skato_q12.log 2>&1; echo "rc=$?"; cat gpurun_out/r02d_skato_q12.log
echo "== skato timing q16"; RVT_B200_LIB_VARIANT=$PWD/rvtests_b200/librvtests_b200_q16.so timeout 600 python tools/overlap_time.py 2500 quick > gpurun_out/r02d_skato_q16.log 2>&1; echo "rc=$?"; cat gpurun_out/r02d_skato_q16.log
echo "== full suite"; timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02d_pytest.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/r02d_pytest.log
