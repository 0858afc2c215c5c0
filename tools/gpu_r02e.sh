#!/bin/bash
mkdir -p gpurun_out
echo "== sanitizer on the failing case"; timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest "tests/test_gpu_zz_fp64_skato.py::test_binary_trait_skato_vs_oracle" -m gpu -q -x > gpurun_out/r02e_sanitizer.log 2>&1; echo "rc=$?"; grep -v "^$" gpurun_out/r02e_sanitizer.log | head -60
echo "== dropin"; timeout 600 python -m pytest tests/test_gpu_dropin.py -m gpu -q > gpurun_out/r02e_dropin.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r02e_dropin.log
for v in "" _ERRBD4 _TRUNC2 _INT2; do
  for pack in 0 1; do
    echo "== variant '$v' pack $pack"; RVT_QAGS_PACK=$pack RVT_B200_LIB_VARIANT=$PWD/rvtests_b200/librvtests_b200$v.so timeout 300 python tools/overlap_time.py 2500 quick 2>&1 | grep "skato 1"
  done
done
echo "== meta tests"; timeout 600 python -m pytest tests/test_gpu_meta.py tests/test_gpu_perm.py tests/test_gpu_lmm.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r02e_meta_tests.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r02e_meta_tests.log
echo "== meta timing"; timeout 900 python tools/meta_time.py > gpurun_out/r02e_meta_time.log 2>&1; echo "rc=$?"; cat gpurun_out/r02e_meta_time.log
