#!/bin/bash
# r02b: second-generation SKAT-O tail on hardware; watchdog A/B of the sweep
mkdir -p gpurun_out
echo "== skato-touching tests"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_golden.py tests/test_gpu_zz_fp64_skato.py tests/test_gpu_adapters.py tests/test_gpu_binary.py -m gpu -q -x > gpurun_out/r02b_skato_tests.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/r02b_skato_tests.log
echo "== skato timing"; timeout 600 python tools/skato_time.py > gpurun_out/r02b_skato.log 2>&1; echo "rc=$?"; cat gpurun_out/r02b_skato.log
echo "== bench (watchdog build)"; timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu --no-e2e > gpurun_out/r02b_bench_wd.json 2> gpurun_out/r02b_bench_wd.err; echo "rc=$?"; python -c "import json;d=json.load(open('gpurun_out/r02b_bench_wd.json'));print(d['value'],d['kernel_ms_per_step'],d['roofline']['frac'])"
echo "== bench (no watchdog build)"; RVT_B200_LIB_VARIANT=$PWD/rvtests_b200/librvtests_b200_nowd.so timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu --no-e2e > gpurun_out/r02b_bench_nowd.json 2> gpurun_out/r02b_bench_nowd.err; echo "rc=$?"; python -c "import json;d=json.load(open('gpurun_out/r02b_bench_nowd.json'));print(d['value'],d['kernel_ms_per_step'],d['roofline']['frac'])"
echo "== bench (watchdog build, again)"; timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu --no-e2e > gpurun_out/r02b_bench_wd2.json 2> gpurun_out/r02b_bench_wd2.err; echo "rc=$?"; python -c "import json;d=json.load(open('gpurun_out/r02b_bench_wd2.json'));print(d['value'],d['kernel_ms_per_step'],d['roofline']['frac'])"
echo "== full suite"; timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02b_pytest.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/r02b_pytest.log
