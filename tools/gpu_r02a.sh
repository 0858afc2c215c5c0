#!/bin/bash
# r02a: the tests that never ran on hardware first (per-test timeouts), then the full suite, a short bench, SKAT-O timing
mkdir -p gpurun_out
echo "== zz tests"; timeout 900 python -m pytest tests/test_gpu_zz_fp64_skato.py -m gpu -q -x --durations=10 > gpurun_out/r02a_zz.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/r02a_zz.log
echo "== full suite"; timeout 1200 python -m pytest tests -m gpu -q --durations=15 > gpurun_out/r02a_pytest.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/r02a_pytest.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > gpurun_out/r02a_smoke.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r02a_smoke.log
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; echo "rc=$?"; cat gpurun_out/r02a_bench.json; tail -3 gpurun_out/r02a_bench.err
echo "== skato"; timeout 600 python tools/skato_time.py > gpurun_out/r02a_skato.log 2>&1; echo "rc=$?"; cat gpurun_out/r02a_skato.log
