#!/bin/bash
# round-2 evidence (same files as tools/gpu_profile.sh + the SKAT-O arm, the meta workload, the packed quadrature capture);
# tools/ncu_summary.py r02 turns gpurun_out/ into the tracked files under profiles/
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -14 gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/smoke.log
echo "== full bench"; timeout 1500 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "rc=$?"; cut -c1-600 gpurun_out/bench_full.json
echo "== bench configs[1] (N=100k x 30, 2000 genes)"; timeout 600 python bench.py --samples 100000 --variants 30 --genes 2000 --no-cpu > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "rc=$?"; cut -c1-300 gpurun_out/bench_c2.json
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$?"; cut -c1-400 gpurun_out/bench_ref.json
echo "== ncu launch list (same command, fewer steps)"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-skato > gpurun_out/ncu_bench.log 2>&1; echo "rc=$?"; grep -c . gpurun_out/launches.csv
echo "== ncu full set on sweep + finalize"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_sweep_tc|k_finalize' -s 2 -c 2 -o gpurun_out/prof -f python bench.py --genes 512 --steps 1 --warmup 3 --no-cpu --no-e2e --no-skato > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_full.log | cut -c1-200; ls -la gpurun_out/prof.ncu-rep
echo "== phases"; timeout 300 python tools/phases.py > gpurun_out/phases.log 2>&1; tail -9 gpurun_out/phases.log
echo "== skato timing"; timeout 300 python tools/overlap_time.py 2500 quick > gpurun_out/skato_time.log 2>&1; cat gpurun_out/skato_time.log
echo "== bench meta"; timeout 900 python bench.py --workload meta > gpurun_out/bench_meta.json 2> gpurun_out/bench_meta.err; echo "rc=$?"; cut -c1-400 gpurun_out/bench_meta.json
echo "== bench meta reference arm"; timeout 600 python bench.py --workload meta --impl reference --steps 2 --warmup 1 > gpurun_out/bench_meta_ref.json 2> gpurun_out/bench_meta_ref.err; echo "rc=$?"
echo "== ncu pair sweep (meta)"; timeout 900 ncu --set full --clock-control none -k regex:'k_sweep_tc' -s 4 -c 1 -o gpurun_out/prof_meta -f python tools/meta_time.py > gpurun_out/ncu_meta.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/ncu_meta.log | cut -c1-200
