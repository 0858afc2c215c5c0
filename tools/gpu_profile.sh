#!/bin/bash
# round-1 evidence: GPU tests, bench line, reference arm, ncu launch list of the same command, ncu --set full of the
# two hot kernels, finalize phase counters, sweep ablation.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_gpu.log
echo "== full bench"; timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "rc=$?"; cut -c1-1200 gpurun_out/bench_full.json
echo "== bench configs[1] (N=100k x 30, 2000 genes)"; timeout 300 python bench.py --samples 100000 --variants 30 --genes 2000 --no-cpu > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "rc=$?"; cut -c1-300 gpurun_out/bench_c2.json
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$?"; cut -c1-400 gpurun_out/bench_ref.json
echo "== ncu launch list (same command, fewer steps)"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_bench.log 2>&1; echo "rc=$?"; grep -c . gpurun_out/launches.csv
echo "== ncu full set on sweep + finalize"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_sweep_tc|k_finalize' -s 2 -c 2 -o gpurun_out/prof -f python bench.py --genes 512 --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_full.log | cut -c1-200; ls -la gpurun_out/prof.ncu-rep
echo "== phases"; timeout 300 python tools/phases.py > gpurun_out/phases.log 2>&1; tail -9 gpurun_out/phases.log
echo "== ablation"; timeout 300 python tools/sweep_knobs.py > gpurun_out/ablation.log 2>&1; tail -9 gpurun_out/ablation.log
