#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
echo "== bench N=$N"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "rc=$?"; cut -c1-1500 gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err
echo "== reference arm under torchrun"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 1 --warmup 1 --cpu-seconds 4 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err; echo "rc=$?"; cut -c1-600 gpurun_out/bench_ref_n$N.json; tail -3 gpurun_out/bench_ref_n$N.err
