#!/bin/bash
mkdir -p gpurun_out
echo "== bench 2 GPUs"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02q_bench_n2.json 2> gpurun_out/r02q_bench_n2.err; echo "rc=$?"; tail -3 gpurun_out/r02q_bench_n2.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02q_bench_n2.json') if l.startswith('{')][-1])
print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['skato']['value'], d['engine']['numa'])
PY
echo "== bench 2 GPUs reference arm"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/r02q_ref_n2.json 2> gpurun_out/r02q_ref_n2.err; echo "rc=$?"; cut -c1-200 gpurun_out/r02q_ref_n2.json
echo "== bench meta 2 GPUs"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --workload meta --gpus 2 --steps 3 --warmup 3 --no-cpu > gpurun_out/r02q_meta_n2.json 2> gpurun_out/r02q_meta_n2.err; echo "rc=$?"; tail -3 gpurun_out/r02q_meta_n2.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02q_meta_n2.json') if l.startswith('{')][-1])
print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'])
PY
