#!/bin/bash
# r02v: 4-GPU check of the default bench (NUMA fallback, e2e legs incl. the binary trait) and the sharded Bolt fit
mkdir -p gpurun_out
N=${1:-4}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n$N rc=$?"; tail -2 gpurun_out/bench_n$N.err | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --workload bolt --gpus $N --bolt-snps 65536 > gpurun_out/bolt_n$N.json 2> gpurun_out/bolt_n$N.err; echo "bolt n$N rc=$?"; tail -2 gpurun_out/bolt_n$N.err | cut -c1-300
python - <<PY
import json
for f in ("bench_n$N", "bolt_n$N"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], (d.get("e2e") or {}).get("value"), (d.get("e2e_binary_trait") or {}).get("value"), d.get("engine", {}).get("numa"), d.get("kernel_ms_per_step"))
    except Exception as e:
        print(f, "unreadable", e)
PY
