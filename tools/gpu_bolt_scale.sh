#!/bin/bash
# Bolt on 1 and 2 GPUs (NCCL all-reduce per H-product) + the reference arm; run under gpurun --gpus 2
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_bolt.py -m gpu -x -q) > gpurun_out/pytest_bolt.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_bolt.log
timeout 600 python bench.py --workload bolt > gpurun_out/bolt_n1.json 2> gpurun_out/bolt_n1.err; echo "n1 rc=$?"; tail -2 gpurun_out/bolt_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --workload bolt --gpus 2 > gpurun_out/bolt_n2.json 2> gpurun_out/bolt_n2.err; echo "n2 rc=$?"; tail -2 gpurun_out/bolt_n2.err
timeout 600 python bench.py --workload bolt --impl reference --steps 1 --warmup 0 > gpurun_out/bolt_ref.json 2> gpurun_out/bolt_ref.err; echo "ref rc=$?"; tail -2 gpurun_out/bolt_ref.err
python - <<'PY'
import json
for f in ("bolt_n1", "bolt_n2", "bolt_ref"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d.get("fit", {}).get("h_products"), d.get("kernel_ms_per_step"), (d.get("roofline") or {}).get("frac"), (d.get("e2e") or {}).get("value"), d.get("engine", {}).get("allreduce_calls_per_fit"))
    except Exception as e:
        print(f, "unreadable", e)
PY
