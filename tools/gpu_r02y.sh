#!/bin/bash
# r02y: BASELINE configs[4] on 8 B200 -- BoltLMM null fit at N = 1M x 131 072 panel SNPs (33 GB of 2-bit rows, 16 384 rows per GPU),
# one ncclAllReduce per H-product, then the score test of 8 192 variants per GPU on the fitted null
mkdir -p gpurun_out
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --workload bolt --gpus $N --bolt-snps $((16384 * N)) --steps 2 --warmup 1 > gpurun_out/bolt_n$N.json 2> gpurun_out/bolt_n$N.err; echo "bolt n$N rc=$?"; tail -2 gpurun_out/bolt_n$N.err | cut -c1-300
python - <<PY
import json
d = json.loads(open("gpurun_out/bolt_n$N.json").read().strip().splitlines()[-1])
print(d["n_gpus"], d["value"], d["ms_per_step"], d["fit"]["h_products"], d["kernel_ms_per_step"], d["roofline"]["frac"], d["roofline"]["share_of_step"], (d.get("e2e") or {}).get("value"), (d.get("score_test") or {}).get("value"), d["engine"]["allreduce_calls_per_fit"])
PY
