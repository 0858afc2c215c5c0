"""GPU: throughput of genes that arrive as PLINK 2-bit rows WITH missing calls (the reference's default --impute mean path,
src/DataConsolidator.cpp:217-245) and of binary-trait genes -- both take the engine's fp64 path -- beside complete
hard-call genes through the integer sweep, end to end from pinned host rows at the benchmark shape."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rvtests_b200  # noqa: E402
from rvtests_b200 import synth  # noqa: E402

N, M, ng = int(os.environ.get("IMP_N", 500_000)), 50, int(os.environ.get("IMP_GENES", 128))
X, y = synth.covariates(20260925, N, 3)
rng = np.random.default_rng(11)
maf = 10 ** rng.uniform(-4, np.log10(0.05), (ng, M))
stride = (N + 3) // 4
beds = {}
for name, miss in ((("complete", 0.0),) if os.environ.get("IMP_ONLY") == "binary" else (("complete", 0.0), ("1% missing", 0.01))):
    buf = torch.empty((ng, M, stride), dtype=torch.uint8).pin_memory().numpy()
    for g in range(ng):
        u = rng.integers(0, 65536, size=(M, N), dtype=np.uint16)
        m = maf[g][:, None]
        G = (u < (65536 * (1 - (1 - m) ** 2))).astype(np.uint8) + (u < (65536 * m * m)).astype(np.uint8)
        code = np.where(G == 0, 0, np.where(G == 1, 2, 3)).astype(np.uint8)
        if miss > 0:
            code[rng.random((M, N)) < miss] = 1
        c4 = np.pad(code, ((0, 0), (0, (-N) % 4))).reshape(M, -1, 4)
        buf[g] = c4[:, :, 0] | (c4[:, :, 1] << 2) | (c4[:, :, 2] << 4) | (c4[:, :, 3] << 6)
    beds[name] = buf
eng = rvtests_b200.GeneEngine(0)


def run(tag, bed, binary, aug=1):
    eng.set_option("aug", aug)
    if os.environ.get("IMP_STREAM"):
        eng.set_option("stream_batch", int(os.environ["IMP_STREAM"]))     # sweep / statistics enqueued under the following copies, as bench.py's e2e leg does
    if os.environ.get("IMP_BINSTREAM"):
        eng.set_option("binary_stream", int(os.environ["IMP_BINSTREAM"]))
    if binary:
        yb = (np.random.default_rng(1).random(N) < 0.3).astype(np.float64)
        eng.set_null_model(X, yb, binary=True)
    else:
        eng.set_null_model(X, y)
    for rep in range(2):
        t = time.perf_counter()
        for g in range(ng):
            eng.push_bed(bed[g], None)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        res = eng.flush()
        dt = time.perf_counter() - t
        dflush = time.perf_counter() - t1
    print(f"{tag:52s}: {ng / dt:8.0f} genes/s end to end ({dt * 1e3:.1f} ms for {ng} genes); flush alone (tiles resident) "
          f"{dflush * 1e3:.2f} ms = {ng / dflush:8.0f} genes/s; augmented genes {int(eng.info('last_aug'))}; status ok {int((res['status'] == 0).sum())}/{ng}", flush=True)


if os.environ.get("IMP_ONLY") != "binary":
    run("complete hard calls (integer sweep)", beds["complete"], False)
    run("1% missing calls -> augmented tensor-core sweep", beds["1% missing"], False)
    run("1% missing calls -> sparse CUDA-core kernel (r01)", beds["1% missing"], False, aug=0)
run("binary trait, complete calls (fp64 path)", beds["complete"], True)
