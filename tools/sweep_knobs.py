"""GPU experiment: where does the tensor-core sweep spend its time?  Disables parts of the kernel
(results are garbage in those runs; only the sweep time is read)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rvtests_b200  # noqa: E402
from rvtests_b200 import synth  # noqa: E402

N, M, ng = 500_000, 50, 1200
keys, t0, t1 = synth.variant_params(20260925, 0, ng * M)
X, y = synth.covariates(20260925, N, 3)
eng = rvtests_b200.GeneEngine(0)
eng.set_null_model(X, y)
eng.synth_load(keys, t0, t1, ng, M)
names = {0: "full kernel", 1: "no collapse", 2: "no MMA", 4: "no E loads", 3: "no collapse, no MMA", 5: "no collapse, no E",
         6: "no MMA, no E", 7: "TMA gene stream only"}
for wide, zc, skip in ((0, 0, 0), (1, 0, 0), (0, 1, 0), (1, 1, 0), (0, 1, 2), (1, 1, 2), (0, 1, 4), (1, 1, 4)):
    eng.set_option("tc_wide", wide)
    eng.set_option("tc_zc", zc)
    eng.set_option("tc_debug_skip", skip)
    ts = []
    for rep in range(4):
        eng.run_loaded()
        ts.append(eng.last_timing()["sweep_ms"])
    t = min(ts[1:])
    print(f"wide {wide} zc {zc} skip {skip} ({names[skip]:24s}): sweep {t:7.3f} ms  {ng * N * M / t / 1e6:7.1f} GB/s", flush=True)
