"""GPU experiment: sweep-kernel time vs pipeline knobs (stages, L2 promotion, splits)."""
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
base = None
for stages in (4, 5):
    for promo in (3, 2, 0):
        for splits in (8, 4, 16):
            eng.set_option("tc_stages", stages)
            eng.set_option("tc_l2promo", promo)
            eng.set_option("splits", splits)
            ts = []
            for rep in range(4):
                res = eng.run_loaded()
                ts.append(eng.last_timing()["sweep_ms"])
            if base is None:
                base = res.tobytes()
            ok = res.tobytes() == base
            t = min(ts[1:])
            print(f"stages {stages} l2promo {promo} splits {splits:2d}: sweep {t:7.3f} ms  {ng * N * M / t / 1e6:7.1f} GB/s  same={ok}", flush=True)
