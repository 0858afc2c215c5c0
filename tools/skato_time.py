"""GPU: cost of the SKAT-O tail at the benchmark shape."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rvtests_b200  # noqa: E402
from rvtests_b200 import synth  # noqa: E402

N, M, ng = 500_000, 50, 600
keys, t0, t1 = synth.variant_params(20260925, 0, ng * M)
X, y = synth.covariates(20260925, N, 3)
eng = rvtests_b200.GeneEngine(0)
eng.set_null_model(X, y)
eng.synth_load(keys, t0, t1, ng, M)
eng.set_option("debug_phases", 1)
for sk in (0, 1):
    eng.set_option("skato", sk)
    for rep in range(2):
        res = eng.run_loaded()
    print("skato", sk, eng.last_timing())
    ph = eng.debug_phases(ng)
    print("  phase medians (cycles):", np.median(ph, axis=0).astype(int).tolist())
print("skato_ok frac", res["skato_ok"].mean(), "median p", np.median(res["skato_p"]), "rho hist", np.unique(res["skato_rho"], return_counts=True))
