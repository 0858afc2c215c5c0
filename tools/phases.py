"""GPU diagnostic: where does the per-gene finalize kernel spend its cycles?"""
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
for rep in range(2):
    res = eng.run_loaded()
ph = eng.debug_phases(ng)
print("timing", eng.last_timing())
names = ["reduce", "K build", "eigen", "davies", "liu+burden", "-"]
tot = ph[:, :5].sum(axis=1)
print("per-gene total cycles: median %.0f  p90 %.0f  max %.0f" % (np.median(tot), np.percentile(tot, 90), tot.max()))
for k in range(5):
    print("  %-12s median %9.0f  p90 %9.0f  max %9.0f  share %.2f" % (names[k], np.median(ph[:, k]), np.percentile(ph[:, k], 90), ph[:, k].max(), ph[:, k].sum() / tot.sum()))
print("n_lambda median", np.median(res["n_lambda"]), "fault frac", (res["davies_fault"] != 0).mean())
