"""GPU: cost of the permutation test (csrc/perm.cuh) at the benchmark shape (N = 500 000 x M = 50)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rvtests_b200  # noqa: E402
from rvtests_b200 import synth  # noqa: E402

N, M, ng = int(os.environ.get("PERM_N", 500_000)), 50, 4
keys, t0, t1 = synth.variant_params(20260925, 0, ng * M)
X, y = synth.covariates(20260925, N, 3)
eng = rvtests_b200.GeneEngine(0)
eng.set_null_model(X, y)
eng.synth_load(keys, t0, t1, ng, M)
base = eng.run_loaded()
CONFIGS = ((256, 256),) if os.environ.get("PERM_QUICK") else ((256, 1024), (1024, 2048))
for batch, nperm in CONFIGS:
    print("config", batch, nperm, flush=True)
    eng.set_option("perm", nperm)
    eng.set_option("perm_alpha", 1.0)       # no early stop: every permutation runs
    eng.set_option("perm_batch", batch)
    eng.set_option("perm_seed", 1)
    t = time.perf_counter()
    res = eng.run_loaded()
    dt = time.perf_counter() - t
    pr = eng.perm_results()
    tot = int(pr["actual_perm"].sum())
    print(f"batch {batch}: {tot} permutations of N={N} x M={M} over {ng} genes in {dt:.3f} s -> {tot / dt:.0f} perm/s "
          f"({tot * (N - 1) / dt / 1e9:.2f} G rand()/s); p_perm {pr['p_perm'].round(4).tolist()} vs analytic {res['p_skat'].round(4).tolist()}")
eng.set_option("perm", 0)
# the reference's own loop on the host (oracle) is timed beside it by tests/diag/perm_vs_reference_loop.py
