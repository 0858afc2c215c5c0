"""GPU diagnostics: one small permutation batch at a given N with per-stage synchronisation (RVT_PERM_TRACE=1)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rvtests_b200  # noqa: E402
from rvtests_b200 import synth  # noqa: E402

N, M, ng = int(sys.argv[1]), int(sys.argv[2]) if len(sys.argv) > 2 else 50, 1
keys, t0, t1 = synth.variant_params(20260925, 0, ng * M)
X, y = synth.covariates(20260925, N, 3)
eng = rvtests_b200.GeneEngine(0)
eng.set_null_model(X, y)
eng.synth_load(keys, t0, t1, ng, M)
B = int(sys.argv[3]) if len(sys.argv) > 3 else 16
eng.set_option("perm", B)
eng.set_option("perm_alpha", 1.0)
eng.set_option("perm_batch", B)
t = time.perf_counter()
eng.run_loaded()
print("ok", N, time.perf_counter() - t, eng.perm_results())
