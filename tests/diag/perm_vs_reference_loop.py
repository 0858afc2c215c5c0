"""Diagnostic (not collected by pytest): the reference's permutation loop -- oracle restatement of src/Model.h:2707-2717 with
glibc rand(), one host thread -- timed on one gene of the benchmark shape beside the device replay of the same shuffles, and
the per-shuffle statistics compared.  Lives under tests/ because it loads the oracle."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import rvtests_b200  # noqa: E402
from rvtests_b200 import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

N, M, nref = int(os.environ.get("PERM_N", 500_000)), 50, 8
keys, t0, t1 = synth.variant_params(20260925, 0, M)
X, y = synth.covariates(20260925, N, 3)
eng = rvtests_b200.GeneEngine(0)
eng.set_null_model(X, y)
eng.synth_load(keys, t0, t1, 1, M)
base = eng.run_loaded()
O.build()
G0 = eng.loaded_read(0, M).T.astype(np.float64)
af = 0.5 * G0.sum(axis=0) / N
nm = O.fit_null_linear(X, y)
t = time.perf_counter()
ref = O.gene_perm(G0, af, nm["resid"], float(base[0]["Q"]), n_perm=nref, alpha=1.0, reseed=1)
dt = time.perf_counter() - t
print(f"reference loop (1 host thread): {nref} permutations in {dt:.3f} s -> {nref / dt:.1f} perm/s")
eng.set_option("perm", nref)
eng.set_option("perm_alpha", 1.0)
eng.set_option("perm_seed", 1)
eng.set_option("debug_perm_q", 1)
eng.run_loaded()
q = eng.perm_debug_q()[:nref]
print("max rel diff of the permuted statistics vs the reference loop (float32 there):", float(np.max(np.abs(q - ref["q"]) / ref["q"])))
