"""debug helper: run the loaded-cohort path with a tc_debug_skip mode (for compute-sanitizer)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rvtests_b200
from rvtests_b200 import synth
N, M, ng = int(sys.argv[2]) if len(sys.argv) > 2 else 70000, 50, 40
skip = int(sys.argv[1]) if len(sys.argv) > 1 else 1
wide = int(sys.argv[3]) if len(sys.argv) > 3 else 1
keys, t0, t1 = synth.variant_params(20260925, 0, ng * M)
X, y = synth.covariates(20260925, N, 3)
eng = rvtests_b200.GeneEngine(0)
eng.set_null_model(X, y)
eng.synth_load(keys, t0, t1, ng, M)
eng.set_option("tc_wide", wide)
eng.set_option("tc_debug_skip", skip)
for rep in range(2):
    r = eng.run_loaded()
print("ok", skip, eng.last_timing())
