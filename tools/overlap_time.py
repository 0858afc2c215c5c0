"""GPU: ring depth of the sweep x statistics-under-sweep overlap, at the benchmark shape (ms per 2 500-gene step)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rvtests_b200  # noqa: E402
from rvtests_b200 import synth  # noqa: E402

N, M, ng = 500_000, 50, int(sys.argv[1]) if len(sys.argv) > 1 else 2500
keys, t0, t1 = synth.variant_params(20260925, 0, ng * M)
X, y = synth.covariates(20260925, N, 3)
eng = rvtests_b200.GeneEngine(0)
eng.set_null_model(X, y)
eng.synth_load(keys, t0, t1, ng, M)
eng.set_option("qags_pack", int(os.environ.get("RVT_QAGS_PACK", "1")))
base = None
quick = len(sys.argv) > 2 and sys.argv[2] == "quick"
for skato in (0, 1):
    eng.set_option("skato", skato)
    for stages in ((5,) if quick else (5, 4, 3)):
        for ovl in ((0,) if quick else (0, 2, 4, 8, 16)):
            eng.set_option("tc_stages", stages)
            eng.set_option("overlap", ovl)
            for _ in range(2):
                res = eng.run_loaded()
            ts = []
            for _ in range(4):
                res = eng.run_loaded()
                ts.append(eng.last_timing())
            tot = np.mean([t["total_ms"] for t in ts])
            print(f"skato {skato} stages {stages} overlap {ovl:2d}: total {tot:7.3f} ms  sweep {np.mean([t['sweep_ms'] for t in ts]):7.3f}  "
                  f"stat {np.mean([t['finalize_ms'] for t in ts]):7.3f}  -> {ng / tot * 1e3:9.0f} genes/s", flush=True)
            if base is None:
                base = res.copy()
            elif skato == 0:
                assert res.tobytes() == base.tobytes(), "records differ from the plain run"
eng.close()
