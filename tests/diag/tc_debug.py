"""GPU diagnostic: run the same genes through the dp4a and the tcgen05 sweeps and diff the raw
integer partials (they must be bit identical).  python tools/tc_debug.py [N] [M] [C]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import rvtests_b200  # noqa: E402
from oracle import oracle as O  # noqa: E402
from util import af_of, make_problem  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
M = int(sys.argv[2]) if len(sys.argv) > 2 else 50
Cc = int(sys.argv[3]) if len(sys.argv) > 3 else 3
G, X, y = make_problem(O, 5, N, M, Cc, maf=np.linspace(0.05, 0.45, M), n_flip=2)
eng = rvtests_b200.GeneEngine(0)
eng.set_null_model(X, y)
print("tc_available", eng.info("tc_available"), "ER", eng.info("ER"))
outs = {}
for which in (1, 2):
    eng.set_option("engine", which)
    eng.push_i8(G.T.copy(), af_of(G))
    eng.push_i8(G.T.copy()[::-1].copy(), af_of(G)[::-1].copy())
    res = eng.flush()
    outs[which] = (res, eng.debug_partials(), eng.last_timing())
    print("engine", which, "splits", eng.info("last_splits"), "Q", res["Q"], "p", res["p_skat"], "status", res["status"], eng.last_timing())
a, b = outs[1][1], outs[2][1]
if len(b) == 2 * len(a):   # wide tensor-core sweep: two partials (even / odd boxes) per unit
    bb = np.zeros(len(a), dtype=b.dtype)
    bb["d"] = b["d"][0::2] + b["d"][1::2]
    bb["coll"] = b["coll"][0::2] + b["coll"][1::2]
    b = bb
ER = int(eng.info("ER"))
NC = 64 + ER
ok = True
for u in range(len(a)):
    Mrows = M
    da, db = a["d"][u][:Mrows, :NC], b["d"][u][:Mrows, :NC]
    # gene x gene block only where both < M ; digit columns all
    cols = list(range(M)) + list(range(64, NC))
    diff = (da[:, cols] != db[:, cols])
    if diff.any():
        ok = False
        rr, cc = np.nonzero(diff)
        print(f"unit {u}: {diff.sum()} mismatching d entries; first rows {sorted(set(rr.tolist()))[:10]} cols {sorted(set(np.array(cols)[cc].tolist()))[:20]}")
        i, j = rr[0], np.array(cols)[cc[0]]
        print("  e.g. d[%d][%d] simt=%d tc=%d" % (i, j, da[i, j], db[i, j]))
        print("  simt row0[:8]", da[0, :8], " tc row0[:8]", db[0, :8])
        print("  simt row1[:8]", da[1, :8], " tc row1[:8]", db[1, :8])
        print("  simt row17[:8]", da[min(17, Mrows-1), :8], " tc", db[min(17, Mrows-1), :8])
        print("  simt digits row0", da[0, 64:NC], "\n  tc   digits row0", db[0, 64:NC])
    ca, cb = a["coll"][u][: 2 * (ER + 1)], b["coll"][u][: 2 * (ER + 1)]
    if (ca != cb).any():
        ok = False
        print(f"unit {u}: collapse sums differ\n  simt {ca}\n  tc   {cb}")
print("PARTIALS IDENTICAL" if ok else "PARTIALS DIFFER")
print("results identical:", outs[1][0].tobytes() == outs[2][0].tobytes())
