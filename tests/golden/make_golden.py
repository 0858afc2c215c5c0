"""Regenerates tests/golden/*.  Run in the BUILD container (needs /root/reference for oracle/_ref):
    python tests/golden/make_golden.py
Golden Davies/Liu vectors come from the REFERENCE's own qfc.c / MixtureChiSquare.cpp / cdflib.cpp
(compiled in place by oracle/Makefile into oracle/_ref/libmixchisq_ref.so)."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as O  # noqa: E402

O.build()
assert O.ref_mix() is not None, "oracle/_ref not built"

# (1) known answers of regression/test/testMixtureChiSquare.cpp:11-40 -- the PROGRAM's output
cases = []
for lam, Q in (([1.0, 2.0, 3.0], 4.0), ([1.0, 1.0, 1.0], 30.0), ([1.0, 1.0, 1.0], 50.0)):
    p, _ = O.mix_pvalue(lam, Q, "reference")
    cases.append(dict(**{"lambda": lam}, Q=Q, davies=p, liu=O.liu_pvalue(lam, Q, "reference")))
json.dump(dict(source="regression/test/testMixtureChiSquare.cpp:11-40 run against the reference build",
               cases=cases), open(os.path.join(HERE, "mixchisq_kat.json"), "w"), indent=1)

# (2) random eigenvalue spectra through the reference's qf
rng = np.random.default_rng(20260925)
T, R = 300, 64
lam = np.zeros((T, R)); n = np.zeros(T, dtype=np.int32); Q = np.zeros(T)
pd = np.zeros(T); pl = np.zeros(T); fault = np.zeros(T, dtype=np.int32)
for t in range(T):
    n[t] = rng.integers(1, R + 1)
    l = np.sort(rng.gamma(0.3, 1.0, n[t]) * 10 ** rng.uniform(-3, 3))[::-1]
    lam[t, : n[t]] = l
    Q[t] = l.sum() * 10 ** rng.uniform(-1.5, 1.3)
    if n[t] > 1:
        v, f, tr = O.qf(l.copy(), Q[t], which="reference")
        fault[t] = f
        pd[t] = -1.0 if f else min(1.0 - v, 1.0)
    else:
        pd[t] = O.liu_pvalue(l.copy(), Q[t], "reference")
    pl[t] = O.liu_pvalue(l.copy(), Q[t], "reference")
np.savez_compressed(os.path.join(HERE, "davies_golden.npz"), lam=lam, n=n, Q=Q, p_davies=pd, p_liu=pl, fault=fault)

# (3) C1 anchor (example/example.vcf, example/pheno y1, example/setFile): inputs transcribed from the
# reference's example files; outputs from the oracle + reference Davies
G = [[1, 0, 0], [0, 0, 0], [0, 0, 0], [2, 0, 0], [1, 0, 1], [0, 0, 0], [1, 1, 0], [1, 0, 0], [0, 1, 0]]
y = [1.911, 2.146, 1.086, 0.704, 2.512, 1.283, 2.384, 3.004, 0.714]
Ga = np.array(G, dtype=float)
nm = O.fit_null_linear(np.ones((9, 1)), np.array(y))
out, l = O.gene(Ga, 0.5 * Ga.sum(0) / 9, np.ones((9, 1)), nm["resid"], nm["sigma2"])
pref, _ = O.skat_final_pvalue(l, out.skat.Q, "reference")
json.dump(dict(G=G, y=y, sigma2=nm["sigma2"], Q=out.skat.Q, **{"lambda": l.tolist()}, fault=out.skat.fault,
               pvalue=pref, cmc_nonref=out.cmc_nonref, zeggini=[1, 0, 0, 1, 2, 0, 2, 1, 1]),
          open(os.path.join(HERE, "c1_anchor.json"), "w"), indent=1)
print("golden written")
