"""Regenerates tests/golden/ref_assoc_golden.npz.  Run in the BUILD container (needs /root/reference):
    make -C oracle ref && python tests/golden/make_golden_ref_assoc.py
The text stored here is what the REFERENCE's own model layer prints -- src/Model.cpp + the fitters of src/Model.h +
src/DataConsolidator.cpp compiled unmodified into oracle/_ref/libmodel_ref.so and driven like the gene loop of
src/Main.cpp:1221-1254 (oracle/ref_model_shim.cpp) -- for the five genes of tests/test_gpu_adapters.py."""
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import oracle as O  # noqa: E402
from util import make_problem  # noqa: E402

O.build()
assert O.ref_model() is not None, "oracle/_ref/libmodel_ref.so not built"
N, C = 1200, 3
genes = []
for gi, (M, nm_, nf) in enumerate([(6, 0, 1), (1, 0, 0), (25, 2, 2), (4, 4, 0), (40, 1, 0)]):
    G, X, y = make_problem(O, 77, N, M, C, maf=np.linspace(0.002, 0.03, M), n_mono=nm_, n_flip=nf)
    rng = np.random.default_rng(gi)
    genes.append(G[:, rng.permutation(M)])
with tempfile.TemporaryDirectory() as d:
    out = O.ref_run_gene_models([g.astype(float) for g in genes], X[:, 1:], y, os.path.join(d, "g"))
text = {m: dict(header=out[m][1], rows=out[m][2]) for m in out}
store = dict(X=X, y=y, assoc=np.array(json.dumps(text)))
for k, g in enumerate(genes):
    store[f"G{k}"] = g.astype(np.int8)
np.savez_compressed(os.path.join(HERE, "ref_assoc_golden.npz"), **store)
for m in text:
    print(m, text[m]["header"][-3:], [r[-3:] for r in text[m]["rows"]])
