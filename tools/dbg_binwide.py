import sys, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from util import af_of, make_problem
from oracle import oracle as O
from oracle import binary_oracle as BIN
import rvtests_b200
from rvtests_b200.synth import pack_bed
seed, N, M, C, miss = 141, 3000, 100, 3, 0.0
G, X, _ = make_problem(O, seed, N, M, C, maf=np.linspace(0.003, 0.05, M), n_flip=2, n_mono=1)
rng = np.random.default_rng(seed)
eta = -0.7 + X[:, 1:] @ np.full(C - 1, 0.4)
y = (rng.random(N) < 1.0 / (1.0 + np.exp(-eta))).astype(np.float64)
nm = BIN.fit_null_logistic(X, y)
af = af_of(G)
eng = rvtests_b200.GeneEngine(0)
eng.set_null_model(X, y, binary=True)
eng.push_bed(pack_bed(G.T), af)
eng.push_i8(G.T.copy(), af)
res = eng.flush()
ref = BIN.gene(G.astype(float), af, X, nm)
print("engine m_poly", res["m_poly"], "status", res["status"], "Q", res["Q"], "ref m_poly", ref["m_poly"], "Q", ref["Q"])
cnt = np.array([(G[:, j] == 0).sum() for j in range(M)]), np.array([(G[:, j] == 1).sum() for j in range(M)]), np.array([(G[:, j] == 2).sum() for j in range(M)])
print("n1 min", cnt[1].min(), "cols with n1==0:", int((cnt[1] == 0).sum()), "n2==0:", int((cnt[2] == 0).sum()))
for Mx in (65, 70, 90, 128, 129):
    G2, _, _ = make_problem(O, seed + Mx, N, Mx, C, maf=np.linspace(0.003, 0.05, Mx))
    eng.push_i8(G2.T.copy(), af_of(G2))
    r = eng.flush()[0]
    ref2 = BIN.gene(G2.astype(float), af_of(G2), X, nm)
    print(Mx, "m_poly", int(r["m_poly"]), ref2["m_poly"], "Q rel", abs(r["Q"] - ref2["Q"]) / ref2["Q"], "status", int(r["status"]))
eng.close()
