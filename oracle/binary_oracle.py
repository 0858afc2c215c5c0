"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the binary-trait path of SkatTest / CMCTest / ZegginiTest.
Pinned on the reference's own LogisticRegression.cpp / LogisticRegressionScoreTest.cpp / Skat.cpp compiled against
oracle/eigen_standin (oracle/_ref/libskat_ref.so): tests/test_oracle_pin_reference_skat.py::test_live_binary_trait.

  fit_null_logistic   LogisticRegression::FitLogisticModel (regression/LogisticRegression.cpp:279-339): Newton rounds from
                      beta = 0; GetDeviance (:75-94) is evaluated with the p of the round's START, the loop stops when two
                      successive deviances differ by < 1e-3 (after round 1), and p / V / covB are NOT refreshed after the
                      last update of beta -- GetPredicted() and GetVariance() belong to the beta of one step earlier
  gene                src/Model.h:2630-2720 (binary branch :2673-2681: ynull = p, v = p(1-p), res = y - p) ->
                      Skat::Fit (regression/Skat.cpp:29-105) with P0 = V - V X (X'VX)^-1 X'V;
                      src/Model.h:845-852, 1201-1208 -> LogisticRegressionScoreTest::TestCovariate(Xnull, y, Xcol)
                      (regression/LogisticRegressionScoreTest.cpp:219-302) on the collapsed genotype:
                      U = S'(y - p), V = S'VS - S'VZ (Z'VZ)^-1 Z'VS, stat = U^2/V, p = chisq_Q(stat, 1).
                      NOT reproduced: with covariates the reference solves its 1 x 1 `SS` against a d x d identity
                      (:292-295), an out-of-bounds access; the intended m = 1 statistic is used.
"""
import numpy as np
from scipy import stats

from . import oracle as O


def fit_null_logistic(X, y, nrrounds=100):
    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    beta = np.zeros(X.shape[1])
    rounds, last = 0, -99999.0
    while rounds < nrrounds:
        p = 1.0 / (1.0 + np.exp(-(X @ beta)))
        V = p * (1.0 - p)
        D = X.T @ (V[:, None] * X)
        r = X.T @ (y - p)
        beta = beta + np.linalg.solve(D, r)
        cur = -2.0 * float(np.sum(y * np.log(p) + (1.0 - y) * np.log(1.0 - p)))
        if rounds > 1 and abs(cur - last) < 1e-3:
            return dict(beta=beta, p=p, v=V, resid=y - p, covB=np.linalg.inv(D), rounds=rounds)
        if not np.isfinite(cur) or cur == 0.0:
            raise RuntimeError("separation")
        last = cur
        rounds += 1
    raise RuntimeError("not enough iterations")


def gene(G_raw, af, X, nm, beta1=1.0, beta2=25.0):
    """G_raw (N, M) as DataConsolidator::getGenotype() hands it over; af per ORIGINAL column (index quirk F9 kept)."""
    G = np.asarray(G_raw, dtype=np.float64).copy()
    N, M = G.shape
    keep = []
    for j in range(M):                                   # convertToMinorAlleleCount + removeMonomorphicMarker
        if G[:, j].sum() > N:
            G[:, j] = 2.0 - G[:, j]
        if G[:, j].min() != G[:, j].max():
            keep.append(j)
    out = dict(m_poly=len(keep))
    if not keep:
        out["status"] = 2
        return out
    G = G[:, keep]
    w = np.zeros(len(keep))
    for i in range(len(keep)):
        f = af[i]
        f = 1.0 - f if f > 0.5 else f
        w[i] = stats.beta.pdf(f, beta1, beta2) ** 2 if f > 1e-30 else 0.0
    v, res = nm["v"], nm["resid"]
    s = G.T @ res
    out["Q"] = float(np.sum(w * s * s))
    GVX = G.T @ (v[:, None] * X)
    A = G.T @ (v[:, None] * G) - GVX @ np.linalg.solve(X.T @ (v[:, None] * X), GVX.T)
    sw = np.sqrt(w)
    lam = np.linalg.eigvalsh(sw[:, None] * A * sw[None, :])[::-1]
    r = 0
    while r < min(N, len(lam)) and lam[r] > 1e-30:
        r += 1
    out["lam"] = lam[:r]
    out["p_skat"], out["fault"] = O.skat_final_pvalue(lam[:r], out["Q"])
    for name, S in (("cmc", (np.trunc(G) > 0).any(axis=1).astype(float)), ("zeg", (np.trunc(G) > 0).sum(axis=1).astype(float))):
        U = float(S @ res)
        SZ = (v * S) @ X
        Vv = float(S @ (v * S)) - float(SZ @ nm["covB"] @ SZ)
        stat = U * U / Vv
        out[name] = dict(U=U, V=Vv, stat=stat, p=float(O.lib().orc_chisq_q(stat, 1.0)), nonref=int((S != 0).sum()))
    out["status"] = 0
    return out
