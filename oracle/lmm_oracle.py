"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the FastLMM score step.  Pinned on the
reference's own FastLMM.cpp compiled against oracle/eigen_standin (oracle/_ref/libskat_ref.so,
tests/test_oracle_pin_reference_skat.py::test_live_fastlmm_score_step; the reference computes in float32 => 2e-3);
this restatement follows its formulas line by line in float64 on the same float32 inputs.

  fit_null_given_delta   FastLMM::Impl::getBetaSigma2 (regression/FastLMM.cpp:300-330, MLE: sigma2 = SSR / n) and the
                         members FitNullModel leaves behind: ux, uy rotated (:52-54), uResid (:126), scaledK (:131-138)
  score                  FastLMM::Impl::TestCovariate, score branch (:215-249)
"""
import numpy as np
from scipy import stats


def fit_null_given_delta(U, lam, X, y, delta):
    lam = np.abs(np.asarray(lam, dtype=np.float64))            # lambda.cwiseAbs(), :46-50
    U = np.asarray(U, dtype=np.float64)
    ux = U.T @ X
    uy = U.T @ y
    d = 1.0 / (lam + delta)
    beta = np.linalg.solve(ux.T @ (d[:, None] * ux), ux.T @ (d * uy))
    resid = uy - ux @ beta
    sigma2 = float(np.sum(resid * resid * d) / len(y))         # MLE
    return dict(ux=ux, uResid=resid, sigma2=sigma2, beta=beta, delta=delta, lam=lam)


def score(U, nm, g):
    """g: (N,) genotypes of one variant.  Returns (Ustat, Vstat, stat, pvalue)."""
    U = np.asarray(U, dtype=np.float64)
    d = 1.0 / (nm["lam"] + nm["delta"])
    ux = nm["ux"]
    ug = U.T @ (g - g.mean())                                  # needToCenterGentype, :218-221
    Ustat = float(np.sum(ug * nm["uResid"] * d) / nm["sigma2"])
    scaledK_ug = d * ug - d * (ux @ np.linalg.solve(ux.T @ (d[:, None] * ux), ux.T @ (d * ug)))
    Vstat = float(ug @ scaledK_ug / nm["sigma2"])
    if Vstat > 0.0:
        stat = Ustat * Ustat / Vstat
        return Ustat, Vstat, stat, float(stats.chi2.sf(stat, 1))
    return Ustat, Vstat, 0.0, 1.0


def meta_cov(U, nm, G, pos, chrom, window):
    """MetaCovFamQtl (src/Model.cpp:437-498) through MetaCovTest::printCovariance (:934-1004) for the variants in the columns
    of G (N, nv): TransformCentered x~ = U'(g - gbar) (regression/FastLMM.cpp:611-625), covXX = x~_v' D x~_w / sigma2 (:538-551),
    covXZ = x~' D ux / sigma2 (:552-595), covZZ = ux' D ux / sigma2, entry = (covXX - covXZ_v covZZ^-1 covXZ_w') / N.
    Returns a dict {(v, w): value} for w >= v inside v's window, both polymorphic (monomorphic variants are never queued)."""
    U = np.asarray(U, dtype=np.float64)
    N, nv = G.shape
    d = 1.0 / (nm["lam"] + nm["delta"])
    ux = nm["ux"]
    Xt = U.T @ (G - G.mean(axis=0, keepdims=True))            # (N, nv)
    covZZ = ux.T @ (d[:, None] * ux) / nm["sigma2"]
    covZZInv = np.linalg.inv(covZZ)
    covXZ = (Xt * d[:, None]).T @ ux / nm["sigma2"]           # (nv, C)
    poly = [G[:, j].min() != G[:, j].max() for j in range(nv)]
    out = {}
    for v in range(nv):
        if not poly[v]:
            continue
        for w in range(v, nv):
            if chrom[w] != chrom[v] or pos[w] - pos[v] > window:
                break
            if not poly[w]:
                continue
            xx = float(np.sum(Xt[:, v] * d * Xt[:, w])) / nm["sigma2"]
            out[(v, w)] = (xx - float(covXZ[v] @ covZZInv @ covXZ[w])) / N
    return out
