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
