"""oracle/skato_oracle.py -- TEST INFRASTRUCTURE ONLY.

numpy restatement of SKAT-O (quantitative traits, and binary traits = type "D"), following the reference line by line:
  SkatOTest::fit          src/Model.h:2787-2860   (UN-squared Beta weights, OLS null, v = sigma2)
  SkatO::Fit / FitSKAT    regression/SkatO.cpp:101-281, 60-99
  getEigen/getMoment/getPvalByMoment/getQvalByMoment/capRhos   regression/SkatO.cpp:350-455
  integrandDavies / integrandLiu                                regression/SkatO.cpp:303-337
  Integration::integrateLU (gsl_integration_qags, limit 1000)   regression/GSLIntegration.cpp:37-49
Linear algebra is numpy (Eigen is not vendored by the reference); the special functions and the
quadrature are the reference's OWN third-party code when oracle/_ref is built: GSL 1.16
(gsl_cdf_chisq_Q/P/Qinv, gsl_ran_chisq_pdf, gsl_integration_qags) and the reference's
MixtureChiSquare (Davies / Liu).  Without oracle/_ref it falls back to scipy + the C oracle and
says so in `info["backend"]`.  Parity status: pinned on the reference's own SkatO.cpp compiled
against oracle/eigen_standin (oracle/_ref/libskat_ref.so; Q / rho / p agree to ~1e-14,
tests/test_oracle_pin_reference_skat.py and tests/golden/ref_skat_golden.npz).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import oracle as O


class _Backend:
    def __init__(self):
        self.gsl = O.ref_gsl()
        self.mix = "reference" if O.ref_mix() is not None else "oracle"
        if self.gsl is None:
            from scipy import stats  # noqa
            self.name = "scipy + oracle Davies"
        else:
            self.name = "GSL 1.16 (vendored tarball) + " + ("reference" if self.mix == "reference" else "oracle") + " Davies"

    def chisq_Q(self, x, df):
        if self.gsl is not None:
            return self.gsl.ref_gsl_cdf_chisq_Q(float(x), float(df))
        from scipy import stats
        return float(stats.chi2.sf(x, df))

    def chisq_P(self, x, df):
        if self.gsl is not None:
            return self.gsl.ref_gsl_cdf_chisq_P(float(x), float(df))
        from scipy import stats
        return float(stats.chi2.cdf(x, df))

    def chisq_Qinv(self, q, df):
        if self.gsl is not None:
            return self.gsl.ref_gsl_cdf_chisq_Qinv(float(q), float(df))
        from scipy import stats
        return float(stats.chi2.isf(q, df))

    def chisq_pdf(self, x, df):
        if self.gsl is not None:
            return self.gsl.ref_gsl_ran_chisq_pdf(float(x), float(df))
        from scipy import stats
        return float(stats.chi2.pdf(x, df))

    def qags(self, f, a, b, epsabs, epsrel, limit=1000):
        """returns (status, result, abserr, n_intervals)"""
        if self.gsl is not None:
            cb = O._QAGS_CB(lambda x, _: f(x))
            res, err, nint = C.c_double(0), C.c_double(0), C.c_int(0)
            st = self.gsl.ref_gsl_qags(cb, None, a, b, epsabs, epsrel, limit, C.byref(res), C.byref(err), C.byref(nint))
            return st, res.value, err.value, nint.value
        from scipy import integrate
        r, e, info, *msg = integrate.quad(f, a, b, epsabs=epsabs, epsrel=epsrel, limit=limit, full_output=1)
        return (1 if msg else 0), r, e, info["last"]


def get_eigen(K):
    """SkatO.cpp:350-382: ascending eigenvalues; keep those >= mean(positive)/1e5, descending."""
    values = np.linalg.eigvalsh(K)
    pos = values[values > 0]
    if len(pos) == 0:
        return None
    t = pos.sum() / len(pos) / 100000.0
    keep = len(values)
    for v in values:
        if v < t:
            keep -= 1
        else:
            break
    return values[::-1][:keep].copy()


def get_moment(lam):
    """SkatO.cpp:383-416 -> (muQ, varQ, df)"""
    c = [np.sum(lam), np.sum(lam ** 2), np.sum(lam ** 2 * lam), np.sum((lam ** 2) ** 2)]
    sigmaQ = np.sqrt(2 * c[1])
    s1 = c[2] / c[1] / np.sqrt(c[1])
    s2 = c[3] / (c[1] * c[1])
    if s1 * s1 > s2:
        a = 1 / (s1 - np.sqrt(s1 * s1 - s2))
        d = (s1 * a - 1.0 * a * a)
        l = a * a - 2 * d
    else:
        l = 1.0 / s2
    return c[0], sigmaQ * sigmaQ, l


def skato(G, w, X, res, be: _Backend | None = None, vv=None):
    """SkatO::Fit.  G (N, M) flipped/polymorphic, w unsquared weights, X (N, C) incl. intercept,
    res null residuals.  vv = None: quantitative trait (type "C"); vv = per-sample variance p(1-p)
    of the logistic null: binary trait (type "D": s2 = 1, SkatO.cpp:133-134; Z1 = V^1/2 (G - X (X'VX)^-1 X'VG),
    :150-158; FitSKAT :72-91).  Returns dict(ok, Q, rho, pvalue, info)."""
    be = be or _Backend()
    N, M = G.shape
    G = G.astype(np.float64) * w[None, :]
    info = {"backend": be.name}
    binary = vv is not None
    if binary:
        vv = np.asarray(vv, dtype=np.float64)

    def davies(Q, lam):
        p, _ = O.mix_pvalue(lam, Q, be.mix)
        return p

    def liu(Q, lam):
        return O.liu_pvalue(lam, Q, be.mix)

    if M == 1:  # FitSKAT, SkatO.cpp:60-99
        temp = res @ G
        Q = float(temp @ temp)
        if not binary:
            s2 = float(res @ res) / (N - 1)
            Q = Q / s2
            W = G.T @ G - (G.T @ X) @ np.linalg.solve(X.T @ X, X.T @ G)
        else:
            VG, VX = vv[:, None] * G, vv[:, None] * X
            W = G.T @ VG - (G.T @ VX) @ np.linalg.solve(X.T @ VX, X.T @ VG)
        Q = Q / 2.0
        W = W / 2
        lam = get_eigen(W)
        if lam is None:
            return dict(ok=False, info=info)
        return dict(ok=True, Q=Q, rho=0.0, pvalue=davies(Q, lam), info=info)

    rhos_orig = np.array([i / 10 for i in range(11)])
    rhos = np.minimum(rhos_orig, 0.999)
    s2 = 1.0 if binary else float(np.linalg.norm(res) ** 2) / (N - 1)
    v = res @ G
    Qs = np.array([(v @ ((1 - r) * np.eye(M) + r * np.ones((M, M)) - 0) @ v) if False else
                   float(v @ (np.where(np.eye(M) > 0, 1.0, r)) @ v) for r in rhos]) / s2 / 2.0
    if not binary:
        Z1 = (G - X @ np.linalg.solve(X.T @ X, X.T @ G)) / np.sqrt(2)
    else:
        vs = np.sqrt(vv)[:, None]
        Z1 = (vs * G - vs * X @ np.linalg.solve(X.T @ (vv[:, None] * X), X.T @ (vv[:, None] * G))) / np.sqrt(2)
    lambdas = []
    for r in rhos:
        R = np.where(np.eye(M) > 0, 1.0, r)
        L = np.linalg.cholesky(R)
        Z2 = Z1 @ L
        lam = get_eigen(Z2.T @ Z2)
        if lam is None:
            return dict(ok=False, info=info)
        lambdas.append(lam)
    z_bar = Z1.sum(axis=1) / M
    z_norm = float(z_bar @ z_bar)
    zz = z_bar @ Z1
    ZMZ = np.outer(zz, zz) / z_norm
    ZIMZ = Z1.T @ Z1 - ZMZ
    lam = get_eigen(ZIMZ)
    if lam is None:
        return dict(ok=False, info=info)
    VarZeta = 4.0 * float((ZMZ * ZIMZ).sum())
    MuQ = float(lam.sum())
    VarQ = 2.0 * float((lam * lam).sum()) + VarZeta
    temp = float((lam * lam).sum())
    KerQ = float((lam ** 4).sum()) / temp / temp * 12
    Df = 12 / KerQ
    taus = M * M * rhos * z_norm + (1.0 - rhos) * float((zz ** 2).sum()) / z_norm
    moments = [get_moment(l) for l in lambdas]
    pvals = np.array([be.chisq_Q((Qs[i] - m[0]) / np.sqrt(m[1]) * np.sqrt(2.0 * m[2]) + m[2], m[2])
                      for i, m in enumerate(moments)])
    minIndex = 0
    minP = pvals[0]
    for i in range(1, 11):
        if pvals[i] < minP:
            minP, minIndex = pvals[i], i
    rho = rhos[minIndex]
    Q = Qs[minIndex]
    Qs_minP = np.array([(be.chisq_Qinv(minP, m[2]) - m[2]) / np.sqrt(2.0 * m[2]) * np.sqrt(m[1]) + m[0]
                        for m in moments])
    lamsum = float(lam.sum())
    neval = [0]

    def integrand_davies(x):
        neval[0] += 1
        kappa = None
        for i in range(11):
            vv = (Qs_minP[i] - taus[i] * x) / (1.0 - rhos[i])
            if i == 0 or vv < kappa:
                kappa = vv
        if kappa > lamsum * 10000:
            temp = 0.0
        else:
            Qx = (kappa - MuQ) * np.sqrt(VarQ - VarZeta) / np.sqrt(VarQ) + MuQ
            temp = davies(Qx, lam)
            if temp <= 0.0 or temp == 1.0:
                temp = liu(Qx, lam)
        return (1.0 - temp) * be.chisq_pdf(x, 1.0)

    def integrand_liu(x):
        kappa = min((Qs_minP[i] - taus[i] * x) / (1.0 - rhos[i]) for i in range(11))
        Qx = (kappa - MuQ) / np.sqrt(VarQ) * np.sqrt(2.0 * Df) + Df
        return be.chisq_P(Qx, Df) * be.chisq_pdf(x, 1.0)

    st, result, abserr, nint = be.qags(integrand_davies, 0.0, 40.0, 1e-25, 0.0001220703)
    info.update(qags_status=st, qags_intervals=nint, davies_evals=neval[0])
    if st:
        st2, result, abserr, nint = be.qags(integrand_liu, 0.0, 40.0, 1e-25, 0.0001220703)
        info.update(qags_status_liu=st2)
    pvalue = 1.0 - result
    multi = 3
    if pvalue <= 0:
        p = minP * multi
        if pvalue < p:
            pvalue = p
    if pvalue == 0.0:
        pvalue = pvals[0]
        for i in range(1, 11):
            if pvals[i] > 0 and pvals[i] < pvalue:
                pvalue = pvals[i]
    if rho >= 0.999:
        rho = 1.0
    info.update(minP=float(minP), pvals=pvals, Qs=Qs, taus=taus, MuQ=MuQ, VarQ=VarQ, VarZeta=VarZeta, Df=Df,
                Qs_minP=Qs_minP, lam=lam, moments=moments, integral=result)
    return dict(ok=True, Q=float(Q), rho=float(rho), pvalue=float(pvalue), info=info)


def skato_gene(G_raw, af, X, resid, beta1=1.0, beta2=25.0, vv=None):
    """SkatOTest::fit on one gene: flip/drop monomorphic (DataConsolidator.cpp:46-142), UN-squared
    weights with the caller-order AF lookup (src/Model.h:2799-2813), then SkatO::Fit -- type "C", or
    type "D" when vv (the logistic null's p(1-p), src/Model.h:2833-2841, 2854-2858) is given."""
    Gc = np.asfortranarray(G_raw, dtype=np.float64)
    N, M = Gc.shape
    out = np.zeros((N, M), order="F")
    keep = np.zeros(M, dtype=np.int32)
    ip = C.POINTER(C.c_int)
    mp = O.lib().orc_flip_minor_polymorphic(N, M, O._p(Gc), O._p(out), keep.ctypes.data_as(ip), None)
    if mp == 0:
        return dict(ok=False, na=True)
    w = np.array([O.lib().orc_skat_weight(float(af[i]), beta1, beta2, 0) for i in range(mp)])
    return skato(np.ascontiguousarray(out[:, :mp]), w, np.asarray(X, dtype=np.float64), np.asarray(resid), vv=vv)
