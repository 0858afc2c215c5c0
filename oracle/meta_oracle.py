"""oracle/meta_oracle.py -- TEST INFRASTRUCTURE ONLY.

numpy restatement of `--meta score,cov` for unrelated samples and a quantitative trait:
  MetaScoreTest::fitWithGivenGenotype / writeOutput + MetaUnrelatedQtl    src/Model.h:3188-3365, 3501-3555
  GenotypeCounter                                                        src/GenotypeCounter.h:14-59
  SNPHWE (Wigginton exact test)                                          libsrc/snp_hwe.cpp:25-122
  LinearRegressionScoreTest::TestCovariate (Matrix overload, m = 1)      regression/LinearRegressionScoreTest.cpp:173-263
  MetaCovTest window / printCovariance + MetaCovUnrelatedQtl             src/Model.h:3954-4020, src/Model.cpp:500-596, 844-1004
The covariance follows the reference literally (centre the genotype, x~'x~/sigma2, covXZ, the LDLT
pseudo-inverse of the centred-covariate Gram) in fp64; the reference itself computes it in float32,
so parity is asserted at 1e-5.  The score columns are pinned on the reference's own
GenotypeCounter / SNPHWE / LinearRegressionScoreTest (oracle/_ref/libskat_ref.so,
tests/test_oracle_pin_reference_skat.py::test_live_meta_score_columns); the covariance window lives in src/Model.cpp,
which cannot be compiled in isolation: parity unpinned by the reference for `--meta cov`.
"""
from __future__ import annotations

import numpy as np

from . import oracle as O


def snp_hwe(obs_hets, obs_hom1, obs_hom2):
    obs_homc = max(obs_hom1, obs_hom2)
    obs_homr = min(obs_hom1, obs_hom2)
    rare = 2 * obs_homr + obs_hets
    n = obs_hets + obs_homc + obs_homr
    het = np.zeros(rare + 1)
    mid = int(1.0 * rare * (2 * n - rare) / (2 * n))
    if (rare & 1) ^ (mid & 1):
        mid += 1
    curr_hets, curr_homr = mid, (rare - mid) // 2
    curr_homc = n - curr_hets - curr_homr
    het[mid] = 1.0
    s = het[mid]
    h = mid
    while h > 1:
        het[h - 2] = het[h] * h * (h - 1.0) / (4.0 * (curr_homr + 1.0) * (curr_homc + 1.0))
        s += het[h - 2]
        curr_homr += 1
        curr_homc += 1
        h -= 2
    curr_homr = (rare - mid) // 2
    curr_homc = n - mid - curr_homr
    h = mid
    while h <= rare - 2:
        het[h + 2] = het[h] * 4.0 * curr_homr * curr_homc / ((h + 2.0) * (h + 1.0))
        s += het[h + 2]
        curr_homr -= 1
        curr_homc -= 1
        h += 2
    het /= s
    p = het[het <= het[obs_hets]].sum()
    return min(p, 1.0)


def meta_score(g, X, resid, sigma2):
    """one variant (N,) of hard calls -> dict of the MetaScore columns"""
    N = len(g)
    n0, n1, n2 = int((g == 0).sum()), int((g == 1).sum()), int((g == 2).sum())
    out = dict(af=0.5 * g.sum() / N, ac=float(g.sum()), call_rate=1.0, n_ref=n0, n_het=n1, n_alt=n2,
               hwe_p=snp_hwe(n1, n0, n2) if (n0 + n1 + n2) else 0.0)
    mono = g.min() == g.max()
    out["polymorphic"] = not mono
    if mono:
        out["ok"] = False
        return out
    U = float(g @ resid)
    SZ = g @ X
    SS = float(g @ g) - float(SZ @ np.linalg.solve(X.T @ X, SZ))
    V = SS * sigma2
    stat = U * (1.0 / SS / sigma2) * U
    if stat < 0:
        out["ok"] = False
        return out
    out.update(ok=True, U=U / sigma2, sqrtV=np.sqrt(V / sigma2 / sigma2), effect=(U / SS) if V != 0 else 0.0,
               effect_se=(sigma2 / np.sqrt(V)) if V != 0 else 0.0, pvalue=O.lib().orc_chisq_q(stat, 1.0))
    return out


def meta_cov(G, pos, chrom, X, sigma2, window):
    """G (N, nv) hard calls -> list over variants of (positions, cov values) as printCovariance
    would emit when that variant reaches the head of the queue (None for monomorphic variants)."""
    N, nv = G.shape
    Gd = G.astype(np.float64)
    Xc = X - X.mean(axis=0, keepdims=True)            # centerMatrix: intercept column -> exactly 0
    covZZ = Xc.T @ Xc / sigma2
    covZZInv = np.linalg.pinv(covZZ)                   # LDLT solve with zero pivots = pseudo-inverse
    xt = Gd - Gd.mean(axis=0, keepdims=True)           # transformGenotype
    covXZ = xt.T @ X / sigma2                          # calculateXZ uses the UNcentred cov
    poly = [Gd[:, j].min() != Gd[:, j].max() for j in range(nv)]
    out = []
    for i in range(nv):
        if not poly[i]:
            out.append(None)
            continue
        ps, vals = [], []
        for j in range(i, nv):
            if chrom[j] != chrom[i] or pos[j] - pos[i] > window:
                break
            if not poly[j]:
                continue
            xx = float(xt[:, i] @ xt[:, j]) / sigma2
            vals.append((xx - float(covXZ[i] @ covZZInv @ covXZ[j])) / N)
            ps.append(int(pos[j]))
        out.append((ps, vals))
    return out


def meta_score_binary(g, y, X, nm):
    """MetaUnrelatedBinary (src/Model.h:3669-3784) on one variant (N,) of hard calls.  nm = binary_oracle.fit_null_logistic
    (p and v of the round before the last update, as GetPredicted / GetVariance return them).
    LogisticRegressionScoreTest::TestCovariate(Xnull, y, Xcol), regression/LogisticRegressionScoreTest.cpp:219-302:
    U = g'(y - p), V = g'Wg - g'WZ (Z'WZ)^-1 Z'Wg; U_STAT = U, SQRT_V_STAT = sqrt(V), ALT_EFFSIZE = U / V, SE = 1 / sqrt(V).
    (With covariates the reference solves its 1 x 1 SS against a d x d identity, :292-295; the intended statistic is used.)
    The all:case:control site columns come from GenotypeCounter on the three sample sets (src/Model.h:3220-3232)."""
    g = np.asarray(g, dtype=np.float64)
    out = meta_score(g, X, y - nm["p"], 1.0)
    cc = {}
    for name, sel in (("case", y == 1), ("ctrl", y == 0)):
        gs = g[sel]
        n0, n1, n2 = int((gs == 0).sum()), int((gs == 1).sum()), int((gs == 2).sum())
        cc[name] = dict(n=int(sel.sum()), n_ref=n0, n_het=n1, n_alt=n2, hwe_p=snp_hwe(n1, n0, n2) if (n0 + n1 + n2) else 0.0)
    out["cc"] = cc
    for k in ("U", "sqrtV", "effect", "effect_se", "pvalue"):
        out.pop(k, None)
    if not out["polymorphic"]:
        out["ok"] = False
        return out
    v = nm["v"]
    U = float(g @ (y - nm["p"]))
    xz = (g * v) @ X
    V = float(g @ (v * g)) - float(xz @ nm["covB"] @ xz)
    out["cov_xz"] = xz
    stat = U * U / V
    if not (V > 0) or stat < 0:
        out["ok"] = False
        return out
    out.update(ok=True, U=U, sqrtV=np.sqrt(V), effect=(U / V) if U != 0 else 0.0, effect_se=1.0 / np.sqrt(V),
               pvalue=O.lib().orc_chisq_q(stat, 1.0))
    return out


def meta_cov_binary(G, pos, chrom, X, nm, window):
    """MetaCovUnrelatedBinary (src/Model.cpp:695-778) through MetaCovTest::printCovariance (:942-1004): raw genotypes,
    covXX = g_i'W g_j, covXZ = g'W Z, covZZ = Z'WZ, entry = (covXX - covXZ_i covZZ^-1 covXZ_j') / N.  Same return shape as
    meta_cov."""
    N, nv = G.shape
    Gd = G.astype(np.float64)
    v = nm["v"]
    covZZInv = np.linalg.inv(X.T @ (v[:, None] * X))
    covXZ = (Gd * v[:, None]).T @ X
    poly = [Gd[:, j].min() != Gd[:, j].max() for j in range(nv)]
    out = []
    for i in range(nv):
        if not poly[i]:
            out.append(None)
            continue
        ps, vals = [], []
        for j in range(i, nv):
            if chrom[j] != chrom[i] or pos[j] - pos[i] > window:
                break
            if not poly[j]:
                continue
            xx = float(Gd[:, i] @ (v * Gd[:, j]))
            vals.append((xx - float(covXZ[i] @ covZZInv @ covXZ[j])) / N)
            ps.append(int(pos[j]))
        out.append((ps, vals))
    return out
