"""GPU: --meta score,cov statistics (src/Model.h:3155-4096, src/Model.cpp:500-1004) against the
numpy oracle: counts bit exact, score statistics 1e-6, HWE 1e-9, covariances 1e-5 (the reference
is float32 there), window membership exact -- including tile pairs across the 64-variant tiles."""
import numpy as np
import pytest

from util import rel

pytestmark = pytest.mark.gpu


def _variants(O, seed, N, nv, maf_hi=0.4):
    vid = np.arange(nv, dtype=np.uint64) + np.uint64(seed * 7919)
    rng = np.random.default_rng(seed)
    maf = 10 ** rng.uniform(np.log10(2.0 / N), np.log10(maf_hi), nv)
    maf[rng.integers(0, nv, max(1, nv // 25))] = 0.9          # ALT-major variants stay unflipped in meta mode
    G = O.synth_genotypes(seed, vid, N, maf=maf)                # (nv, N)
    for j in rng.integers(0, nv, max(1, nv // 30)):
        G[j] = 0                                                # monomorphic
    return G


@pytest.mark.parametrize("case", [(1, 3000, 50, 1, 400), (2, 5000, 200, 3, 3000), (3, 2500, 130, 2, 100000)])
def test_meta_score_and_cov(engine_cls, oracle, case):
    from oracle import meta_oracle as MO
    O = oracle
    seed, N, nv, C, window = case
    G = _variants(O, seed, N, nv)
    X, y = O.synth_covariates(seed, N, C)
    rng = np.random.default_rng(seed + 100)
    pos = np.cumsum(rng.integers(1, 60, nv)).astype(np.int32)
    chrom = np.ones(nv, dtype=np.int32)
    chrom[int(nv * 0.7):] = 2                                   # a chromosome change inside a tile
    pos[int(nv * 0.7):] -= pos[int(nv * 0.7)] - 5
    eng = engine_cls(0)
    eng.set_null_model(X, y)
    nm = O.fit_null_linear(X, y)
    for b0 in range(0, nv, 37):                                 # pushes need not be tile aligned
        eng.push_i8(G[b0:b0 + 37].copy(), None)
    vout, band, wmax = eng.meta_flush(nv, pos, chrom, window)
    ref_cov = MO.meta_cov(G.T, pos, chrom, X, nm["sigma2"], window)
    n_cov = 0
    for v in range(nv):
        ref = MO.meta_score(G[v].astype(np.float64), X, nm["resid"], nm["sigma2"])
        r = vout[v]
        assert (int(r["n_ref"]), int(r["n_het"]), int(r["n_alt"])) == (ref["n_ref"], ref["n_het"], ref["n_alt"])
        assert r["af"] == ref["af"] and r["ac"] == ref["ac"] and r["call_rate"] == 1.0
        assert rel(r["hwe_p"], ref["hwe_p"]) <= 1e-9, (v, r["hwe_p"], ref["hwe_p"])
        assert bool(r["ok"]) == ref["ok"] and bool(r["polymorphic"]) == ref["polymorphic"]
        if ref["ok"]:
            for k in ("U", "sqrtV", "effect", "effect_se"):
                assert rel(r[k], ref[k]) <= 1e-6, (v, k, r[k], ref[k])
            assert rel(r["pvalue"], ref["pvalue"]) <= 1e-6
        # covariance row: the non-NaN entries of the band, in order, are the reference's COV list
        row = band[v]
        if ref_cov[v] is None:
            assert np.all(np.isnan(row))
            continue
        ps, vals = ref_cov[v]
        got = row[~np.isnan(row)]
        got_pos = pos[v:v + wmax + 1][~np.isnan(row[: len(pos[v:v + wmax + 1])])]
        assert list(got_pos) == ps
        assert len(got) == len(vals)
        scale = max(abs(vals[0]), 1e-300)                        # entry 0 is the variant's own variance
        assert np.max(np.abs(got - np.array(vals))) <= 1e-5 * scale
        n_cov += len(vals)
    assert n_cov > nv
    eng.close()


def test_meta_score_only_and_hwe_extremes(engine_cls, oracle):
    from oracle import meta_oracle as MO
    O = oracle
    N = 4000
    rng = np.random.default_rng(5)
    rows = []
    for n1, n2 in ((0, 0), (1, 0), (0, 1), (N, 0), (0, N), (N // 2, N // 4), (17, 3), (1999, 1000), (N - 1, 0)):
        g = np.zeros(N, dtype=np.int8)
        idx = rng.permutation(N)
        g[idx[:n1]] = 1
        g[idx[n1:n1 + n2]] = 2
        rows.append(g)
    G = np.array(rows)
    X, y = O.synth_covariates(9, N, 2)
    eng = engine_cls(0)
    eng.set_null_model(X, y)
    nm = O.fit_null_linear(X, y)
    eng.push_i8(G, None)
    vout, band, wmax = eng.meta_flush(len(rows), want_cov=False)
    assert band is None
    for v in range(len(rows)):
        ref = MO.meta_score(G[v].astype(np.float64), X, nm["resid"], nm["sigma2"])
        assert rel(vout[v]["hwe_p"], ref["hwe_p"]) <= 1e-9
        assert bool(vout[v]["ok"]) == ref["ok"]
    eng.close()


def test_bolt_score_step(engine_cls, oracle):
    """BoltLMM::TestCovariate (regression/BoltLMM.cpp:315-338, projDot/projNorm2 :1064-1138,
    BoltPlinkLoader::projectCovariate .cpp:266-271) restated in numpy, for an arbitrary H^-1 y:
      g_test = [g ; Z'g],  u = g.h - (Z'g).(Z'h),  v = (|g|^2 - |Z'g|^2) * |H^-1y|^2_proj * calib / N,
      p = chisq_Q(u^2/v, 1), af = sum(g)/2N
    against rvt_set_null_residual + rvt_meta_flush."""
    O = oracle
    N, nv, C = 6000, 90, 3
    G = _variants(O, 11, N, nv, maf_hi=0.3)
    X, _ = O.synth_covariates(11, N, C)
    rng = np.random.default_rng(11)
    Z, _r = np.linalg.qr(X)                       # orthonormal covariate basis (BoltPlinkLoader z_)
    h = rng.standard_normal(N)                    # stands in for H^-1 y
    hz = Z.T @ h
    h_norm2 = float(h @ h - hz @ hz)              # projNorm2(H_inv_y_)
    calib = 1.0371
    kappa = h_norm2 * calib / N
    r_b = h - Z @ hz
    eng = engine_cls(0)
    eng.set_null_residual(X, r_b, kappa)
    for b0 in range(0, nv, 64):
        eng.push_i8(G[b0:b0 + 64].copy(), None)
    vout, _band, _w = eng.meta_flush(nv, want_cov=False)
    for j in range(nv):
        g = G[j].astype(np.float64)
        zg = Z.T @ g
        u = float(g @ h - zg @ hz)
        v = float(g @ g - zg @ zg) * h_norm2 * calib / N
        r = vout[j]
        assert r["af"] == 0.5 * g.sum() / N
        if g.min() == g.max():
            assert not r["ok"]
            continue
        assert abs(r["U"] * kappa - u) <= 1e-6 * max(abs(u), np.sqrt(v))      # U_STAT = U / sigma2
        assert rel(r["sqrtV"] ** 2 * kappa ** 2, v) <= 1e-6
        assert rel(r["pvalue"], O.lib().orc_chisq_q(u * u / v, 1.0)) <= 1e-6
    eng.close()


def test_bolt_covariance_band(engine_cls, oracle):
    """BoltLMM::GetCovXX (regression/BoltLMM.cpp:435-460) as MetaCovFamQtlBolt::calculateXX uses it (src/Model.cpp:780-805),
    printed divided by N (printCovariance, :990-996): g1'(I - ZZ')g2 * xVx_xx_ratio / N, restated in numpy, against the band of
    rvt_meta_flush with option "meta_cov_scale"."""
    O = oracle
    N, nv, C, ratio = 5000, 130, 3, 0.8731
    G = _variants(O, 12, N, nv, maf_hi=0.3)
    X, _ = O.synth_covariates(12, N, C)
    Z, _r = np.linalg.qr(X)
    rng = np.random.default_rng(12)
    h = rng.standard_normal(N)
    r_b = h - Z @ (Z.T @ h)
    pos = (500 * np.arange(nv)).astype(np.int32)
    chrom = np.ones(nv, dtype=np.int32)
    window = 20_000
    eng = engine_cls(0)
    try:
        eng.set_null_residual(X, r_b, 1.3)
        eng.set_option("meta_cov_scale", ratio)
        for b0 in range(0, nv, 64):
            eng.push_i8(G[b0:b0 + 64].copy(), None)
        vout, band, wmax = eng.meta_flush(nv, pos, chrom, window)
    finally:
        eng.close()
    Gd = G.astype(np.float64)
    PG = Gd - (Gd @ Z) @ Z.T                      # rows projected off the covariates
    poly = Gd.min(axis=1) != Gd.max(axis=1)
    checked = 0
    for i in range(nv):
        for d in range(wmax + 1):
            j = i + d
            if j >= nv or pos[j] - pos[i] > window:
                continue
            if not (poly[i] and poly[j]):
                assert np.isnan(band[i, d])
                continue
            want = float(PG[i] @ Gd[j]) * ratio / N
            # an entry is a small difference of two large sums (A_ij - B_i (X'X)^-1 B_j'): judge the error on the scale
            # of the two variants' variances (the covariates enter through 2^-30 fixed-point digits)
            scale = np.sqrt(float(PG[i] @ Gd[i]) * float(PG[j] @ Gd[j])) * ratio / N
            assert abs(band[i, d] - want) <= 1e-7 * scale, (i, d, band[i, d], want, scale)
            checked += 1
    assert checked > 1000


def _binary_problem(O, seed, N, nv, C):
    G = _variants(O, seed, N, nv)
    X, _y = O.synth_covariates(seed, N, C)
    rng = np.random.default_rng(seed + 7)
    eta = -0.4 + (0.5 * X[:, 1] if C > 1 else 0.0) + 0.3 * (G[3] - G[3].mean())
    y = (rng.uniform(size=N) < 1.0 / (1.0 + np.exp(-eta))).astype(np.float64)
    return G, X, y


@pytest.mark.parametrize("case", [(11, 3000, 50, 1, 400), (12, 5000, 200, 3, 3000), (13, 70000, 90, 2, 100000)])
def test_meta_binary_trait_score_and_cov(engine_cls, oracle, case):
    """MetaUnrelatedBinary (src/Model.h:3669-3784) + MetaCovUnrelatedBinary (src/Model.cpp:695-778): the weighted sums come
    off the integer tensor-core sweep through base-128 digits of the weights (csrc/meta.cuh); counts among cases / controls
    exact, statistics 1e-6 (the weights are rounded to 2^-28), band 1e-6 of the variant's own variance."""
    from oracle import meta_oracle as MO
    from oracle import binary_oracle as BIN
    O = oracle
    seed, N, nv, C, window = case
    G, X, y = _binary_problem(O, seed, N, nv, C)
    rng = np.random.default_rng(seed + 100)
    pos = np.cumsum(rng.integers(1, 60, nv)).astype(np.int32)
    chrom = np.ones(nv, dtype=np.int32)
    chrom[int(nv * 0.7):] = 2
    pos[int(nv * 0.7):] -= pos[int(nv * 0.7)] - 5
    eng = engine_cls(0)
    if eng.info("tc_available") != 1:
        pytest.skip("binary-trait meta statistics need the tensor-core sweep")
    eng.set_null_model(X, y, binary=True)
    nm = BIN.fit_null_logistic(X, y)
    for b0 in range(0, nv, 37):
        eng.push_i8(G[b0:b0 + 37].copy(), None)
    vout, band, wmax = eng.meta_flush(nv, pos, chrom, window)
    cc, xz, zz = eng.meta_binary_extras(nv, C)
    assert np.max(np.abs(zz - X.T @ (nm["v"][:, None] * X))) <= 1e-9 * np.max(np.abs(zz))
    ref_cov = MO.meta_cov_binary(G.T, pos, chrom, X, nm, window)
    n_cov = n_ok = 0
    for v in range(nv):
        ref = MO.meta_score_binary(G[v].astype(np.float64), y, X, nm)
        r = vout[v]
        assert (int(r["n_ref"]), int(r["n_het"]), int(r["n_alt"])) == (ref["n_ref"], ref["n_het"], ref["n_alt"])
        assert r["af"] == ref["af"] and r["ac"] == ref["ac"]
        assert rel(r["hwe_p"], ref["hwe_p"]) <= 1e-9
        for w, name in enumerate(("case", "ctrl")):
            rc = ref["cc"][name]
            assert (int(cc[v]["n"][w]), int(cc[v]["n_ref"][w]), int(cc[v]["n_het"][w]), int(cc[v]["n_alt"][w])) == \
                (rc["n"], rc["n_ref"], rc["n_het"], rc["n_alt"]), (v, name)
            assert rel(cc[v]["hwe_p"][w], rc["hwe_p"]) <= 1e-9, (v, name, cc[v]["hwe_p"][w], rc["hwe_p"])
        assert bool(r["ok"]) == ref["ok"] and bool(r["polymorphic"]) == ref["polymorphic"], (v, r, ref)
        if ref["ok"]:
            n_ok += 1
            assert abs(r["U"] - ref["U"]) <= 1e-6 * max(abs(ref["U"]), ref["sqrtV"]), (v, r["U"], ref["U"])
            for k in ("sqrtV", "effect_se"):
                assert rel(r[k], ref[k]) <= 1e-6, (v, k, r[k], ref[k])
            assert abs(r["effect"] - ref["effect"]) <= 1e-6 * max(abs(ref["effect"]), ref["effect_se"])
            assert abs(r["pvalue"] - ref["pvalue"]) <= 1e-6 * max(ref["pvalue"], 1e-12) + 1e-9
            assert np.max(np.abs(xz[v] - ref["cov_xz"])) <= 1e-6 * max(np.max(np.abs(ref["cov_xz"])), 1e-12)
        row = band[v]
        if ref_cov[v] is None:
            assert np.all(np.isnan(row))
            continue
        ps, vals = ref_cov[v]
        got = row[~np.isnan(row)]
        got_pos = pos[v:v + wmax + 1][~np.isnan(row[: len(pos[v:v + wmax + 1])])]
        assert list(got_pos) == ps
        scale = max(abs(vals[0]), 1e-300)
        assert np.max(np.abs(got - np.array(vals))) <= 1e-6 * scale, (v, got[:4], vals[:4])
        n_cov += len(vals)
    assert n_cov > nv and n_ok > nv // 2
    # score only: no band, same records
    for b0 in range(0, nv, 64):
        eng.push_i8(G[b0:b0 + 64].copy(), None)
    vout2, _b, _w = eng.meta_flush(nv, want_cov=False)
    for k in ("U", "sqrtV", "pvalue"):
        assert np.allclose(vout2[k], vout[k], rtol=1e-12, atol=0)
    eng.close()
