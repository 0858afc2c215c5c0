"""GPU (-m gpu): A13, the FastLMM score step (regression/FastLMM.cpp:215-249) against the numpy restatement."""
import numpy as np
import pytest

from util import make_problem, rel

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", [(90, 600, 1, 0.7), (91, 2000, 3, 0.2), (92, 1037, 2, 3.0)])
def test_fastlmm_score_step(engine_cls, oracle, case):
    from oracle import lmm_oracle as LO
    from rvtests_b200.synth import pack_bed
    O = oracle
    seed, N, C, delta = case
    eng = engine_cls(0)
    if eng.info("tc_available") != 1:
        pytest.skip("the mixed-model score step needs the tensor-core sweep")
    rng = np.random.default_rng(seed)
    # a kinship with structure: K = Z Z' / m over standardised common variants, plus family blocks
    Z = rng.binomial(2, 0.3, size=(N, 3 * N // 2)).astype(np.float64)
    Z = (Z - Z.mean(axis=0)) / Z.std(axis=0)
    K = Z @ Z.T / Z.shape[1]
    fam = rng.integers(0, N // 4, size=N)
    K += 0.25 * (fam[:, None] == fam[None, :])
    lam, U = np.linalg.eigh(K)
    G, X, y = make_problem(O, seed, N, 150, C, maf=np.linspace(0.01, 0.4, 150), n_mono=1)
    y = y + U @ (np.sqrt(np.maximum(lam, 0)) * rng.normal(size=N)) * 0.5     # a polygenic component
    U32, lam32 = U.astype(np.float32), lam.astype(np.float32)
    nm = LO.fit_null_given_delta(U32, lam32, X, y, delta)                      # what FitNullModel leaves behind
    eng.lmm_set_null(U32, lam32, delta, nm["sigma2"], nm["uResid"], nm["ux"])
    nm32 = dict(nm, uResid=nm["uResid"].astype(np.float32).astype(np.float64), ux=nm["ux"].astype(np.float32).astype(np.float64),
                lam=np.abs(lam32.astype(np.float64)))
    M = G.shape[1]
    eng.push_i8(G[:, :64].T.copy())
    eng.push_bed(pack_bed(G[:, 64:128].T))
    eng.push_i8(G[:, 128:].T.copy())
    res = eng.lmm_flush(M)
    assert len(res) == M
    for j in range(M):
        Us, Vs, st, p = LO.score(U32, nm32, G[:, j].astype(np.float64))
        r = res[j]
        assert int(r["ok"]) == 1
        assert rel(r["af"], 0.5 * G[:, j].mean()) <= 1e-12
        scale = max(abs(Us), np.sqrt(abs(Vs)), 1e-300)
        assert abs(r["U"] - Us) <= 1e-5 * scale, (j, r["U"], Us)
        if Vs > 1e-9:
            assert rel(r["V"], Vs) <= 1e-5, (j, r["V"], Vs)
            assert abs(r["stat"] - st) <= 1e-4 * max(st, 1e-3), (j, r["stat"], st)
            assert abs(r["pvalue"] - p) <= 1e-4 * max(p, 1e-12) + 1e-12, (j, r["pvalue"], p)
    eng.close()


def test_fastlmm_covariance_band(engine_cls, oracle):
    """MetaCovFamQtl: FastLMM::TransformCentered / GetCovXX / GetCovXZ / GetCovZZ (regression/FastLMM.cpp:538-625) through
    MetaCovTest::printCovariance -- the band of rvt_lmm_meta_flush against the numpy restatement (oracle/lmm_oracle.py)."""
    from oracle import lmm_oracle as LO
    O = oracle
    seed, N, C, delta, nv = 95, 900, 2, 0.6, 150
    eng = engine_cls(0)
    if eng.info("tc_available") != 1:
        pytest.skip("the mixed-model score step needs the tensor-core sweep")
    rng = np.random.default_rng(seed)
    Z = rng.binomial(2, 0.3, size=(N, 3 * N // 2)).astype(np.float64)
    Z = (Z - Z.mean(axis=0)) / Z.std(axis=0)
    K = Z @ Z.T / Z.shape[1]
    lam, U = np.linalg.eigh(K)
    G, X, y = make_problem(O, seed, N, nv, C, maf=np.linspace(0.01, 0.4, nv), n_mono=1)
    U32, lam32 = U.astype(np.float32), lam.astype(np.float32)
    nm = LO.fit_null_given_delta(U32, lam32, X, y, delta)
    eng.lmm_set_null(U32, lam32, delta, nm["sigma2"], nm["uResid"], nm["ux"])
    nm32 = dict(nm, uResid=nm["uResid"].astype(np.float32).astype(np.float64), ux=nm["ux"].astype(np.float32).astype(np.float64),
                lam=np.abs(lam32.astype(np.float64)))
    pos = np.cumsum(rng.integers(100, 900, nv)).astype(np.int32)
    chrom = np.ones(nv, dtype=np.int32)
    chrom[110:] = 2
    window = 6000
    for b0 in range(0, nv, 64):
        eng.push_i8(G[:, b0:b0 + 64].T.copy())
    res, band, wmax = eng.lmm_meta_flush(nv, pos, chrom, window)
    eng.close()
    ref = LO.meta_cov(U32, nm32, G.astype(np.float64), pos, chrom, window)
    assert len(ref) > 500
    seen = 0
    for v in range(nv):
        for dd in range(wmax + 1):
            w = v + dd
            val = band[v, dd]
            if (v, w) in ref:
                scale = np.sqrt(abs(ref[(v, v)] * ref[(w, w)]))
                assert abs(val - ref[(v, w)]) <= 1e-6 * scale, (v, w, val, ref[(v, w)])
                seen += 1
            else:
                assert np.isnan(val), (v, w, val)
    assert seen == len(ref)
    # the score records are those of rvt_lmm_flush
    Us, Vs, st, p = LO.score(U32, nm32, G[:, 5].astype(np.float64))
    assert abs(res[5]["U"] - Us) <= 1e-5 * max(abs(Us), np.sqrt(abs(Vs)))
