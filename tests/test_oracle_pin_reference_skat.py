"""CPU: pin the oracle on the REFERENCE's own Skat.cpp / SkatO.cpp / LinearRegression.cpp /
LinearRegressionScoreTest.cpp.  Those files need Eigen, which the reference downloads at build time and
which is absent here; oracle/Makefile compiles them UNMODIFIED against oracle/eigen_standin (a
from-scratch stand-in for the Eigen API they use) and the vendored GSL 1.16 into
oracle/_ref/libskat_ref.so.  Two layers:
  (1) tests/golden/ref_skat_golden.npz -- inputs + outputs of that build (tests/golden/make_golden_ref_skat.py),
      committed, so this layer runs wherever the repository is checked out;
  (2) the live build, when oracle/_ref is present, on further random problems.
Tolerances: Skat.cpp computes in float32 (MatrixXf), so its Q carries ~1e-7 relative noise and its
p-value inherits the float32 eigenvalues (<= 5e-4 relative seen here); SkatO.cpp and the regression
files are double precision and agree with the restatement to ~1e-10."""
import ctypes as C
import os

import numpy as np
import pytest

from util import af_of, make_problem, rel

GOLD = os.path.join(os.path.dirname(__file__), "golden", "ref_skat_golden.npz")
# float32 accumulation over N samples on the reference side (seen: Q <= 9.2e-7, p <= 3.6e-5)
TOL_Q32, TOL_P32 = 3e-6, 5e-4


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def _cases(g):
    return range(len(g["cases"]))


def test_golden_null_model(oracle, gold):
    """A3: LinearRegression::FitLinearModel (regression/LinearRegression.cpp:20-69)."""
    O = oracle
    for k in _cases(gold):
        X, y = gold[f"X{k}"], gold[f"y{k}"]
        Cc = X.shape[1]
        lin = gold[f"lin{k}"]
        nm = O.fit_null_linear(X, y)
        assert rel(nm["sigma2"], lin[0]) <= 1e-11, k
        assert np.max(np.abs(nm["beta"] - lin[1:1 + Cc])) <= 1e-10 * max(1.0, np.max(np.abs(lin[1:1 + Cc]))), k
        assert rel(np.abs(nm["resid"]).sum(), lin[1 + Cc]) <= 1e-11, k
        assert np.max(np.abs(nm["resid"][:8] - lin[2 + Cc:10 + Cc])) <= 1e-10, k


def test_golden_flip_weights_collapse(oracle, gold):
    """The stored, hand-checkable inputs of the reference calls: flipped polymorphic genotypes have minor-allele
    dosage (column mean <= 1), weights are dbeta(maf, 1, 25) = 25 (1 - maf)^24, CMC is the carrier indicator and
    Zeggini the rare-allele count of the flipped block."""
    for k in _cases(gold):
        Gf, w1 = gold[f"Gf{k}"].astype(float), gold[f"w1_{k}"]
        if Gf.shape[1] == 0:
            continue
        assert np.all(Gf.mean(0) <= 1.0 + 1e-12) and np.all(Gf.std(0) > 0)
        af = af_of(gold[f"G{k}"])[: Gf.shape[1]]  # F9: caller-order lookup by kept index
        maf = np.minimum(af, 1 - af)
        expect = np.where(maf > 1e-30, 25.0 * (1 - maf) ** 24, 0.0)
        assert np.allclose(w1, expect, rtol=1e-10)
        assert np.array_equal(gold[f"cmc{k}"], (Gf > 0).any(1).astype(np.int8))
        assert np.array_equal(gold[f"zeg{k}"], (Gf > 0).sum(1).astype(np.int16))


def test_golden_gene_statistics(oracle, gold):
    """A4/A5/A8-A10: SKAT Q and p (Skat.cpp:29-105), burden U/V/p (LinearRegressionScoreTest.cpp:173-263)."""
    O = oracle
    for k in _cases(gold):
        G, X, y = gold[f"G{k}"], gold[f"X{k}"], gold[f"y{k}"]
        nm = O.fit_null_linear(X, y)
        out, lam = O.gene(G.astype(float), af_of(G), X, nm["resid"], nm["sigma2"])
        sk, cm, zg = gold[f"skat{k}"], gold[f"cmcst{k}"], gold[f"zegst{k}"]
        assert out.m_poly == gold[f"Gf{k}"].shape[1]
        if out.m_poly == 0:
            assert out.status == 2 and sk[0] == -1
            continue
        assert rel(out.skat.Q, sk[1]) <= TOL_Q32, (k, out.skat.Q, sk[1])
        assert rel(out.skat.pvalue, sk[2]) <= TOL_P32, (k, out.skat.pvalue, sk[2])
        for pre, ref in (("cmc", cm), ("zeg", zg)):
            assert getattr(out, pre + "_ok") == (1 if ref[0] == 0 else 0), (k, pre)
            if ref[0] == 0:
                u, v = getattr(out, pre + "_U"), getattr(out, pre + "_V")
                assert abs(u - ref[1]) <= 1e-9 * max(abs(ref[1]), np.sqrt(ref[2])), (k, pre)
                assert rel(v, ref[2]) <= 1e-9, (k, pre)
                assert rel(getattr(out, pre + "_stat"), ref[3]) <= 1e-8, (k, pre)
                assert rel(getattr(out, pre + "_p"), ref[4]) <= 1e-8, (k, pre)
        assert out.cmc_nonref == int(gold[f"cmc{k}"].sum())


def test_golden_perm_statistic(gold):
    """A6: Skat::GetQFromNewResidual (Skat.cpp:107-116) = sum_j w_j (g_j . r)^2 on a shuffled residual."""
    for k in _cases(gold):
        Gf, w1, perms, sk = gold[f"Gf{k}"].astype(float), gold[f"w1_{k}"], gold[f"perm{k}"], gold[f"skat{k}"]
        if Gf.shape[1] == 0:
            continue
        q = ((perms @ Gf) ** 2 * (w1 * w1)).sum(1)
        assert np.max(np.abs(q - sk[3:]) / np.maximum(q, 1e-300)) <= 2e-6, k


def test_golden_skato(oracle, gold):
    """A7: SkatO::Fit (SkatO.cpp:100-282), incl. the single-variant FitSKAT branch (:60-98)."""
    from oracle import skato_oracle as SO
    O = oracle
    if O.ref_gsl() is None or O.ref_mix() is None:
        pytest.skip("the SKAT-O oracle runs on oracle/_ref (GSL 1.16 + reference Davies)")
    for k in _cases(gold):
        G, X, y = gold[f"G{k}"], gold[f"X{k}"], gold[f"y{k}"]
        ref = gold[f"skato{k}"]
        nm = O.fit_null_linear(X, y)
        r = SO.skato_gene(G.astype(float), af_of(G), X, nm["resid"])
        if gold[f"Gf{k}"].shape[1] == 0:
            assert not r["ok"]
            continue
        assert r["ok"] == (ref[0] == 0), k
        assert rel(r["Q"], ref[1]) <= 1e-9, (k, r["Q"], ref[1])
        assert r["rho"] == ref[2], k
        assert rel(r["pvalue"], ref[3]) <= 1e-8, (k, r["pvalue"], ref[3])


# ------------------------------------------------------------------------------------------------
# live reference build
# ------------------------------------------------------------------------------------------------
def _prep(O, G):
    Gc = np.asfortranarray(G, dtype=np.float64)
    N, M = Gc.shape
    out = np.zeros((N, M), order="F")
    keep = np.zeros(M, dtype=np.int32)
    mp = O.lib().orc_flip_minor_polymorphic(N, M, O._p(Gc), O._p(out), keep.ctypes.data_as(C.POINTER(C.c_int)), None)
    af = af_of(G)[:mp]
    w1 = np.array([O.lib().orc_skat_weight(float(a), 1.0, 25.0, 0) for a in af])
    return np.ascontiguousarray(out[:, :mp]), w1


LIVE = [(201, 400, 9, 1, 0, 2), (202, 700, 20, 2, 1, 3), (203, 1200, 40, 3, 2, 0), (204, 333, 3, 5, 0, 1),
        (205, 2500, 64, 3, 0, 6), (206, 150, 16, 1, 0, 0)]


@pytest.mark.parametrize("case", LIVE)
def test_live_reference_build(oracle, case):
    from oracle import skato_oracle as SO
    O = oracle
    if O.ref_skat() is None:
        pytest.skip("oracle/_ref/libskat_ref.so not built (no /root/reference here)")
    seed, N, M, Cc, n_mono, n_flip = case
    # keep the carrier fraction below 1: a constant CMC indicator is collinear with the intercept, V is then
    # rounding noise around 0 and its sign decides fitOK (in the reference as well)
    G, X, y = make_problem(O, seed, N, M, Cc, maf=np.linspace(0.004, 0.3 if M <= 12 else 0.03, M), n_mono=n_mono,
                           n_flip=n_flip)
    nm = O.fit_null_linear(X, y)
    lin = O.ref_linear_fit(X, y)
    assert rel(nm["sigma2"], lin["sigma2"]) <= 1e-11
    assert np.max(np.abs(nm["resid"] - lin["resid"])) <= 1e-10
    Gf, w1 = _prep(O, G)
    out, lam = O.gene(G.astype(float), af_of(G), X, nm["resid"], nm["sigma2"])
    sk = O.ref_skat_fit(lin["resid"], np.full(N, lin["sigma2"]), X, Gf, w1 * w1)
    assert rel(out.skat.Q, sk["Q"]) <= TOL_Q32
    assert rel(out.skat.pvalue, sk["pvalue"]) <= TOL_P32, (out.skat.pvalue, sk["pvalue"])
    so = O.ref_skato_fit(lin["resid"], np.full(N, lin["sigma2"]), X, Gf, w1)
    r = SO.skato_gene(G.astype(float), af_of(G), X, nm["resid"])
    assert r["ok"] == (so["rc"] == 0)
    assert rel(r["Q"], so["Q"]) <= 1e-9 and r["rho"] == so["rho"]
    assert rel(r["pvalue"], so["pvalue"]) <= 1e-8, (r["pvalue"], so["pvalue"])
    # burden: CMCTest::fit runs the MATRIX overload of TestCovariate (its collapsed genotype is a Matrix,
    # src/Model.h:855, :902); the single-column overload gives the same numbers but additionally refuses
    # V < 1e-6 (LinearRegressionScoreTest.cpp:109-112), e.g. when every sample is a carrier.
    cmc = (Gf > 0).any(1).astype(float)
    assert cmc.mean() < 1.0
    for force in (True, False):
        sc = O.ref_score_test(X, y, cmc, force_matrix=force)
        degenerate = sc["V"][0, 0] < 1e-6
        if force:
            assert (sc["rc"] == 0) == bool(out.cmc_ok)
        else:
            assert (sc["rc"] == 0) == (bool(out.cmc_ok) and not degenerate)
        if out.cmc_ok and not degenerate:
            assert abs(out.cmc_U - sc["U"][0]) <= 1e-9 * max(abs(sc["U"][0]), np.sqrt(sc["V"][0, 0]))
            assert rel(out.cmc_V, sc["V"][0, 0]) <= 1e-9
            assert rel(out.cmc_p, sc["pvalue"]) <= 1e-8


PERM = [(301, 200, 6, 1, 400, 0.05), (302, 1501, 20, 3, 300, 0.05), (303, 640, 11, 2, 250, 1.0)]


def test_live_permutation_loop(oracle):
    """A6: the oracle's permutation loop against the reference's own permute() + Permutation + Skat::GetQFromNewResidual
    (src/LinearAlgebra.h:8-21, src/Permutation.h:49-98, regression/Skat.cpp:107-116) on the same glibc rand() stream:
    identical shuffles => the same sequence of permuted Q (to the float32 noise of the reference side), hence the same
    ActualPerm / NumGreater / NumEqual, across consecutive genes that continue one stream."""
    O = oracle
    if O.ref_skat() is None:
        pytest.skip("oracle/_ref/libskat_ref.so not built (no /root/reference here)")
    problems = []
    for seed, N, M, Cc, n_perm, alpha in PERM:
        G, X, y = make_problem(O, seed, N, M, Cc, maf=np.linspace(0.01, 0.3, M), n_flip=1)
        problems.append((G, X, y, n_perm, alpha))
    refs, orcs = [], []
    for i, (G, X, y, n_perm, alpha) in enumerate(problems):  # the stream is process-wide: one pass per implementation
        lin = O.ref_linear_fit(X, y)
        Gf, w1 = _prep(O, G)
        refs.append(O.ref_skat_perm(lin["resid"], np.full(len(y), lin["sigma2"]), X, Gf, w1 * w1, n_perm=n_perm, alpha=alpha,
                                    reseed=1 if i == 0 else 0))
    for i, (G, X, y, n_perm, alpha) in enumerate(problems):
        nm = O.fit_null_linear(X, y)
        out, _ = O.gene(G.astype(float), af_of(G), X, nm["resid"], nm["sigma2"])
        orcs.append(O.gene_perm(G.astype(float), af_of(G), nm["resid"], out.skat.Q, n_perm=n_perm, alpha=alpha,
                                reseed=1 if i == 0 else 0))
    for i, (r, o) in enumerate(zip(refs, orcs)):
        assert r["actual"] == o["actual"] and r["actual"] > 0, (i, r["actual"], o["actual"])
        assert np.max(np.abs(r["q"] - o["q"]) / np.maximum(o["q"], 1e-300)) <= TOL_Q32, i
        assert (r["greater"], r["equal"]) == (o["greater"], o["equal"]), i
        assert r["p"] == pytest.approx(o["p"], rel=1e-12), i
    assert any(r["actual"] < p[3] for r, p in zip(refs, problems))  # the early stop was exercised


def test_live_fastlmm_score_step(oracle):
    """A13: the FastLMM restatement (oracle/lmm_oracle.py) against the reference's own FastLMM.cpp: FitNullModel
    (:28-140, delta by grid + Brent) then the score branch of TestCovariate (:215-249) per variant.  The restatement is
    given the reference's delta (the device API takes the fitted null model from the caller, rvt_lmm_set_null) and must
    reproduce beta, sigma2_g and every U / V / p; the reference computes in float32 => 2e-3 on the statistics."""
    from oracle import lmm_oracle as LO
    O = oracle
    if O.ref_skat() is None:
        pytest.skip("oracle/_ref/libskat_ref.so not built (no /root/reference here)")
    rng = np.random.default_rng(5)
    for N, Cc, h2 in ((240, 2, 0.5), (400, 3, 0.2)):
        Zm = rng.binomial(2, 0.3, size=(N, 600)).astype(float)
        Zm = (Zm - Zm.mean(0)) / Zm.std(0)
        K = Zm @ Zm.T / Zm.shape[1]
        lam, U = np.linalg.eigh(K)
        U32, lam32 = U.astype(np.float32), lam.astype(np.float32)
        X = np.c_[np.ones(N), rng.normal(size=(N, Cc - 1))]
        y = X @ rng.normal(size=Cc) + np.linalg.cholesky(h2 * K + (1 - h2) * np.eye(N)) @ rng.normal(size=N)
        G = rng.binomial(2, 0.15, size=(N, 12)).astype(float)
        ref = O.ref_fastlmm_score(X, y, U32, lam32, G)
        assert ref["rc"] == 0 and ref["delta"] > 0
        nm = LO.fit_null_given_delta(U32, lam32, X, y, ref["delta"])
        assert rel(nm["sigma2"], ref["sigma2"]) <= 2e-4, (nm["sigma2"], ref["sigma2"])
        assert np.max(np.abs(nm["beta"] - ref["beta"])) <= 2e-4 * max(1.0, np.max(np.abs(ref["beta"])))
        for j in range(G.shape[1]):
            Us, Vs, st, p = LO.score(U32, nm, G[:, j])
            assert abs(Us - ref["U"][j]) <= 2e-3 * max(abs(ref["U"][j]), np.sqrt(ref["V"][j])), (j, Us, ref["U"][j])
            assert rel(Vs, ref["V"][j]) <= 2e-3, (j, Vs, ref["V"][j])
            assert abs(p - ref["pvalue"][j]) <= 5e-3 * max(ref["pvalue"][j], 1e-3), (j, p, ref["pvalue"][j])


def test_live_meta_score_columns(oracle):
    """A11: the --meta score restatement (oracle/meta_oracle.py) against the reference's GenotypeCounter + SNPHWE
    (src/GenotypeCounter.h, libsrc/snp_hwe.cpp) and LinearRegressionScoreTest (Matrix overload, as MetaUnrelatedQtl
    calls it, src/Model.h:3516-3549: U / sigma2, V / sigma2^2, beta, sigma2 / sqrt(V), p)."""
    from oracle import meta_oracle as MO
    O = oracle
    if O.ref_skat() is None:
        pytest.skip("oracle/_ref/libskat_ref.so not built (no /root/reference here)")
    G, X, y = make_problem(O, 401, 3000, 40, 3, maf=np.r_[np.linspace(0.0005, 0.5, 36), [0.7, 0.9, 0.98, 0.3]], n_mono=2)
    nm = O.fit_null_linear(X, y)
    seen_mono = False
    for j in range(G.shape[1]):
        g = G[:, j].astype(float)
        o = MO.meta_score(g, X, nm["resid"], nm["sigma2"])
        c = O.ref_genotype_counter(g)
        assert (o["n_ref"], o["n_het"], o["n_alt"]) == (c["n_ref"], c["n_het"], c["n_alt"]) and c["n_missing"] == 0
        assert o["af"] == c["af"] and o["ac"] == c["ac"] and o["call_rate"] == c["call_rate"]
        assert o["hwe_p"] == pytest.approx(c["hwe_p"], rel=1e-12, abs=1e-300), j
        if not o["polymorphic"]:
            seen_mono = True
            continue
        s = O.ref_score_test(X, y, g, force_matrix=True)
        assert o["ok"] == (s["rc"] == 0)
        s2 = s["sigma2"]
        assert abs(o["U"] - s["U"][0] / s2) <= 1e-9 * max(abs(o["U"]), o["sqrtV"])
        assert rel(o["sqrtV"], np.sqrt(s["V"][0, 0] / s2 / s2)) <= 1e-9
        assert abs(o["effect"] - s["beta"][0]) <= 1e-9 * max(abs(o["effect"]), o["effect_se"])
        assert rel(o["effect_se"], s["se_beta"]) <= 1e-9
        assert rel(o["pvalue"], s["pvalue"]) <= 1e-8
    assert seen_mono


def test_live_binary_trait(oracle):
    """the binary-trait restatement (oracle/binary_oracle.py) against the reference's own
    LogisticRegression.cpp (Newton rounds, its stopping rule, and p / V left one step behind beta), Skat::Fit with
    v = p(1-p), res = y - p (the general-covariate branch of P0, Skat.cpp:58-66) and, for an intercept-only null model,
    LogisticRegressionScoreTest::TestCovariate on the CMC / Zeggini collapse."""
    from oracle import binary_oracle as BIN
    O = oracle
    if O.ref_skat() is None:
        pytest.skip("oracle/_ref/libskat_ref.so not built (no /root/reference here)")
    for seed, N, M, Cc in ((501, 800, 10, 3), (502, 1500, 25, 2), (503, 600, 8, 1)):
        G, X, _ = make_problem(O, seed, N, M, Cc, maf=np.linspace(0.005, 0.3 if M <= 12 else 0.03, M), n_flip=1, n_mono=1)
        rng = np.random.default_rng(seed)
        eta = X @ np.r_[-0.4, rng.normal(size=Cc - 1) * 0.5] + 0.4 * G[:, 0]
        y = (rng.random(N) < 1 / (1 + np.exp(-eta))).astype(float)
        nm = BIN.fit_null_logistic(X, y)
        ref = O.ref_logistic_fit(X, y)
        assert ref["rc"] == 0
        assert np.max(np.abs(nm["beta"] - ref["beta"])) <= 1e-10
        assert np.max(np.abs(nm["p"] - ref["p"])) <= 1e-12 and np.max(np.abs(nm["v"] - ref["v"])) <= 1e-12
        # the quirk: p belongs to the beta of one Newton step earlier
        assert np.max(np.abs(ref["p"] - 1 / (1 + np.exp(-(X @ ref["beta"]))))) > 0
        assert np.max(np.abs(nm["covB"] - ref["covB"])) <= 1e-10 * np.max(np.abs(ref["covB"]))
        out = BIN.gene(G.astype(float), af_of(G), X, nm)
        Gf, w1 = _prep(O, G)
        sk = O.ref_skat_fit(y - ref["p"], ref["v"], X, Gf, w1 * w1)
        assert rel(out["Q"], sk["Q"]) <= TOL_Q32, (out["Q"], sk["Q"])
        assert rel(out["p_skat"], sk["pvalue"]) <= TOL_P32, (out["p_skat"], sk["pvalue"])
        if Cc == 1:
            for name, S in (("cmc", (Gf > 0).any(1).astype(float)), ("zeg", (Gf > 0).sum(1).astype(float))):
                st = O.ref_logistic_score_test(X, y, S)
                assert st["rc"] == 0
                assert abs(out[name]["U"] - st["U"]) <= 1e-10 * max(abs(st["U"]), np.sqrt(st["V"]))
                assert rel(out[name]["V"], st["V"]) <= 1e-10
                assert rel(out[name]["p"], st["pvalue"]) <= 1e-8
        else:
            assert O.ref_logistic_score_test(X, y, (Gf > 0).any(1).astype(float))["rc"] == -3


def _binary_problem(O, seed, N, M, Cc):
    G, X, _ = make_problem(O, seed, N, M, Cc, maf=np.linspace(0.005, 0.3 if M <= 12 else 0.03, M), n_flip=1 if M > 2 else 0,
                           n_mono=1 if M > 2 else 0)
    rng = np.random.default_rng(seed)
    eta = X @ np.r_[-0.4, rng.normal(size=Cc - 1) * 0.5] + 0.4 * G[:, 0]
    y = (rng.random(N) < 1 / (1 + np.exp(-eta))).astype(float)
    return G, X, y


BINARY_SKATO_CASES = ((511, 800, 10, 3), (512, 1500, 25, 2), (513, 600, 8, 1), (514, 700, 1, 2), (515, 900, 2, 3))


@pytest.mark.parametrize("case", BINARY_SKATO_CASES)
def test_live_binary_skato(oracle, case):
    """SkatO::Fit type "D" (src/Model.h:2833-2841, 2854-2858; SkatO.cpp:72-91, 133-134, 150-158): the numpy restatement
    (oracle/skato_oracle.py with vv = p(1-p)) against the reference's own SkatO.cpp driven with the reference's own
    logistic null model, incl. the single-variant branch (FitSKAT) and a two-variant gene."""
    from oracle import binary_oracle as BIN
    from oracle import skato_oracle as SO
    O = oracle
    if O.ref_skat() is None:
        pytest.skip("oracle/_ref/libskat_ref.so not built (no /root/reference here)")
    seed, N, M, Cc = case
    G, X, y = _binary_problem(O, seed, N, M, Cc)
    ref = O.ref_logistic_fit(X, y)
    assert ref["rc"] == 0
    nm = BIN.fit_null_logistic(X, y)
    Gf, w1 = _prep(O, G)
    want = O.ref_skato_fit(y - ref["p"], ref["v"], X, Gf, w1, binary=True)
    assert want["rc"] == 0
    got = SO.skato_gene(G.astype(float), af_of(G), X, nm["resid"], vv=nm["v"])
    assert got["ok"]
    assert rel(got["Q"], want["Q"]) <= 1e-10, (got["Q"], want["Q"])
    assert got["rho"] == want["rho"]
    assert rel(got["pvalue"], want["pvalue"]) <= 1e-8, (got["pvalue"], want["pvalue"])
    # and it is NOT the quantitative formula on the same inputs (s2 and the V-weighting both matter)
    other = SO.skato_gene(G.astype(float), af_of(G), X, nm["resid"])
    assert rel(other["Q"], want["Q"]) > 1e-3
