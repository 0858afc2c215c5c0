"""CPU: pin the oracle (oracle/skat_oracle.c) on everything the reference offers for this path:
 (1) the three known-answer vectors of regression/test/testMixtureChiSquare.cpp:11-40 (program
     output), (2) the reference's own MixtureChiSquare/qfc/cdflib compiled in place
     (oracle/_ref), (3) GSL 1.16 as vendored by the reference, (4) the C1 example anchor,
 (5) the literal float32 N x N restatement of regression/Skat.cpp against the reduced fp64 algebra."""
import json
import os

import numpy as np
import pytest

from util import af_of, make_problem, rel

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_mixture_chisq_known_answers(oracle):
    O = oracle
    kat = json.load(open(os.path.join(GOLD, "mixchisq_kat.json")))
    for case in kat["cases"]:
        p, fault = O.mix_pvalue(case["lambda"], case["Q"])
        assert p == pytest.approx(case["davies"], rel=2e-5, abs=1e-300)
        assert O.liu_pvalue(case["lambda"], case["Q"]) == pytest.approx(case["liu"], rel=2e-5)


def test_golden_davies_vectors(oracle):
    """Golden vectors generated from the REFERENCE's qfc.c (tests/golden/make_golden.py)."""
    O = oracle
    g = np.load(os.path.join(GOLD, "davies_golden.npz"))
    lam, n, Q = g["lam"], g["n"], g["Q"]
    for i in range(len(Q)):
        l = lam[i, : n[i]].copy()
        p, fault = O.mix_pvalue(l, Q[i])
        assert (fault or 0) == g["fault"][i], i
        if g["fault"][i] == 0:
            assert abs(p - g["p_davies"][i]) <= 1e-12, i
        assert rel(O.liu_pvalue(l, Q[i]), g["p_liu"][i]) <= 1e-6, i


def test_against_reference_build(oracle):
    O = oracle
    if O.ref_mix() is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    rng = np.random.default_rng(7)
    for t in range(400):
        n = int(rng.integers(2, 64))
        lam = np.sort(rng.gamma(0.3, 1.0, n) * 10 ** rng.uniform(-3, 3))[::-1].copy()
        Q = lam.sum() * 10 ** rng.uniform(-1.5, 1.3)
        vo, fo, tro = O.qf(lam, Q)
        vr, fr, trr = O.qf(lam, Q, which="reference")
        assert fo == fr
        assert tro[6] == trr[6]  # identical number of errbd/truncation/cfe evaluations
        if fo in (0, 2):
            assert abs(vo - vr) <= 1e-13
        assert rel(O.liu_pvalue(lam, Q), O.liu_pvalue(lam, Q, "reference")) <= 1e-9


def test_against_gsl(oracle):
    O = oracle
    G = O.ref_gsl()
    if G is None:
        pytest.skip("oracle/_ref/libgsl_ref.so not built")
    L = O.lib()
    for x in (1e-6, 0.003, 0.2, 1.0, 3.84, 10.0, 40.0, 200.0):
        assert rel(L.orc_chisq_q(x, 1.0), G.ref_gsl_cdf_chisq_Q(x, 1.0)) <= 1e-12
    for maf in (1e-4, 0.001, 0.01, 0.05, 0.3, 0.5):
        assert rel(L.orc_beta_pdf(maf, 1.0, 25.0), G.ref_gsl_ran_beta_pdf(maf, 1.0, 25.0)) <= 1e-12
        assert rel(L.orc_beta_pdf(maf, 0.5, 0.5), G.ref_gsl_ran_beta_pdf(maf, 0.5, 0.5)) <= 1e-12


def test_c1_anchor(oracle):
    """example/example.vcf + example/pheno + example/setFile (N=9, M=3): SURVEY.md 8(c)."""
    O = oracle
    a = json.load(open(os.path.join(GOLD, "c1_anchor.json")))
    G = np.array(a["G"], dtype=float)
    y = np.array(a["y"])
    X = np.ones((9, 1))
    nm = O.fit_null_linear(X, y)
    assert nm["sigma2"] == pytest.approx(a["sigma2"], rel=1e-8)
    out, lam = O.gene(G, af_of(G), X, nm["resid"], nm["sigma2"])
    assert out.skat.Q == pytest.approx(a["Q"], rel=1e-8)
    assert out.skat.fault == 1
    assert out.skat.pvalue == pytest.approx(a["pvalue"], rel=2e-6)
    assert out.cmc_nonref == a["cmc_nonref"]
    assert lam[0] == pytest.approx(a["lambda"][0], rel=1e-8)
    assert lam[1] == pytest.approx(a["lambda"][1], rel=1e-8)
    col = np.zeros(9)
    Gc = np.asfortranarray(G)
    import ctypes as C
    O.lib().orc_zeggini_collapse(9, 3, Gc.ctypes.data_as(C.POINTER(C.c_double)), col.ctypes.data_as(C.POINTER(C.c_double)))
    assert col.tolist() == a["zeggini"]


@pytest.mark.parametrize("N,M,C", [(60, 4, 1), (300, 12, 3), (800, 30, 3)])
def test_reduced_algebra_matches_literal_float32(oracle, N, M, C):
    """regression/Skat.cpp literally (float32, explicit N x N P0) vs the O(N M^2) fp64 form."""
    import ctypes as Ct
    O = oracle
    G, X, y = make_problem(O, 11 + N, N, M, C, maf=np.linspace(0.02, 0.3, M))
    nm = O.fit_null_linear(X, y)
    Gd = np.asfortranarray(G.astype(float))
    keep = G.std(axis=0) > 0
    Gd = np.asfortranarray(Gd[:, keep])
    Mp = Gd.shape[1]
    w = np.array([O.lib().orc_skat_weight(a, 1.0, 25.0, 1) for a in af_of(G)[keep]])
    v = np.full(N, nm["sigma2"])
    Xc = np.asfortranarray(X)
    dp = Ct.POINTER(Ct.c_double)
    p = lambda a: a.ctypes.data_as(dp)
    o64, o32 = O.SkatOut(), O.SkatOut()
    l64, l32 = np.zeros(Mp), np.zeros(Mp)
    O.lib().orc_skat_reduced64(N, Mp, C, p(Gd), p(Xc), p(nm["resid"]), p(v), p(w), Ct.byref(o64), p(l64))
    O.lib().orc_skat_faithful32(N, Mp, C, p(Gd), p(Xc), p(nm["resid"]), p(v), p(w), Ct.byref(o32), p(l32))
    assert rel(o64.Q, o32.Q) <= 2e-5          # float32 accumulation noise
    assert rel(l64[0], l32[0]) <= 2e-4
    assert rel(o64.pvalue, o32.pvalue) <= 5e-3
