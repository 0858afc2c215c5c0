"""GPU (-m gpu): the CUDA path, through the C ABI, against outputs of the REFERENCE's own sources.
tests/golden/ref_skat_golden.npz holds inputs and the results of regression/Skat.cpp, SkatO.cpp,
LinearRegression.cpp and LinearRegressionScoreTest.cpp compiled unmodified in the build container
(oracle/Makefile -> oracle/_ref/libskat_ref.so, Eigen replaced by oracle/eigen_standin;
generator: tests/golden/make_golden_ref_skat.py).  No oracle in between: device numbers vs reference numbers.
Tolerances: Skat.cpp is float32 (Eigen::MatrixXf) => its Q carries up to ~1e-6 relative accumulation noise and
its p-value the float32 eigenvalues; SkatO.cpp and the score test are double => 1e-6 / 1e-5 as in the oracle tests."""
import os

import numpy as np
import pytest

from util import af_of, rel

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "ref_skat_golden.npz")
TOL_Q32, TOL_P32 = 3e-6, 5e-4


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


@pytest.fixture(scope="module")
def eng(engine_cls):
    e = engine_cls(0)
    yield e
    e.close()


def _check_skat_burden(r, gold, k, ctx):
    sk, cm, zg = gold[f"skat{k}"], gold[f"cmcst{k}"], gold[f"zegst{k}"]
    m_poly = gold[f"Gf{k}"].shape[1]
    assert int(r["m_poly"]) == m_poly, ctx
    if m_poly == 0:
        assert int(r["status"]) == 2, ctx
        return
    assert int(r["status"]) == 0, ctx
    assert rel(r["Q"], sk[1]) <= TOL_Q32, (ctx, r["Q"], sk[1])
    assert rel(r["p_skat"], sk[2]) <= TOL_P32, (ctx, r["p_skat"], sk[2])
    assert int(r["cmc_nonref"]) == int(gold[f"cmc{k}"].sum()), ctx
    for pre, ref in (("cmc", cm), ("zeg", zg)):
        assert int(r[pre + "_ok"]) == (1 if ref[0] == 0 else 0), (ctx, pre)
        if ref[0] == 0:
            assert abs(r[pre + "_U"] - ref[1]) <= 1e-6 * max(abs(ref[1]), np.sqrt(ref[2])), (ctx, pre, r[pre + "_U"], ref[1])
            assert rel(r[pre + "_V"], ref[2]) <= 1e-6, (ctx, pre)
            assert rel(r[pre + "_stat"], ref[3]) <= 1e-5, (ctx, pre)
            assert rel(r[pre + "_p"], ref[4]) <= 1e-4, (ctx, pre, r[pre + "_p"], ref[4])


@pytest.mark.parametrize("which", [1, 2])  # RVT_ENGINE_SIMT, RVT_ENGINE_TC
def test_skat_and_burden_vs_reference_outputs(eng, gold, which):
    if which == 2 and eng.info("tc_available") != 1:
        pytest.skip("tensor-core engine not available in this build")
    eng.set_option("engine", which)
    try:
        for k in range(len(gold["cases"])):
            G, X, y = gold[f"G{k}"], gold[f"X{k}"], gold[f"y{k}"]
            eng.set_null_model(X, y)
            nm = eng.get_null_model()
            lin = gold[f"lin{k}"]
            assert rel(nm["sigma2"], lin[0]) <= 1e-10, k
            assert np.max(np.abs(nm["resid"][:8] - lin[2 + X.shape[1]:10 + X.shape[1]])) <= 1e-9, k
            eng.push_i8(np.ascontiguousarray(G.T), af_of(G))
            r = eng.flush()[0]
            _check_skat_burden(r, gold, k, f"case {k} engine {which}")
    finally:
        eng.set_option("engine", 0)


def test_skato_vs_reference_outputs(eng, gold):
    eng.set_option("engine", 0)
    eng.set_option("skato", 1)
    try:
        for k in range(len(gold["cases"])):
            G, X, y = gold[f"G{k}"], gold[f"X{k}"], gold[f"y{k}"]
            ref = gold[f"skato{k}"]
            eng.set_null_model(X, y)
            eng.push_i8(np.ascontiguousarray(G.T), af_of(G))
            r = eng.flush()[0]
            _check_skat_burden(r, gold, k, f"case {k} with skato on")
            if gold[f"Gf{k}"].shape[1] == 0:
                continue
            assert int(r["skato_ok"]) == (1 if ref[0] == 0 else 0), k
            if ref[0] == 0:
                assert rel(r["skato_Q"], ref[1]) <= 1e-6, (k, r["skato_Q"], ref[1])
                assert r["skato_rho"] == ref[2], (k, r["skato_rho"], ref[2])
                assert rel(r["skato_p"], ref[3]) <= 1e-5, (k, r["skato_p"], ref[3])
    finally:
        eng.set_option("skato", 0)
