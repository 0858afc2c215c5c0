"""CPU: the product's __host__ __device__ math headers (rvtests_b200/csrc/{mathdev,davies,eigen}.cuh)
compiled with g++ (tests/hostcheck) must agree with the oracle and with the golden vectors.  This
checks LOGIC only; the GPU parity tests (-m gpu) check the kernels that instantiate the same code."""
import ctypes as C
import os

import numpy as np
import pytest

from util import rel

GOLD = os.path.join(os.path.dirname(__file__), "golden")
dp = C.POINTER(C.c_double)


def _p(a):
    return a.ctypes.data_as(dp)


def test_davies_liu_vs_golden(hostcheck):
    g = np.load(os.path.join(GOLD, "davies_golden.npz"))
    for i in range(len(g["Q"])):
        l = g["lam"][i, : g["n"][i]].copy()
        f = C.c_int(0)
        p = hostcheck.hc_mixchisq(_p(l), len(l), float(g["Q"][i]), C.byref(f))
        assert f.value == g["fault"][i], i
        if f.value == 0:
            assert abs(p - g["p_davies"][i]) <= 1e-12, i
        else:
            assert p == -1.0
        assert rel(hostcheck.hc_liu(_p(l), len(l), float(g["Q"][i])), g["p_liu"][i]) <= 1e-6, i


def test_davies_vs_oracle_random(hostcheck, oracle):
    rng = np.random.default_rng(3)
    for t in range(1500):
        n = int(rng.integers(1, 64))
        lam = np.sort(rng.gamma(0.3, 1.0, n) * 10 ** rng.uniform(-3, 3))[::-1].copy()
        Q = lam.sum() * 10 ** rng.uniform(-1.5, 1.3)
        f = C.c_int(0)
        p = hostcheck.hc_mixchisq(_p(lam), n, Q, C.byref(f))
        po, fo = oracle.mix_pvalue(lam, Q)
        assert f.value == fo
        assert abs(p - po) <= 1e-13
        assert rel(hostcheck.hc_liu(_p(lam), n, Q), oracle.liu_pvalue(lam, Q)) <= 1e-10


def test_budget_exhaustion_fault4(hostcheck, oracle):
    """lim small enough that the evaluation budget (qfc.c:77-83 counter/longjmp) trips."""
    lam = np.array([5.0, 1.0, 0.2, 0.01])
    for lim in (3, 10, 30):
        f = C.c_int(0)
        hostcheck.hc_qf(_p(lam), 4, 3.0, lim, 1e-6, C.byref(f))
        v, fo, tr = oracle.qf(lam, 3.0, lim=lim)
        assert f.value == fo


def test_special_functions(hostcheck, oracle):
    L = oracle.lib()
    for a in (0.5, 1.3, 7.0, 25.5, 120.0):
        for x in (1e-3, 0.4, 2.0, 9.0, 60.0, 300.0):
            assert rel(hostcheck.hc_gamma_q(a, x), L.orc_gamma_q(a, x)) <= 1e-11
    for x in (1e-8, 0.1, 3.84, 30.0, 700.0):
        assert rel(hostcheck.hc_chisq_q(x, 1.0), L.orc_chisq_q(x, 1.0)) <= 1e-13
    for f in (0.0, 1e-31, 1e-4, 0.01, 0.3, 0.5, 0.7, 1.0):
        for sq in (0, 1):
            assert rel(hostcheck.hc_beta_weight(f, 1.0, 25.0, sq), L.orc_skat_weight(f, 1.0, 25.0, sq)) <= 1e-13


@pytest.mark.parametrize("n", [1, 2, 3, 7, 16, 31, 50, 64])
def test_parallel_jacobi(hostcheck, n):
    rng = np.random.default_rng(n)
    for rank in (n + 2, max(1, n // 3)):
        B = rng.standard_normal((n, rank)) * 10 ** rng.uniform(-2, 2, (n, 1))
        A = B @ B.T
        out = np.zeros(n)
        hostcheck.hc_eigen(_p(np.ascontiguousarray(A)), n, _p(out))
        ref = np.linalg.eigvalsh(A)[::-1]
        assert np.max(np.abs(out - ref)) <= 1e-12 * ref.max()
        assert np.all(np.diff(out) <= 0)
        out2 = np.zeros(n)
        hostcheck.hc_eigen_tridiag(_p(np.ascontiguousarray(A)), n, _p(out2))
        assert np.max(np.abs(out2 - ref)) <= 1e-13 * ref.max(), (n, rank)
        assert np.all(np.diff(out2) <= 0)


def test_tridiag_special_matrices(hostcheck):
    """diagonal, already-tridiagonal, repeated and zero eigenvalues"""
    for A in (np.diag([3.0, 1.0, 2.0, 0.0]), np.diag([5.0] * 6), np.zeros((5, 5)),
              np.diag([2.0] * 5) + np.diag([1.0] * 4, 1) + np.diag([1.0] * 4, -1),
              np.ones((7, 7)), np.array([[2.0, -1.0], [-1.0, 2.0]])):
        n = A.shape[0]
        out = np.zeros(n)
        hostcheck.hc_eigen_tridiag(_p(np.ascontiguousarray(A)), n, _p(out))
        ref = np.linalg.eigvalsh(A)[::-1]
        assert np.max(np.abs(out - ref)) <= 1e-13 * max(1.0, np.abs(ref).max()), A


def test_chisq_quantile_vs_gsl(hostcheck, oracle):
    G = oracle.ref_gsl()
    if G is None:
        pytest.skip("oracle/_ref/libgsl_ref.so not built")
    for q in (0.999, 0.9, 0.5, 0.1, 1e-3, 1e-8, 1e-15):
        for df in (0.3, 1.0, 2.5, 17.3, 120.0):
            assert rel(hostcheck.hc_chisq_qinv(q, df), G.ref_gsl_cdf_chisq_Qinv(q, df)) <= 1e-10


def test_qags_machine_matches_gsl_bitwise(hostcheck, oracle):
    """the resumable QAGS machine must reproduce gsl_integration_qags (GSL 1.16) node for node"""
    import math
    G = oracle.ref_gsl()
    if G is None:
        pytest.skip("oracle/_ref/libgsl_ref.so not built")
    CB = C.CFUNCTYPE(C.c_double, C.c_double)
    fs = [lambda x: math.exp(-x) * math.sqrt(x), lambda x: 1 / math.sqrt(x) if x > 0 else 0.0,
          lambda x: math.log(x) * math.sin(5 * x) if x > 0 else 0.0,
          lambda x: math.exp(-0.5 * x) / math.sqrt(2 * math.pi * x), lambda x: 1.0 / (1e-4 + (x - 7.3) ** 2),
          lambda x: abs(x - 13.1) ** 0.3]
    for f in fs:
        r, e, n = C.c_double(), C.c_double(), C.c_int()
        st = hostcheck.hc_qags(CB(f), 0.0, 40.0, 1e-25, 0.0001220703, C.byref(r), C.byref(e), C.byref(n))
        r2, e2, n2 = C.c_double(), C.c_double(), C.c_int()
        st2 = G.ref_gsl_qags(oracle._QAGS_CB(lambda x, _: f(x)), None, 0.0, 40.0, 1e-25, 0.0001220703, 1000,
                             C.byref(r2), C.byref(e2), C.byref(n2))
        assert (st != 0) == (st2 != 0)
        assert n.value == n2.value
        assert r.value == r2.value and e.value == e2.value


@pytest.mark.parametrize("case", [(1, 500, 5, 1), (2, 2000, 20, 3), (3, 3000, 50, 3), (4, 800, 1, 2), (5, 1500, 2, 3)])
def test_skato_tail_vs_oracle(hostcheck, oracle, case):
    """device SKAT-O tail (host build) on the M x M statistics vs the literal N x M numpy oracle"""
    import sys
    sys.path.insert(0, os.path.dirname(__file__))
    from util import af_of, make_problem
    from oracle import skato_oracle as SO
    O = oracle
    seed, N, M, Cc = case
    Gm, X, y = make_problem(O, seed, N, M, Cc, maf=np.linspace(0.01, 0.3, M))
    nm = O.fit_null_linear(X, y)
    ref = SO.skato_gene(Gm.astype(float), af_of(Gm), X, nm["resid"])
    keep = [j for j in range(M) if Gm[:, j].min() != Gm[:, j].max()]
    w = np.array([O.lib().orc_skat_weight(float(a), 1.0, 25.0, 0) for a in af_of(Gm)[: len(keep)]])
    Gw = Gm[:, keep].astype(float) * w[None, :]
    Wm = np.ascontiguousarray((Gw.T @ Gw - (Gw.T @ X) @ np.linalg.solve(X.T @ X, X.T @ Gw)) / 2)
    v = np.ascontiguousarray(nm["resid"] @ Gw)
    s2 = float(nm["resid"] @ nm["resid"]) / (N - 1)
    out = np.zeros(4)
    hostcheck.hc_skato_tail(_p(Wm), len(keep), _p(v), s2, _p(out))
    assert out[3] == 1.0 and ref["ok"]
    assert rel(out[0], ref["Q"]) <= 1e-10
    assert out[1] == ref["rho"]
    assert rel(out[2], ref["pvalue"]) <= 1e-8


@pytest.mark.parametrize("case", [(21, 800, 10, 3), (22, 1500, 25, 2), (23, 600, 1, 2), (24, 900, 2, 3), (25, 2500, 50, 3)])
def test_skato_tail_binary_vs_oracle(hostcheck, oracle, case):
    """binary trait (SkatO::Fit type "D"): the device tail with s2 = 1 on the p(1-p)-weighted M x M statistics the fp64
    path hands it (K / 2 = W (G'VG - G'VX (X'VX)^-1 X'VG) W / 2, v = W G'(y - p)) vs the literal N x M oracle, itself
    pinned on the reference's SkatO.cpp (tests/test_oracle_pin_reference_skat.py::test_live_binary_skato)"""
    import sys
    sys.path.insert(0, os.path.dirname(__file__))
    from util import af_of, make_problem
    from oracle import binary_oracle as BIN
    from oracle import skato_oracle as SO
    O = oracle
    seed, N, M, Cc = case
    Gm, X, _ = make_problem(O, seed, N, M, Cc, maf=np.linspace(0.01, 0.3, M))
    rng = np.random.default_rng(seed)
    y = (rng.random(N) < 1 / (1 + np.exp(0.3 - 0.5 * X[:, -1]))).astype(float)
    nm = BIN.fit_null_logistic(X, y)
    ref = SO.skato_gene(Gm.astype(float), af_of(Gm), X, nm["resid"], vv=nm["v"])
    keep = [j for j in range(M) if Gm[:, j].min() != Gm[:, j].max()]
    w = np.array([O.lib().orc_skat_weight(float(a), 1.0, 25.0, 0) for a in af_of(Gm)[: len(keep)]])
    Gw = Gm[:, keep].astype(float) * w[None, :]
    VG, VX = nm["v"][:, None] * Gw, nm["v"][:, None] * X
    Wm = np.ascontiguousarray((Gw.T @ VG - (Gw.T @ VX) @ np.linalg.solve(X.T @ VX, X.T @ VG)) / 2)
    v = np.ascontiguousarray(nm["resid"] @ Gw)
    out = np.zeros(4)
    hostcheck.hc_skato_tail(_p(Wm), len(keep), _p(v), 1.0, _p(out))
    assert out[3] == 1.0 and ref["ok"]
    assert rel(out[0], ref["Q"]) <= 1e-10
    assert out[1] == ref["rho"]
    assert rel(out[2], ref["pvalue"]) <= 1e-8


def test_tridiag_hard_spectra(hostcheck):
    """division-free Sturm counts (rescaled polynomial recurrence): graded, clustered, glued and
    rank-deficient spectra, the shapes the SKAT kernel matrix W^1/2 (G'PG) W^1/2 produces"""
    rng = np.random.default_rng(11)
    mats = []
    n = 21                                                    # Wilkinson W21+: pairs of close eigenvalues
    mats.append(np.diag(np.abs(np.arange(n) - 10.0)) + np.diag(np.ones(n - 1), 1) + np.diag(np.ones(n - 1), -1))
    for n in (8, 33, 50, 64):
        Qm, _ = np.linalg.qr(rng.standard_normal((n, n)))
        mats.append((Qm * 10.0 ** np.linspace(0, -18, n)) @ Qm.T)          # graded over 18 decades
        lam = np.concatenate([np.full(n // 2, 1.0) + 1e-13 * rng.standard_normal(n // 2), np.full(n - n // 2, 1e-9)])
        mats.append((Qm * lam) @ Qm.T)                                       # two tight clusters
        B = rng.standard_normal((n, 3))
        mats.append(B @ B.T * 1e12)                                          # rank 3, huge scale
        mats.append(B @ B.T * 1e-12)                                         # rank 3, tiny scale
        D = np.diag(rng.uniform(0.5, 2, n))                                  # glued blocks, 1e-160 coupling
        D[n // 2, n // 2 - 1] = D[n // 2 - 1, n // 2] = 1e-160
        mats.append(D)
    for A in mats:
        A = 0.5 * (A + A.T)
        n = A.shape[0]
        out = np.zeros(n)
        hostcheck.hc_eigen_tridiag(_p(np.ascontiguousarray(A)), n, _p(out))
        ref = np.linalg.eigvalsh(A)[::-1]
        assert np.max(np.abs(out - ref)) <= 2e-13 * np.abs(ref).max(), (n, np.max(np.abs(out - ref)) / np.abs(ref).max())
        assert np.all(np.diff(out) <= 0)
