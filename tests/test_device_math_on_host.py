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
