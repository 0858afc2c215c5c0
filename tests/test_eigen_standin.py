"""CPU: oracle/eigen_standin -- the stand-in for the Eigen API that lets the reference's own sources compile
(oracle/_ref/libskat_ref.so) -- held against numpy, operation by operation.  The reference build pins the oracle;
this pins the stand-in."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
dp = C.POINTER(C.c_double)


def P(a):
    return a.ctypes.data_as(dp)


@pytest.fixture(scope="module")
def st():
    src = os.path.join(ROOT, "oracle", "eigen_standin_selftest.cpp")
    so = os.path.join(ROOT, "oracle", "libeigen_standin_selftest.so")
    inc = os.path.join(ROOT, "oracle", "eigen_standin")
    hdr = [os.path.join(inc, "third", "eigen", "Eigen", f) for f in os.listdir(os.path.join(inc, "third", "eigen", "Eigen"))]
    if not os.path.exists(so) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in [src] + hdr):
        subprocess.run(["g++", "-O2", "-std=c++11", "-fPIC", "-shared", "-I" + inc, src, "-o", so], check=True)
    return C.CDLL(so)


def F(a):
    return np.asfortranarray(a, dtype=np.float64)


def test_products_and_scalars(st):
    rng = np.random.default_rng(1)
    n, k, m = 37, 5, 4
    A, B, Cm, d = F(rng.normal(size=(n, k))), F(rng.normal(size=(n, m))), F(rng.normal(size=(k, m))), rng.normal(size=n)
    out = F(np.zeros((k, m)))
    st.st_algebra(n, k, m, P(A), P(d), P(B), P(Cm), C.c_double(3.0), P(out))
    assert np.allclose(out, A.T @ (d[:, None] * B) - Cm / 3.0, rtol=1e-13, atol=1e-13)


def test_solvers(st):
    rng = np.random.default_rng(2)
    for n, m in ((1, 1), (3, 3), (8, 2), (20, 5)):
        R = rng.normal(size=(n + 3, n))
        A, B = F(R.T @ R), F(rng.normal(size=(n, m)))
        inv, llt, ldlt, L = F(np.zeros((n, n))), F(np.zeros((n, m))), F(np.zeros((n, m))), F(np.zeros((n, n)))
        det, rank = C.c_double(0), C.c_int(0)
        st.st_solvers(n, m, P(A), P(B), P(inv), P(llt), P(ldlt), P(L), C.byref(det), C.byref(rank))
        assert np.allclose(inv, np.linalg.inv(A), rtol=1e-9, atol=1e-12)
        assert np.allclose(llt, np.linalg.solve(A, B), rtol=1e-9, atol=1e-12)
        assert np.allclose(ldlt, np.linalg.solve(A, B), rtol=1e-9, atol=1e-12)
        assert np.allclose(L, np.linalg.cholesky(A), rtol=1e-10, atol=1e-12)
        assert det.value == pytest.approx(np.linalg.det(A), rel=1e-9)
        assert rank.value == n
    # LDLT of a matrix with an exactly-zero row / column (Z'Z of centred covariates incl. the intercept): pseudo-inverse
    Z = rng.normal(size=(50, 3))
    Z -= Z.mean(0)
    Zc = np.c_[np.zeros(50), Z]
    A, B, x = F(Zc.T @ Zc), F(np.eye(4)), F(np.zeros((4, 4)))
    st.st_ldlt(4, 4, P(A), P(B), P(x))
    assert np.all(np.isfinite(x)) and np.allclose(x, np.linalg.pinv(A), rtol=1e-9, atol=1e-12)
    # and a matrix that needs the pivoting (leading zero on the diagonal, still non-singular after the permutation)
    A2 = F(np.array([[0.0, 0.0, 0.0], [0.0, 4.0, 1.0], [0.0, 1.0, 3.0]]))
    x2 = F(np.zeros((3, 3)))
    st.st_ldlt(3, 3, P(A2), P(F(np.eye(3))), P(x2))
    assert np.allclose(x2, np.linalg.pinv(A2), atol=1e-12)
    low = F(rng.normal(size=(9, 3)) @ rng.normal(size=(3, 7)))
    assert st.st_rank(9, 7, P(low)) == 3


def test_self_adjoint_eigen_solver(st):
    rng = np.random.default_rng(3)
    for n in (1, 2, 7, 30, 64):
        R = rng.normal(size=(n, n))
        A = F(R + R.T)
        if n == 30:  # a rank-deficient PSD matrix, as K of a gene with collinear variants
            Z = rng.normal(size=(n, 5))
            A = F(Z @ Z.T)
        v64, V, v32 = np.zeros(n), F(np.zeros((n, n))), np.zeros(n)
        st.st_eigen(n, P(A), P(v64), P(V), P(v32))
        ref = np.linalg.eigvalsh(A)
        scale = max(1.0, np.max(np.abs(ref)))
        assert np.all(np.diff(v64) >= 0), "increasing order is what the reference's back-to-front loops rely on"
        assert np.max(np.abs(v64 - ref)) <= 1e-12 * scale
        assert np.max(np.abs(v32 - ref)) <= 2e-5 * scale
        assert np.allclose(V @ np.diag(v64) @ V.T, A, atol=1e-11 * scale)
        assert np.allclose(V.T @ V, np.eye(n), atol=1e-12)


def test_reductions_blocks_comma_map(st):
    rng = np.random.default_rng(4)
    n, m = 6, 4
    A = F(rng.normal(size=(n, m)))
    rowsum, colmean, cen, arr = np.zeros(n), np.zeros(m), F(np.zeros((n, m))), F(np.zeros((n, m)))
    blocks, comma, mapped, vfr, sc = F(np.zeros((n, m))), F(np.zeros((n, 2 * m))), F(np.zeros((n, m))), np.zeros(m), np.zeros(8)
    st.st_misc(n, m, P(A), P(rowsum), P(colmean), P(cen), P(arr), P(blocks), P(comma), P(mapped), P(vfr), P(sc))
    assert np.allclose(rowsum, A.sum(1)) and np.allclose(colmean, A.mean(0))
    assert np.allclose(cen, A - A.mean(0))
    assert np.allclose(arr, (A * A - 1.0) ** 2 / 2.0)
    bl = np.zeros((n, m))
    bl[:, 0] = A[:, m - 1]
    bl[n - 1, :] = A[0, :]
    bl[np.arange(m), np.arange(m)] += A[:m, 0]
    assert np.allclose(blocks, bl)
    assert np.allclose(comma, np.c_[A, A - A.mean(0)])
    assert np.allclose(mapped, 2 * A)
    assert np.allclose(vfr, A.sum(0)) and (sc[4], sc[5]) == (m, 1)
    assert np.allclose(sc[:4], [A.sum(), (A * A).sum(), np.sqrt((A * A).sum()), np.trace(A)])
    assert (sc[6], sc[7]) == (A.min(), A.max())


def test_bolt_api_blocks_arrays_and_svd(st):
    """the part of the API regression/BoltLMM.cpp and BoltPlinkLoader.cpp add: blocks of blocks, block arrays that write
    through, the projected column products as the reference spells them, array comparisons, the thin SVD"""
    rng = np.random.default_rng(5)
    n, c, k = 41, 3, 4
    A, B = F(rng.normal(size=(n + c, k))), F(rng.normal(size=(n + c, k)))
    Z = F(np.column_stack([np.ones(n), rng.normal(size=(n, c - 1))]))
    pd, pn, cen, proj = np.zeros(k), np.zeros(k), np.zeros(n), F(np.zeros((n + c, k)))
    colhead, blkdiv, sv, U = np.zeros(n), np.zeros(n), np.zeros(c), F(np.zeros((n, c)))
    all_lt, any_lt = C.c_int(-1), C.c_int(-1)
    st.st_bolt_api(n, c, k, P(A), P(B), P(Z), P(pd), P(pn), P(cen), P(proj), P(colhead), C.byref(all_lt), C.byref(any_lt), P(sv), P(U), P(blkdiv))
    assert np.allclose(pd, (A[:n] * B[:n]).sum(0) - (A[n:] * B[n:]).sum(0), rtol=1e-13, atol=1e-13)
    assert np.allclose(pn, (A[:n] ** 2).sum(0) - (A[n:] ** 2).sum(0), rtol=1e-13, atol=1e-13)
    assert np.allclose(cen, A[:n, 0] - A[:n, 0].mean(), rtol=1e-13, atol=1e-13)
    want = B.copy()
    want[n:] = Z.T @ B[:n]
    assert np.allclose(proj, want, rtol=1e-12, atol=1e-12)
    assert np.allclose(colhead, B[:n, 0]) and np.allclose(blkdiv, A[:n, 0] / 4.0)
    assert all_lt.value == 1 and any_lt.value == 0
    s_np = np.linalg.svd(Z, compute_uv=False)
    assert np.allclose(sv, s_np, rtol=1e-12)
    assert np.allclose(U.T @ U, np.eye(c), atol=1e-12)                       # orthonormal
    assert np.allclose(U @ (U.T @ Z), Z, atol=1e-10)                         # spans the columns of Z
