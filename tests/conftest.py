import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def hostcheck():
    """g++ build of the product's __host__ __device__ math headers (logic check on the CPU)."""
    import ctypes as C
    import subprocess
    d = os.path.join(ROOT, "tests", "hostcheck")
    so = os.path.join(d, "libhostcheck.so")
    src = os.path.join(d, "hostcheck.cpp")
    deps = [src] + [os.path.join(ROOT, "rvtests_b200", "csrc", f)
                    for f in os.listdir(os.path.join(ROOT, "rvtests_b200", "csrc")) if f.endswith(".cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(p) > os.path.getmtime(so) for p in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", src, "-o", so], check=True)
    H = C.CDLL(so)
    dp = C.POINTER(C.c_double)
    H.hc_gamma_q.restype = C.c_double
    H.hc_gamma_q.argtypes = [C.c_double, C.c_double]
    H.hc_chisq_q.restype = C.c_double
    H.hc_chisq_q.argtypes = [C.c_double, C.c_double]
    H.hc_beta_weight.restype = C.c_double
    H.hc_beta_weight.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int]
    H.hc_liu.restype = C.c_double
    H.hc_liu.argtypes = [dp, C.c_int, C.c_double]
    H.hc_mixchisq.restype = C.c_double
    H.hc_mixchisq.argtypes = [dp, C.c_int, C.c_double, C.POINTER(C.c_int)]
    H.hc_qf.restype = C.c_double
    H.hc_qf.argtypes = [dp, C.c_int, C.c_double, C.c_int, C.c_double, C.POINTER(C.c_int)]
    H.hc_eigen.restype = C.c_int
    H.hc_eigen.argtypes = [dp, C.c_int, dp]
    H.hc_chisq_qinv.restype = C.c_double
    H.hc_chisq_qinv.argtypes = [C.c_double, C.c_double]
    H.hc_qags.restype = C.c_int
    H.hc_qags.argtypes = [C.CFUNCTYPE(C.c_double, C.c_double), C.c_double, C.c_double, C.c_double, C.c_double, dp, dp,
                          C.POINTER(C.c_int)]
    H.hc_skato_tail.restype = C.c_int
    H.hc_skato_tail.argtypes = [dp, C.c_int, dp, C.c_double, dp]
    H.hc_lfg_draws.restype = None
    H.hc_lfg_draws.argtypes = [C.c_uint, C.c_ulonglong, C.c_int, C.c_void_p]
    H.hc_fy_roots.restype = None
    H.hc_fy_roots.argtypes = [C.c_void_p, C.c_uint, C.c_void_p]
    H.hc_eigen_tridiag.restype = C.c_int
    H.hc_eigen_tridiag.argtypes = [dp, C.c_int, dp]
    return H


@pytest.fixture(scope="session")
def engine_cls():
    import rvtests_b200
    return rvtests_b200.GeneEngine
