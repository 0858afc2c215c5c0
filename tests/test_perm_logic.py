"""CPU: the host-testable core of the permutation test (rvtests_b200/csrc/permlogic.cuh, compiled with g++ by
tests/hostcheck) against glibc's own rand() and a literal Fisher-Yates shuffle (src/LinearAlgebra.h:8-21)."""
import numpy as np
import pytest


def test_lfg_jump_ahead_reproduces_glibc_rand(hostcheck, oracle):
    """rand() after srand(1) -- the state the reference runs with, it never seeds -- at arbitrary stream positions"""
    ref = oracle.glibc_rand(200000, reseed=1)
    assert ref[0] == 1804289383            # the well-known first value of glibc's rand()
    for pos, n in ((0, 1000), (1, 64), (30, 100), (31, 100), (12345, 5000), (199000, 1000)):
        out = np.zeros(n, dtype=np.int32)
        hostcheck.hc_lfg_draws(1, pos, n, out.ctypes.data)
        assert np.array_equal(out, ref[pos:pos + n]), pos
    # another seed, far position (the oracle skips 3e6 draws by calling rand())
    far = oracle.glibc_rand(500, reseed=20260925, skip=3_000_000)
    out = np.zeros(500, dtype=np.int32)
    hostcheck.hc_lfg_draws(20260925, 3_000_000, 500, out.ctypes.data)
    assert np.array_equal(out, far)


@pytest.mark.parametrize("n", [1, 2, 3, 17, 1000, 50001])
def test_fisher_yates_chains_equal_the_literal_shuffle(hostcheck, n):
    rng = np.random.default_rng(n)
    for rep in range(3):
        draws = rng.integers(0, 2**31, size=max(n - 1, 1), dtype=np.uint32)
        if rep == 2:
            draws[:] = 0                  # every step swaps with position 0: the longest possible chains
        v = np.arange(n)
        for s in range(n - 1):             # permute(): i = n-1..1, j = rand() % (i+1)
            i = n - 1 - s
            j = int(draws[s]) % (i + 1)
            v[i], v[j] = v[j], v[i]
        root = np.zeros(n, dtype=np.uint32)
        hostcheck.hc_fy_roots(draws.ctypes.data, n, root.ctypes.data)
        assert np.array_equal(root, v)
