"""Host side of the synthetic cohort: per-variant keys and integer thresholds for
the device generator (csrc/prep.cuh k_synth_rows), covariates and the null quantitative trait.
Pure parameter generation -- genotypes themselves are produced on the GPU.  (oracle/oracle.py holds
an independent twin used by the tests to cross-check the device stream.)"""
from __future__ import annotations

import numpy as np


def _mix64(z):
    z = np.asarray(z, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z


def variant_params(seed: int, vid0: int, n: int, maf_lo: float = 1e-4, maf_hi: float = 0.05):
    """keys (uint64), t0, t1 (uint32) for variants vid0 .. vid0+n-1; MAF log-uniform[lo, hi], HWE."""
    vid = np.arange(vid0, vid0 + n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        keys = _mix64(np.uint64(seed) + vid * np.uint64(0x9E3779B97F4A7C15))
    k2 = _mix64(keys ^ np.uint64(0xA5A5A5A5A5A5A5A5))
    u = (k2 >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    maf = maf_lo * (maf_hi / maf_lo) ** u
    q0 = (1.0 - maf) ** 2
    q01 = q0 + 2.0 * maf * (1.0 - maf)
    t0 = np.minimum(np.floor(q0 * 4294967296.0), 4294967295.0).astype(np.uint64).astype(np.uint32)
    t1 = np.minimum(np.floor(q01 * 4294967296.0), 4294967295.0).astype(np.uint64).astype(np.uint32)
    return keys, t0, t1


def covariates(seed: int, N: int, C: int = 3):
    """intercept + (C-1) N(0,1) covariates; y = 0.5 x1 - 0.3 x2 + N(0,1) (null for G)."""
    rng = np.random.Generator(np.random.Philox(key=int(seed)))
    X = np.ones((N, C))
    if C > 1:
        X[:, 1:] = rng.standard_normal((N, C - 1))
    y = rng.standard_normal(N)
    if C > 1:
        y = y + 0.5 * X[:, 1]
    if C > 2:
        y = y - 0.3 * X[:, 2]
    return X, y


def pack_bed(Gt, missing=None):
    """Hard calls (M, N) in {0,1,2} -> PLINK .bed SNP-major rows (M, ceil(N/4)) uint8, the coding the
    reference reads back in libVcf/PlinkInputFile.cpp:23-47: 0 -> 00, 1 -> 10, 2 -> 11, missing -> 01;
    sample p occupies bits 2(p&3)..2(p&3)+1 of byte p>>2.  missing: optional boolean (M, N) mask."""
    Gt = np.asarray(Gt)
    M, N = Gt.shape
    code = np.array([0, 2, 3], dtype=np.uint8)[Gt.astype(np.int64)]
    if missing is not None:
        code = np.where(missing, np.uint8(1), code)
    pad = (-N) % 4
    if pad:
        code = np.concatenate([code, np.zeros((M, pad), dtype=np.uint8)], axis=1)
    c = code.reshape(M, -1, 4)
    return np.ascontiguousarray((c[:, :, 0] | (c[:, :, 1] << 2) | (c[:, :, 2] << 4) | (c[:, :, 3] << 6)).astype(np.uint8))
