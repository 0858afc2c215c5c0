"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the BoltLMM null-model fit.  parity unpinned:
the reference runs this in float32 on Eigen 3.3.9 (absent here); this restatement follows it step by step
in float64 and reproduces its random numbers bit for bit (libsrc/Random.cpp: MT19937 seeded with 12345, polar
Box-Muller), so the Monte-Carlo REML path is the reference's own path up to float rounding.

  Random                      libsrc/Random.cpp:128-140 (InitMersenne), :146-183 (Next), :269-288 (Normal);
                              seed 12345: regression/BoltPlinkLoader.h:21
  prepare                     BoltPlinkLoader::extractCovariateBasis / preparePhenotype / prepareGenotype
                              (regression/BoltPlinkLoader.cpp:115-264): orthonormal covariate basis Z, centred
                              phenotype, genotypes normalised to (g - 2p)/sqrt(2p(1-p)) with missing -> 0
  working_data                WorkingData::init (regression/BoltLMM.cpp:88-123)
  Hx, solve                   computeHx (:931-993), solve (:745-859): multi-RHS conjugate gradients, tolerance 5e-4 on
                              the PROJECTED squared residual norm, at most min(N, 250) iterations
  eval_reml, fit              evalREML (:669-724), EstimateHeritabilityBolt (:575-668), EstimateInfStatCalibration
                              (:1141-1214), MCtrial (:465)
Every (N+C)-row vector of the reference is [v ; Z'v]; "proj" products are v'w - (Z'v)'(Z'w) (:1064-1138).
"""
import numpy as np


class Random:
    def __init__(self, seed=12345):
        mt = np.zeros(624, dtype=np.uint64)
        mt[0] = seed & 0xFFFFFFFF
        for i in range(1, 624):
            mt[i] = (1812433253 * (int(mt[i - 1]) ^ (int(mt[i - 1]) >> 30)) + i) & 0xFFFFFFFF
        self.mt = [int(x) for x in mt]
        self.mti = 624
        self.saved = None

    def _refill(self):
        mt = self.mt
        for kk in range(624):
            y = (mt[kk] & 0x80000000) | (mt[(kk + 1) % 624] & 0x7FFFFFFF)
            mt[kk] = mt[(kk + 397) % 624] ^ (y >> 1) ^ (0x9908B0DF if y & 1 else 0)
        self.mti = 0

    def next(self):
        if self.mti >= 624:
            self._refill()
        y = self.mt[self.mti]
        self.mti += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9D2C5680
        y ^= (y << 15) & 0xEFC60000
        y ^= y >> 18
        return (1.0 / 4294967296.0) * (float(y) + 0.5)

    def normal(self):
        if self.saved is not None:
            v, self.saved = self.saved, None
            return v
        while True:
            v1 = 2.0 * self.next() - 1.0
            v2 = 2.0 * self.next() - 1.0
            rsq = v1 * v1 + v2 * v2
            if 0.0 < rsq < 1.0:
                break
        fac = np.sqrt(-2.0 * np.log(rsq) / rsq)
        self.saved = v1 * fac
        return v2 * fac


def prepare(G, covar, y, binary=False):
    """G: (M, N) int8 hard calls, -1 = missing; covar: (N, C) with the intercept in column 0; y: (N,).
    Returns X (N, M) normalised, Z (N, C') orthonormal, yc (centred phenotype; binary: as it is -- enableBinaryMode,
    BoltPlinkLoader.cpp:155-158)."""
    G = np.asarray(G)
    M, N = G.shape
    U, s, _ = np.linalg.svd(np.asarray(covar, dtype=np.float64), full_matrices=False)
    keep = 1 + int(np.sum(s[1:] > s[0] * 1e-8))
    Z = U[:, :keep]
    yc = np.asarray(y, dtype=np.float64) - (0.0 if binary else np.mean(y))
    X = np.zeros((N, M))
    for m in range(M):
        g = G[m].astype(np.float64)
        obs = g >= 0
        af = 0.5 * g[obs].sum() / obs.sum()
        sd = np.sqrt(2.0 * af * (1.0 - af))
        if sd > 0:
            X[obs, m] = (g[obs] - 2.0 * af) / sd
    return X, Z, yc


class Fit:
    def __init__(self, X, Z, yc, rng=None, mc_trials=None):
        self.X, self.Z = X, Z
        self.N, self.M = X.shape
        self.C = Z.shape[1]
        self.rng = rng or Random(12345)
        self.mc = mc_trials or max(min(int(4e9 / self.N / self.N), 15), 3)
        self.PX = X - Z @ (Z.T @ X)                       # never formed by the reference; same products
        self.cg_iters = []
        N, M, mc = self.N, self.M, self.mc
        beta = np.zeros((M, mc))
        for i in range(M):                                # WorkingData::init, row-major draw order
            for j in range(mc):
                beta[i, j] = self.rng.normal() / np.sqrt(float(M))
        self.beta_rand = beta
        self.x_beta = X @ beta                            # top rows; the bottom rows are Z' of it
        e = np.zeros((N, mc))
        for i in range(N):
            for j in range(mc):
                e[i, j] = self.rng.normal()
        self.e_rand = e
        self.yc = yc

    # projected products on top rows
    def pdot(self, a, b):
        return np.sum(a * b, axis=0) - np.sum((self.Z.T @ a) * (self.Z.T @ b), axis=0)

    def Hx(self, delta, v):
        Xy = self.X.T @ v - (self.X.T @ self.Z) @ (self.Z.T @ v)
        return self.X @ Xy / self.M + delta * v

    def solve(self, y, delta):
        x = y / delta
        r = y - self.Hx(delta, x)
        p = r.copy()
        rsold = self.pdot(r, r)
        tol = 5e-4
        it = 0
        for it in range(1, min(self.N, 250) + 1):
            ap = self.Hx(delta, p)
            with np.errstate(divide="ignore", invalid="ignore"):
                alpha = rsold / self.pdot(p, ap)
            alpha[~np.isfinite(alpha)] = 0.0
            x = x + p * alpha
            r = r - ap * alpha
            rsnew = self.pdot(r, r)
            if np.all(rsnew < tol):
                break
            if np.max(np.abs(rsnew - rsold)) < tol:
                break
            with np.errstate(divide="ignore", invalid="ignore"):
                ratio = rsnew / rsold
            ratio[(rsnew < tol) | ~np.isfinite(ratio)] = 0.0
            p = r + p * ratio
            rsold = rsnew
        self.cg_iters.append(it)
        return x

    def eval_reml(self, log_delta):
        delta = np.exp(log_delta)
        Y = np.column_stack([self.yc, self.x_beta + np.sqrt(delta) * self.e_rand])
        H = self.solve(Y, delta)
        self.H_inv_y = H
        PH = H - self.Z @ (self.Z.T @ H)
        beta_hat = self.X.T @ PH / self.M                 # estimateBetaAndE: g' H - (Z'g)' (Z'H)
        e_hat = delta * H
        en = self.pdot(e_hat, e_hat)
        bn = np.sum(beta_hat * beta_hat, axis=0)
        r_data = (bn[0], en[0])
        r_rand = (bn[1:].sum(), en[1:].sum())
        return float(np.log((r_data[0] / r_data[1]) / (r_rand[0] / r_rand[1])))

    def fit(self):
        h2 = [0.25]
        ld = [np.log((1 - h2[0]) / h2[0])]
        f = [self.eval_reml(ld[0])]
        h2.append(0.125 if f[0] < 0 else min(0.5, 0.5 * 0.25 + 0.5))
        ld.append(np.log((1 - h2[1]) / h2[1]))
        f.append(self.eval_reml(ld[1]))
        i = 2
        while i < 7:
            with np.errstate(divide="ignore", invalid="ignore"):
                nld = (ld[i - 2] * f[i - 1] - ld[i - 1] * f[i - 2]) / (f[i - 1] - f[i - 2])
            if not np.isfinite(nld):
                i -= 1
                break
            nld = min(max(nld, -10.0), 5.0)
            ld.append(nld)
            h2.append(1.0 / (1.0 + np.exp(nld)))
            if abs(ld[i] - ld[i - 1]) < 0.01:
                break
            f.append(self.eval_reml(ld[i]))
            i += 1
        if i == 7:
            i -= 1
        self.log_delta, self.f = ld, f
        return self.finish(float(np.exp(ld[i])), h2[i])

    def finish(self, delta, h2):
        """BoltLMM.cpp:649-666: the variance components from the LAST solve (which may be one secant step behind delta)"""
        self.delta = delta
        y0, H0 = self.yc[:, None], self.H_inv_y[:, :1]
        self.sigma2_g = float(self.pdot(y0, H0)[0] / (self.N - self.C))
        self.sigma2_e = self.delta * self.sigma2_g
        self.h2 = h2
        self.h = self.H_inv_y[:, 0] / self.sigma2_g      # H_inv_y_
        self.h_norm2 = float(self.pdot(self.h[:, None], self.h[:, None])[0])
        return self

    def calibrate(self):
        n_snp = min(30, self.M)
        idx = [int(self.rng.next() * self.M) for _ in range(n_snp)]
        g = self.X[:, idx]
        V = self.solve(g, self.sigma2_e / self.sigma2_g) / self.sigma2_g
        xVy = self.pdot(g, np.repeat(self.h[:, None], n_snp, axis=1))
        xVx = self.pdot(g, V)
        xx = self.pdot(g, g)
        prosp = xVy ** 2 / xVx
        retro = self.N * xVy ** 2 / (xx * self.h_norm2)
        sel = prosp < 5.0
        r0, r1 = retro[sel].sum(), prosp[sel].sum()
        self.calibration = float(r0 / r1) if r1 != 0 else 1.0
        ratio = xVx.sum() / xx.sum()
        self.xvx_xx_ratio = float(ratio) if np.isfinite(ratio) else 1.0
        self.calib_idx = idx
        return self
