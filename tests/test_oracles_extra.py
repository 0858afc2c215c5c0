"""CPU: the numpy restatements added for the mixed-model and binary-trait rows are themselves checked against independent
formulations (the FastLMM and binary-trait ones are additionally pinned on the reference build in
test_oracle_pin_reference_skat.py; the Bolt null fit stays "parity unpinned": BoltLMM.cpp cannot be built here)."""
import numpy as np
import pytest


def test_bolt_random_is_mt19937_12345():
    """libsrc/Random.cpp seeds MT19937 with init_genrand(12345) (regression/BoltPlinkLoader.h:21) and maps a 32-bit output y
    to (y + 0.5) / 2^32; numpy's legacy RandomState uses the same seeding and the same generator"""
    from oracle import bolt_oracle as BO
    r = BO.Random(12345)
    raw = np.random.RandomState(12345).randint(0, 2**32, size=2000, dtype=np.uint64)   # crosses a 624-word refill
    got = np.array([r.next() for _ in range(2000)])
    assert np.array_equal(got, (raw.astype(np.float64) + 0.5) / 4294967296.0)
    # polar Box-Muller: second deviate first, the saved one next (Random.cpp:269-288)
    r2, r3 = BO.Random(7), BO.Random(7)
    while True:
        v1, v2 = 2 * r3.next() - 1, 2 * r3.next() - 1
        rsq = v1 * v1 + v2 * v2
        if 0 < rsq < 1:
            break
    fac = np.sqrt(-2 * np.log(rsq) / rsq)
    assert r2.normal() == v2 * fac and r2.normal() == v1 * fac


def test_bolt_oracle_recovers_heritability_and_solves_H():
    from oracle import bolt_oracle as BO
    rng = np.random.default_rng(5)
    N, M, C, h2 = 600, 400, 2, 0.5
    maf = rng.uniform(0.1, 0.5, M)
    G = rng.binomial(2, maf[:, None], size=(M, N)).astype(np.int8)
    covar = np.column_stack([np.ones(N), rng.normal(size=N)])
    X, Z, _ = BO.prepare(G, covar, np.zeros(N))
    assert np.allclose(Z.T @ Z, np.eye(C), atol=1e-12)
    assert np.allclose(X.mean(axis=0), 0, atol=1e-12) and np.allclose((X * X).sum(axis=0) / N / (1 - 0), (X * X).mean(axis=0))
    y = X @ rng.normal(size=M) * np.sqrt(h2 / M) + rng.normal(size=N) * np.sqrt(1 - h2) + covar @ np.array([2.0, 0.5])
    X, Z, yc = BO.prepare(G, covar, y)
    fit = BO.Fit(X, Z, yc)
    # the CG solve against a dense solve of the projected system: (P X X' P / M + delta I) x = P y
    delta = 1.3
    P = np.eye(N) - Z @ Z.T
    yv = (P @ yc)[:, None]
    x = fit.solve(yv, delta)
    dense = np.linalg.solve(P @ X @ X.T @ P / M + delta * np.eye(N), yv)
    assert np.max(np.abs(P @ x - dense)) <= 2e-3 * np.max(np.abs(dense))     # BOLT's tolerance is 5e-4 on |r|^2
    fit.fit().calibrate()
    assert 0.2 < fit.h2 < 0.8 and fit.sigma2_g > 0 and len(fit.log_delta) <= 7
    assert np.isfinite(fit.calibration) and fit.calibration > 0


def test_lmm_oracle_equals_the_sample_space_form():
    """FastLMM score branch in rotated coordinates == g_c' H^-1 r / sigma2 and g_c' (H^-1 - H^-1 X (X'H^-1X)^-1 X'H^-1) g_c / sigma2"""
    from oracle import lmm_oracle as LO
    rng = np.random.default_rng(0)
    N = 300
    Zm = rng.normal(size=(N, 500))
    K = Zm @ Zm.T / 500
    lam, U = np.linalg.eigh(K)
    X = np.c_[np.ones(N), rng.normal(size=N)]
    y = rng.normal(size=N)
    delta = 0.5
    nm = LO.fit_null_given_delta(U, lam, X, y, delta)
    g = rng.binomial(2, 0.2, size=N).astype(float)
    Us, Vs, st, p = LO.score(U, nm, g)
    Hinv = np.linalg.inv(K + delta * np.eye(N))
    gc = g - g.mean()
    Pm = Hinv - Hinv @ X @ np.linalg.inv(X.T @ Hinv @ X) @ X.T @ Hinv
    assert Us == pytest.approx(gc @ Hinv @ (y - X @ nm["beta"]) / nm["sigma2"], rel=1e-9)
    assert Vs == pytest.approx(gc @ Pm @ gc / nm["sigma2"], rel=1e-9)
    assert st == pytest.approx(Us * Us / Vs) and 0 <= p <= 1


def test_binary_oracle_logistic_fit_is_the_mle_up_to_its_stopping_rule(oracle):
    """the Newton loop of LogisticRegression::FitLogisticModel stops on a deviance change < 1e-3: its beta is within a Newton
    step of the maximum-likelihood estimate, and p / V are those of the PREVIOUS beta"""
    from oracle import binary_oracle as BIN
    from scipy import optimize
    rng = np.random.default_rng(3)
    N = 3000
    X = np.c_[np.ones(N), rng.normal(size=(N, 2))]
    y = (rng.random(N) < 1 / (1 + np.exp(-(X @ np.array([-0.5, 0.8, -0.4]))))).astype(float)
    nm = BIN.fit_null_logistic(X, y)

    def nll(b):
        eta = X @ b
        return float(np.sum(np.logaddexp(0, eta) - y * eta))

    mle = optimize.minimize(nll, np.zeros(3), method="BFGS", options=dict(gtol=1e-10)).x
    assert np.max(np.abs(nm["beta"] - mle)) <= 1e-5
    p_new = 1 / (1 + np.exp(-(X @ nm["beta"])))
    assert np.max(np.abs(nm["p"] - p_new)) > 0                       # one step stale ...
    assert np.max(np.abs(nm["p"] - p_new)) <= 1e-3                   # ... by a converged Newton step
    assert np.allclose(nm["v"], nm["p"] * (1 - nm["p"])) and np.allclose(nm["resid"], y - nm["p"])
    # a gene: the SKAT matrix is the projection form  W^1/2 G' (V - V X (X'VX)^-1 X'V) G W^1/2   (Skat.cpp:55-76)
    G = rng.binomial(2, 0.03, size=(N, 6)).astype(float)
    af = 0.5 * G.mean(axis=0)
    out = BIN.gene(G, af, X, nm)
    V = np.diag(nm["v"])
    P0 = V - V @ X @ np.linalg.inv(X.T @ V @ X) @ X.T @ V
    from scipy import stats
    w = stats.beta.pdf(af, 1, 25) ** 2
    Kfull = np.sqrt(w)[:, None] * (G.T @ P0 @ G) * np.sqrt(w)[None, :]
    assert np.allclose(np.sort(np.linalg.eigvalsh(Kfull))[::-1][:len(out["lam"])], out["lam"], rtol=1e-9)
