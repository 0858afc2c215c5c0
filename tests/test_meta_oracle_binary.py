"""CPU: the numpy restatement of MetaUnrelatedBinary / MetaCovUnrelatedBinary (oracle/meta_oracle.py) pinned on the reference's
own LogisticRegressionScoreTest.cpp compiled in oracle/_ref/libskat_ref.so (intercept-only null models: with covariates the
reference's Matrix overload solves a 1 x 1 matrix against a d x d identity, LogisticRegressionScoreTest.cpp:292-295)."""
import numpy as np
import pytest

from oracle import binary_oracle as BIN
from oracle import meta_oracle as MO
from oracle import oracle as O


@pytest.mark.parametrize("seed", [5, 6])
def test_binary_meta_score_matches_the_reference_score_test(seed):
    try:
        O.ref_skat()
    except Exception as e:  # pragma: no cover
        pytest.skip(f"reference build unavailable: {e}")
    rng = np.random.default_rng(seed)
    N = 1500
    X = np.ones((N, 1))
    y = (rng.uniform(size=N) < 0.35).astype(np.float64)
    nm = BIN.fit_null_logistic(X, y)
    for maf in (0.02, 0.3):
        g = rng.binomial(2, maf, N).astype(np.float64)
        ref = O.ref_logistic_score_test(X, y, g)
        got = MO.meta_score_binary(g, y, X, nm)
        assert ref["rc"] == 0 and got["ok"]
        assert abs(got["U"] - ref["U"]) <= 1e-9 * max(abs(ref["U"]), 1.0)
        assert abs(got["sqrtV"] ** 2 - ref["V"]) <= 1e-9 * ref["V"]
        assert abs(got["pvalue"] - ref["pvalue"]) <= 1e-9 * ref["pvalue"]
        n1 = int(y.sum())
        assert got["cc"]["case"]["n"] == n1 and got["cc"]["ctrl"]["n"] == N - n1
        c = O.ref_genotype_counter(g[y == 1])
        assert (c["n_ref"], c["n_het"], c["n_alt"]) == tuple(got["cc"]["case"][k] for k in ("n_ref", "n_het", "n_alt"))
        assert abs(c["hwe_p"] - got["cc"]["case"]["hwe_p"]) <= 1e-9


def test_binary_meta_cov_restatement_is_the_projected_weighted_gram():
    rng = np.random.default_rng(3)
    N, nv = 800, 12
    X = np.column_stack([np.ones(N), rng.normal(size=N)])
    y = (rng.uniform(size=N) < 0.4).astype(np.float64)
    nm = BIN.fit_null_logistic(X, y)
    G = rng.binomial(2, 0.2, (N, nv)).astype(np.float64)
    pos = np.arange(nv) * 10
    out = MO.meta_cov_binary(G, pos, np.ones(nv, int), X, nm, 1000)
    W = np.diag(nm["v"])
    P = W - W @ X @ np.linalg.inv(X.T @ W @ X) @ X.T @ W
    full = G.T @ P @ G / N
    for i in range(nv):
        ps, vals = out[i]
        assert np.allclose(vals, full[i, i:], rtol=1e-10, atol=1e-14)
