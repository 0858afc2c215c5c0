"""GPU (-m gpu): A6, the permutation p-value of `--kernel skat` (src/Model.h:2707-2717), against the oracle's literal
loop (glibc rand() Fisher-Yates on the host, float32 statistic): the same shuffles, hence the same counts."""
import numpy as np
import pytest

from util import af_of, make_problem, rel

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng(engine_cls):
    e = engine_cls(0)
    if e.info("tc_available") != 1:
        pytest.skip("the permutation test needs the tensor-core sweep")
    yield e
    e.close()


def _genes(O, N, C, specs):
    X, y = None, None
    out = []
    for seed, M, kw in specs:
        G, X, y = make_problem(O, seed, N, M, C, **kw)
        out.append(G)
    return out, X, y


@pytest.mark.parametrize("case", [(80, 257, 1, 40, 0.3), (81, 3001, 3, 300, 0.05), (82, 20000, 3, 64, 0.5)])
def test_perm_counts_equal_the_reference_loop(eng, oracle, case):
    """three genes in a row (one of them wider than a tile, one with flipped + monomorphic variants): ActualPerm,
    NumGreater, NumEqual and the stream position of every gene equal the serial host loop that calls rand()."""
    O = oracle
    seed, N, C, n_perm, alpha = case
    genes, X, y = _genes(O, N, C, [(seed, 12, dict(maf=np.linspace(0.01, 0.2, 12), n_flip=2, n_mono=1)),
                                   (seed, 70, dict(maf=np.linspace(0.005, 0.1, 70), n_flip=1)),
                                   (seed, 5, dict(maf=0.1))])
    eng.set_option("engine", 0)
    eng.set_null_model(X, y)
    nm = O.fit_null_linear(X, y)
    eng.set_option("perm", n_perm)
    eng.set_option("perm_alpha", alpha)
    eng.set_option("perm_batch", 32)
    eng.set_option("perm_seed", 1)            # the state of a fresh process (the reference never calls srand)
    try:
        for G in genes:
            eng.push_i8(G.T.copy(), af_of(G))
        res = eng.flush()
        pr = eng.perm_results()
    finally:
        eng.set_option("perm", 0)
    assert len(pr) == len(genes)
    pos = 0
    for g, G in enumerate(genes):
        ref = O.gene_perm(G.astype(float), af_of(G), nm["resid"], float(res[g]["Q"]), n_perm=n_perm, alpha=alpha,
                          reseed=1 if g == 0 else 0)
        assert ref["rc"] == 0 and int(pr[g]["done"]) == 1
        assert int(pr[g]["num_perm"]) == n_perm
        assert int(pr[g]["stream_pos"]) == pos
        assert pr[g]["stat"] == res[g]["Q"]
        # the statistics of the very same shuffles: float32 in the reference, exact here
        assert int(pr[g]["actual_perm"]) == ref["actual"], (g, pr[g], ref["actual"])
        assert int(pr[g]["num_greater"]) == ref["greater"], (g, pr[g], ref["greater"])
        assert int(pr[g]["num_equal"]) == ref["equal"]
        assert rel(pr[g]["p_perm"], ref["p"]) <= 1e-12
        pos += ref["actual"] * (N - 1)
    assert eng.info("perm_stream_pos") == pos


def test_perm_binary_trait(engine_cls, oracle):
    """binary trait: the same loop on r = y - p of the logistic null (src/Model.h:2673-2717) -- counts and stream positions
    of two genes against the serial host loop"""
    from oracle import binary_oracle as BIN
    O = oracle
    N, C, n_perm, alpha = 2500, 2, 200, 0.1
    genes, X, _ = _genes(O, N, C, [(85, 12, dict(maf=np.linspace(0.01, 0.2, 12), n_flip=2, n_mono=1)), (85, 30, dict(maf=0.05))])
    rng = np.random.default_rng(85)
    y = (rng.random(N) < 1.0 / (1.0 + np.exp(0.4 - 0.5 * X[:, 1]))).astype(np.float64)
    nm = BIN.fit_null_logistic(X, y)
    e = engine_cls(0)
    try:
        e.set_option("perm", n_perm)
        e.set_option("perm_alpha", alpha)
        e.set_option("perm_batch", 32)
        e.set_option("perm_seed", 1)
        e.set_null_model(X, y, binary=True)
        for G in genes:
            e.push_i8(G.T.copy(), af_of(G))
        res = e.flush()
        pr = e.perm_results()
    finally:
        e.close()
    pos = 0
    for g, G in enumerate(genes):
        ref = O.gene_perm(G.astype(float), af_of(G), nm["resid"], float(res[g]["Q"]), n_perm=n_perm, alpha=alpha,
                          reseed=1 if g == 0 else 0)
        assert ref["rc"] == 0 and int(pr[g]["done"]) == 1, (g, pr[g])
        assert int(pr[g]["stream_pos"]) == pos
        assert (int(pr[g]["actual_perm"]), int(pr[g]["num_greater"]), int(pr[g]["num_equal"])) == (ref["actual"], ref["greater"], ref["equal"]), (g, pr[g], ref)
        pos += ref["actual"] * (N - 1)


def test_perm_genes_with_missing_calls(engine_cls, oracle):
    """a gene with missing calls (mean-imputed, augmented sweep) between two complete genes: its permutations run -- on the
    operand tiles H and M, s = H'r_pi + delta M'r_pi -- and consume the rand() stream as the reference's loop does, so the
    counts AND stream positions of the genes after it still replay the serial host loop."""
    from rvtests_b200.synth import pack_bed
    O = oracle
    N, C, n_perm, alpha = 3001, 3, 150, 0.2
    genes, X, y = _genes(O, N, C, [(87, 10, dict(maf=np.linspace(0.01, 0.2, 10))),
                                   (88, 30, dict(maf=np.linspace(0.005, 0.1, 30), n_flip=2, n_mono=1)),
                                   (89, 6, dict(maf=0.1))])
    nm = O.fit_null_linear(X, y)
    rng = np.random.default_rng(87)
    mask = rng.random((30, N)) < 0.02
    mask[3, : N // 4] = True
    bed1 = pack_bed(genes[1].T, mask)
    raw = O.bed_decode_fast(bed1, N).T
    Gd1 = O.impute_mean(raw)
    af1 = 0.5 * np.where(raw >= 0, raw, 0.0).sum(axis=0) / N
    mats = [genes[0].astype(float), Gd1, genes[2].astype(float)]
    afs = [af_of(genes[0]), af1, af_of(genes[2])]
    e = engine_cls(0)
    try:
        e.set_option("perm", n_perm)
        e.set_option("perm_alpha", alpha)
        e.set_option("perm_batch", 32)
        e.set_option("perm_seed", 1)
        e.set_null_model(X, y)
        e.push_bed(pack_bed(genes[0].T), afs[0])
        e.push_bed(bed1, af1)
        e.push_bed(pack_bed(genes[2].T), afs[2])
        res = e.flush()
        pr = e.perm_results()
        assert int(e.info("last_aug")) == 1
    finally:
        e.close()
    pos = 0
    for g in range(3):
        ref = O.gene_perm(mats[g], afs[g], nm["resid"], float(res[g]["Q"]), n_perm=n_perm, alpha=alpha, reseed=1 if g == 0 else 0)
        assert ref["rc"] == 0 and int(pr[g]["done"]) == 1, (g, pr[g])
        assert int(pr[g]["stream_pos"]) == pos, (g, pr[g], pos)
        assert (int(pr[g]["actual_perm"]), int(pr[g]["num_greater"]), int(pr[g]["num_equal"])) == (ref["actual"], ref["greater"], ref["equal"]), (g, pr[g], ref)
        pos += ref["actual"] * (N - 1)


def test_perm_statistics_match_per_shuffle(eng, oracle):
    """with alpha = 1 every permutation runs: compare each permuted statistic, not only the counts"""
    O = oracle
    N, C, n_perm = 5000, 2, 48
    G, X, y = make_problem(O, 83, N, 30, C, maf=np.linspace(0.004, 0.1, 30), n_flip=2)
    eng.set_option("engine", 0)
    eng.set_null_model(X, y)
    nm = O.fit_null_linear(X, y)
    eng.set_option("perm", n_perm)
    eng.set_option("perm_alpha", 1.0)
    eng.set_option("perm_batch", 16)
    eng.set_option("perm_seed", 1)
    eng.set_option("debug_perm_q", 1)
    try:
        eng.push_i8(G.T.copy(), af_of(G))
        res = eng.flush()
        pr = eng.perm_results()
        q = eng.perm_debug_q()
    finally:
        eng.set_option("perm", 0)
        eng.set_option("debug_perm_q", 0)
    ref = O.gene_perm(G.astype(float), af_of(G), nm["resid"], float(res[0]["Q"]), n_perm=n_perm, alpha=1.0, reseed=1)
    assert int(pr[0]["actual_perm"]) == ref["actual"] == n_perm
    assert len(q) == n_perm
    assert np.max(np.abs(q - ref["q"]) / ref["q"]) <= 2e-5     # the reference's statistic is float32


def test_perm_off_by_default_and_na_gene(eng, oracle):
    O = oracle
    G, X, y = make_problem(O, 84, 600, 4, 1, maf=0.2, n_mono=4)   # every variant monomorphic: fit() == -1
    eng.set_null_model(X, y)
    eng.push_i8(G.T.copy(), af_of(G))
    eng.flush()
    assert len(eng.perm_results()) == 0
    eng.set_option("perm", 100)
    try:
        eng.push_i8(G.T.copy(), af_of(G))
        eng.flush()
        pr = eng.perm_results()
    finally:
        eng.set_option("perm", 0)
    assert len(pr) == 1 and int(pr[0]["done"]) == 0 and int(pr[0]["actual_perm"]) == 0


def test_device_rand_equals_glibc_rand(eng, oracle):
    """k_lfg_draws (CTA + thread jump-ahead, state in registers) against glibc's own rand(): 60 M consecutive values
    from a fresh stream, and a window at a far position"""
    n = 60_000_000
    ref = oracle.glibc_rand(n, reseed=1)
    dev = eng.debug_rand(n, seed=1, pos=0)
    bad = np.flatnonzero(dev != ref)
    assert bad.size == 0, (bad[:5], dev[bad[:5]], ref[bad[:5]])
    far = oracle.glibc_rand(100_000, reseed=7, skip=5_000_000)
    assert np.array_equal(eng.debug_rand(100_000, seed=7, pos=5_000_000), far)


def test_perm_reports_lost_stream_parity_after_an_uncovered_gene(engine_cls, oracle):
    """A gene with dosages pushed as doubles is not permuted (done = 0) although the reference would have shuffled for it:
    every LATER record of the context says stream_ok = 0 (valid permutation statistics, but not the reference's own draws)
    until the stream position is set again; a gene the reference itself skips (fit() == -1) costs nothing."""
    O = oracle
    N = 500
    G1, X, y = make_problem(O, 91, N, 6, 1, maf=0.2)
    Gm, _, _ = make_problem(O, 92, N, 4, 1, maf=0.2, n_mono=4)
    G2, _, _ = make_problem(O, 93, N, 7, 1, maf=0.2)
    Gd = G2.astype(np.float64)
    Gd[::7, 0] = 0.37                                   # dosages: the fp64 path
    eng = engine_cls(0)
    try:
        eng.set_null_model(X, y)
        eng.set_option("perm", 200)
        eng.push_i8(G1.T.copy(), af_of(G1))
        eng.push_i8(Gm.T.copy(), af_of(Gm))             # NA in the reference too: no draws consumed
        eng.push_f64(Gd, af_of(G2))                     # testable in the reference, not covered here
        eng.push_i8(G2.T.copy(), af_of(G2))
        res = eng.flush()
        pr = eng.perm_results()
        assert [int(r["status"]) for r in res][0] == 0 and int(res[2]["status"]) == 0
        assert [int(p["done"]) for p in pr] == [1, 0, 0, 1]
        assert [int(p["stream_ok"]) for p in pr] == [1, 1, 1, 0]
        eng.push_i8(G1.T.copy(), af_of(G1))
        eng.flush()
        assert int(eng.perm_results()[0]["stream_ok"]) == 0          # sticky across flushes
        eng.set_option("perm_stream_pos", 0)
        eng.push_i8(G1.T.copy(), af_of(G1))
        eng.flush()
        p2 = eng.perm_results()[0]
        assert int(p2["stream_ok"]) == 1 and int(p2["num_greater"]) == int(pr[0]["num_greater"]) and int(p2["actual_perm"]) == int(pr[0]["actual_perm"])
    finally:
        eng.close()
