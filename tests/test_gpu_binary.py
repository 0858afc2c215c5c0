"""GPU (-m gpu): binary traits: logistic null model on the device + the variance-weighted SKAT / CMC /
Zeggini statistics (src/Model.h:2673-2681, regression/LogisticRegression.cpp:279-339, LogisticRegressionScoreTest.cpp:219-302)
against the numpy restatement."""
import numpy as np
import pytest

from util import af_of, make_problem, rel

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", [(110, 900, 12, 1, 0.3), (111, 4000, 40, 3, 0.05), (112, 2500, 64, 2, 0.1)])
def test_binary_trait_vs_oracle(engine_cls, oracle, case):
    from oracle import binary_oracle as BIN
    from rvtests_b200.synth import pack_bed
    O = oracle
    seed, N, M, C, hi = case
    G, X, _ = make_problem(O, seed, N, M, C, maf=np.linspace(0.004, hi, M), n_flip=2, n_mono=1 if M > 12 else 0)
    rng = np.random.default_rng(seed)
    eta = -0.8 + (X[:, 1:] @ np.full(C - 1, 0.5) if C > 1 else 0.0)
    y = (rng.random(N) < 1.0 / (1.0 + np.exp(-eta))).astype(np.float64)
    nm = BIN.fit_null_logistic(X, y)
    eng = engine_cls(0)
    eng.set_null_model(X, y, binary=True)
    got = eng.get_null_model()
    assert np.max(np.abs(got["resid"] - nm["resid"])) <= 1e-10           # y - p of the SAME Newton round
    assert got["sigma2"] == 1.0
    assert np.max(np.abs(got["xtx_inv"] - nm["covB"])) <= 1e-9 * np.max(np.abs(nm["covB"]))
    af = af_of(G)
    eng.push_i8(G.T.copy(), af)
    eng.push_f64(G.astype(float), af)
    eng.push_bed(pack_bed(G.T), af)
    res = eng.flush()
    ref = BIN.gene(G.astype(float), af, X, nm)
    for k in range(3):
        r = res[k]
        assert int(r["m_poly"]) == ref["m_poly"] and int(r["status"]) == 0
        assert rel(r["Q"], ref["Q"]) <= 1e-6
        assert rel(r["lambda_max"], ref["lam"][0]) <= 1e-7
        assert int(r["davies_fault"]) == ref["fault"]
        assert rel(r["p_skat"], ref["p_skat"]) <= 1e-4
        for pre in ("cmc", "zeg"):
            b = ref[pre]
            assert abs(r[pre + "_U"] - b["U"]) <= 1e-6 * max(abs(b["U"]), np.sqrt(b["V"]))
            assert rel(r[pre + "_V"], b["V"]) <= 1e-6
            assert rel(r[pre + "_p"], b["p"]) <= 1e-4
        assert int(r["cmc_nonref"]) == ref["cmc"]["nonref"]
    # back to a quantitative trait on the same context: the linear path is unaffected
    Xq, yq = O.synth_covariates(seed, N, C)
    eng.set_null_model(Xq, yq)
    eng.push_i8(G.T.copy(), af)
    rq = eng.flush()[0]
    nmq = O.fit_null_linear(Xq, yq)
    refq, lamq = O.gene(G.astype(float), af, Xq, nmq["resid"], nmq["sigma2"])
    assert rel(rq["Q"], refq.skat.Q) <= 1e-6 and rel(rq["p_skat"], refq.skat.pvalue) <= 1e-4
    eng.close()


def test_binary_trait_rejects_other_codes(engine_cls):
    import rvtests_b200
    eng = engine_cls(0)
    X = np.ones((10, 1))
    with pytest.raises(rvtests_b200.RvtError):
        eng.set_null_model(X, np.arange(10.0), binary=True)      # case/control must already be 0/1 at this boundary
    eng.close()


def test_dense_genotypes_take_the_rescan_path(engine_cls, oracle):
    """common variants: most samples carry more non-zero calls than the per-sample list of k_tile_sparse holds, so the
    kernel re-reads their bytes; SKAT and Zeggini against the oracle (the CMC indicator is constant here: skipped)"""
    from oracle import binary_oracle as BIN
    from rvtests_b200.synth import pack_bed
    O = oracle
    seed, N, M, C = 113, 1300, 40, 2
    G, X, _ = make_problem(O, seed, N, M, C, maf=np.linspace(0.2, 0.45, M), n_flip=3)
    rng = np.random.default_rng(seed)
    y = (rng.random(N) < 1.0 / (1.0 + np.exp(0.3 - 0.5 * X[:, 1]))).astype(np.float64)
    nm = BIN.fit_null_logistic(X, y)
    assert np.median((G != 0).sum(axis=1)) > 14
    eng = engine_cls(0)
    eng.set_null_model(X, y, binary=True)
    af = af_of(G)
    eng.push_bed(pack_bed(G.T), af)
    r = eng.flush()[0]
    ref = BIN.gene(G.astype(float), af, X, nm)
    assert int(r["status"]) == 0 and int(r["m_poly"]) == ref["m_poly"]
    ctx = dict(Q=(r["Q"], ref["Q"]), lam=(r["lambda_max"], ref["lam"][0]), p=(r["p_skat"], ref["p_skat"]),
               zegU=(r["zeg_U"], ref["zeg"]["U"]), zegV=(r["zeg_V"], ref["zeg"]["V"]))
    assert rel(r["Q"], ref["Q"]) <= 1e-6, ctx
    assert rel(r["lambda_max"], ref["lam"][0]) <= 1e-7, ctx
    assert rel(r["p_skat"], ref["p_skat"]) <= 1e-4
    assert abs(r["zeg_U"] - ref["zeg"]["U"]) <= 1e-6 * max(abs(ref["zeg"]["U"]), np.sqrt(ref["zeg"]["V"]))
    assert rel(r["zeg_V"], ref["zeg"]["V"]) <= 1e-6 and rel(r["zeg_p"], ref["zeg"]["p"]) <= 1e-4
    assert int(r["cmc_nonref"]) == ref["cmc"]["nonref"]
    eng.close()


def test_binary_streaming_is_invisible_in_the_results(engine_cls, oracle):
    """options binary_stream / stream_batch: with a binary null model the fp64 statistics and the tail of the tile genes are
    enqueued behind their copies (batches of <= 8, a stream of their own) instead of at flush.  Same records in the same order
    -- a gene with missing calls (imputed on the fly), a mean-imputed Matrix push, a dosage gene (which keeps the flush path) and
    SKAT-O included -- whether streamed, streamed in small batches or all at flush.  (fp64 shared-memory atomics make the last
    bits of the sums schedule-dependent: 1e-10, not bitwise.)"""
    from oracle import binary_oracle as BIN
    from rvtests_b200.synth import pack_bed
    O = oracle
    N, C = 6001, 3
    rng = np.random.default_rng(191)
    X, _ = O.synth_covariates(191, N, C)
    y = (rng.random(N) < 1.0 / (1.0 + np.exp(0.7 - 0.4 * X[:, 1]))).astype(np.float64)
    Ms = [5, 50, 64, 1, 33, 62, 17, 40, 8, 21, 12]
    genes = []
    for g, M in enumerate(Ms):
        G, _, _ = make_problem(O, 1910 + g, N, M, C, maf=np.linspace(0.004, 0.06, M), n_flip=1 if M > 1 else 0)
        genes.append(G)
    mask = rng.random((Ms[4], N)) < 0.02
    raw = O.bed_decode_fast(pack_bed(genes[6].T, rng.random((Ms[6], N)) < 0.01), N).T
    G6 = O.impute_mean(raw)                                        # the Matrix form of a gene with missing calls
    G9 = genes[9].astype(np.float64)
    G9[11, 3] = 0.4                                                 # a real dosage

    def run(stream_batch, binary_stream):
        eng = engine_cls(0)
        try:
            eng.set_null_model(X, y, binary=True)
            eng.set_option("skato", 1)
            eng.set_option("stream_batch", stream_batch)
            eng.set_option("binary_stream", binary_stream)
            for g, G in enumerate(genes):
                if g == 4:
                    eng.push_bed(pack_bed(G.T, mask), af_of(G))
                elif g == 6:
                    eng.push_f64(G6, af_of(genes[6]))
                elif g == 9:
                    eng.push_f64(G9, af_of(G))
                elif g % 3 == 1:
                    eng.push_i8(G.T.copy(), af_of(G))
                else:
                    eng.push_bed(pack_bed(G.T), af_of(G))
            return eng.flush()
        finally:
            eng.close()

    base = run(0, 0)
    assert len(base) == len(Ms) and np.all(base["status"] == 0)
    nm = BIN.fit_null_logistic(X, y)
    ref = BIN.gene(genes[1].astype(float), af_of(genes[1]), X, nm)
    assert rel(base[1]["Q"], ref["Q"]) <= 1e-6
    for sb, bs in ((0, 1), (3, 1), (64, 1), (3, 0)):
        r = run(sb, bs)
        assert len(r) == len(Ms)
        for k in range(len(Ms)):
            for f in ("status", "m_poly", "cmc_nonref", "davies_fault", "skato_ok"):
                assert int(r[k][f]) == int(base[k][f]), (sb, bs, k, f)
            for f in ("Q", "p_skat", "cmc_p", "zeg_p", "skato_Q", "skato_p", "skato_rho"):
                assert rel(r[k][f], base[k][f]) <= 1e-9, (sb, bs, k, f, r[k][f], base[k][f])


@pytest.mark.parametrize("case", [(141, 3000, 100, 3, 0.0), (142, 2200, 180, 2, 0.01), (143, 900, 65, 1, 0.03)])
def test_binary_trait_wide_genes_vs_oracle(engine_cls, oracle, case):
    """A binary trait with a gene of more than 64 variants (the reference has no width limit: Skat::Fit sizes itself to the
    gene, regression/Skat.cpp:29-105): the p(1-p)-weighted statistics come from k_wide_sparse (fp64, from the int8 tiles of all
    T tiles, missing calls imputed on the fly), the tail is k_wide_finalize's.  SKAT, CMC, Zeggini and SKAT-O (type "D") against
    the numpy restatements; hard calls through all three host forms, missing calls through the 2-bit form and through the
    mean-imputed Matrix; an ordinary gene in the same flush."""
    from oracle import binary_oracle as BIN
    from oracle import skato_oracle as SO
    from rvtests_b200.synth import pack_bed
    O = oracle
    seed, N, M, C, miss = case
    G, X, _ = make_problem(O, seed, N, M, C, maf=np.linspace(0.003, 0.05, M), n_flip=2, n_mono=1)
    Gs, _, _ = make_problem(O, seed + 50, N, 20, C, maf=np.linspace(0.01, 0.1, 20))
    rng = np.random.default_rng(seed)
    eta = -0.7 + (X[:, 1:] @ np.full(C - 1, 0.4) if C > 1 else 0.0)
    y = (rng.random(N) < 1.0 / (1.0 + np.exp(-eta))).astype(np.float64)
    nm = BIN.fit_null_logistic(X, y)
    if miss > 0:
        mask = rng.random((M, N)) < miss
        mask[2] = False
        bed = pack_bed(G.T, mask)
        raw = O.bed_decode_fast(bed, N).T
        Gd = O.impute_mean(raw)
        af = 0.5 * np.where(raw >= 0, raw, 0.0).sum(axis=0) / N
    else:
        bed, Gd, af = pack_bed(G.T), G.astype(np.float64), af_of(G)
    eng = engine_cls(0)
    try:
        eng.set_null_model(X, y, binary=True)
        eng.set_option("skato", 1)
        eng.push_i8(Gs.T.copy(), af_of(Gs))
        eng.push_bed(bed, af)
        eng.push_f64(Gd, af)
        if miss == 0:
            eng.push_i8(G.T.copy(), af)
        res = eng.flush()
    finally:
        eng.close()
    assert np.all(res["status"] == 0)
    ref = BIN.gene(Gd, af, X, nm)
    so = SO.skato_gene(Gd, af, X, nm["resid"], vv=nm["v"])
    for r in res[1:]:
        assert int(r["m_poly"]) == ref["m_poly"]
        assert rel(r["Q"], ref["Q"]) <= 1e-6
        assert int(r["davies_fault"]) == ref["fault"]
        assert rel(r["p_skat"], ref["p_skat"]) <= 1e-4
        for pre in ("cmc", "zeg"):
            b = ref[pre]
            if b["V"] <= 1e-9 * N:     # with this many variants nearly every sample carries one: the CMC indicator is the intercept,
                assert abs(r[pre + "_V"]) <= 1e-9 * N                  # its variance rounding noise around zero on both sides
                continue
            assert abs(r[pre + "_U"] - b["U"]) <= 1e-6 * max(abs(b["U"]), np.sqrt(b["V"]))
            assert rel(r[pre + "_V"], b["V"]) <= 1e-6
            assert rel(r[pre + "_p"], b["p"]) <= 1e-4
        assert int(r["cmc_nonref"]) == ref["cmc"]["nonref"]
        assert int(r["skato_ok"]) == int(so["ok"]) == 1
        assert rel(r["skato_Q"], so["Q"]) <= 1e-6 and r["skato_rho"] == so["rho"]
        assert rel(r["skato_p"], so["pvalue"]) <= 1e-5
    refs = BIN.gene(Gs.astype(float), af_of(Gs), X, nm)
    assert rel(res[0]["Q"], refs["Q"]) <= 1e-6 and rel(res[0]["p_skat"], refs["p_skat"]) <= 1e-4
