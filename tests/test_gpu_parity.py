"""GPU (-m gpu): the CUDA path, called through the C ABI, against the oracle on the same seeded
inputs.  Integer/count outputs bit exact; Q <= 1e-6 rel; p-values <= 1e-4 rel with the same
Davies fault flag (BASELINE.json north_star tolerances)."""
import json
import os

import numpy as np
import pytest

from util import af_of, check_gene, make_problem, rel

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
ENGINES = [1, 2]  # RVT_ENGINE_SIMT, RVT_ENGINE_TC


def _oracle_gene(O, G, X, nm, af=None):
    return O.gene(G.astype(float), af_of(G) if af is None else af, X, nm["resid"], nm["sigma2"])


@pytest.fixture(scope="module")
def eng(engine_cls):
    e = engine_cls(0)
    yield e
    e.close()


def _set_engine(eng, which):
    if which == 2 and eng.info("tc_available") != 1:
        pytest.skip("tensor-core engine not available in this build")
    eng.set_option("engine", which)


@pytest.mark.parametrize("which", ENGINES)
def test_c1_anchor(eng, oracle, which):
    _set_engine(eng, which)
    a = json.load(open(os.path.join(GOLD, "c1_anchor.json")))
    G = np.array(a["G"], dtype=np.int8)
    X = np.ones((9, 1))
    y = np.array(a["y"])
    eng.set_null_model(X, y)
    nm = eng.get_null_model()
    assert nm["sigma2"] == pytest.approx(a["sigma2"], rel=1e-12)
    eng.push_f64(G.astype(float), af_of(G))
    r = eng.flush()[0]
    assert r["Q"] == pytest.approx(a["Q"], rel=1e-8)
    assert int(r["davies_fault"]) == 1
    assert r["p_skat"] == pytest.approx(a["pvalue"], rel=1e-5)
    assert int(r["cmc_nonref"]) == a["cmc_nonref"]
    # what the reference prints (src/Model.h:2743 "%g\t%g")
    assert "%g" % r["Q"] == "23.741" and "%g" % r["p_skat"] == "0.324321"


def test_null_model_matches_oracle(eng, oracle):
    O = oracle
    for seed, N, C in ((1, 17, 1), (2, 1000, 3), (3, 40000, 5)):
        X, y = O.synth_covariates(seed, N, C)
        eng.set_null_model(X, y)
        nm = eng.get_null_model()
        ref = O.fit_null_linear(X, y)
        assert rel(nm["sigma2"], ref["sigma2"]) <= 1e-12
        assert np.max(np.abs(nm["resid"] - ref["resid"])) <= 1e-11
        assert np.max(np.abs(nm["xtx_inv"] - ref["xtx_inv"])) <= 1e-12 * np.max(np.abs(ref["xtx_inv"]))


CASES = [
    # seed, N, M, C, n_mono, n_flip
    (10, 50, 1, 1, 0, 0),
    (11, 97, 5, 1, 1, 1),
    (12, 1000, 30, 3, 2, 3),
    (13, 1000, 64, 3, 0, 5),
    (14, 4099, 50, 3, 3, 4),
    (15, 20000, 50, 3, 0, 0),
    (16, 513, 7, 2, 7, 0),      # every variant monomorphic -> NA
    (17, 2048, 33, 4, 1, 1),
]


@pytest.mark.parametrize("which", ENGINES)
@pytest.mark.parametrize("case", CASES)
def test_push_f64_and_i8_vs_oracle(eng, oracle, which, case):
    _set_engine(eng, which)
    O = oracle
    seed, N, M, C, n_mono, n_flip = case
    maf = None if N >= 1000 else np.linspace(0.05, 0.4, M)
    G, X, y = make_problem(O, seed, N, M, C, maf=maf, n_mono=n_mono, n_flip=n_flip)
    eng.set_null_model(X, y)
    nm = O.fit_null_linear(X, y)
    af = af_of(G)
    eng.push_f64(G.astype(float), af)          # reference boundary (double Matrix)
    eng.push_i8(G.T.copy(), af)                # packed hard calls
    eng.push_i8(G.T.copy(), None)              # AF derived by the engine
    res = eng.flush()
    ref, lam = _oracle_gene(O, G, X, nm)
    check_gene(res[0], ref, lam, ctx=f"f64 {case}")
    check_gene(res[1], ref, lam, ctx=f"i8 {case}")
    # f64 and i8 paths share the integer pipeline: identical bits
    for k in ("Q", "p_skat", "cmc_U", "zeg_U", "cmc_nonref", "lambda_max"):
        assert res[0][k] == res[1][k], k
    # without caller AF the engine uses the kept columns' own frequencies (no index quirk)
    keep = [j for j in range(M) if G[:, j].min() != G[:, j].max()]
    if keep:
        af2 = np.zeros(M)
        af2[: len(keep)] = af[keep]
        ref2, lam2 = _oracle_gene(O, G, X, nm, af=af2)
        check_gene(res[2], ref2, lam2, ctx=f"i8/noaf {case}")


@pytest.mark.parametrize("which", ENGINES)
def test_split_invariance_bitwise(eng, oracle, which):
    """exact integer accumulation => the split count cannot change a single bit"""
    _set_engine(eng, which)
    O = oracle
    G, X, y = make_problem(O, 21, 30000, 40, 3, n_flip=2)
    eng.set_null_model(X, y)
    outs = []
    for s in (1, 3, 8):
        eng.set_option("splits", s)
        eng.push_i8(G.T.copy(), af_of(G))
        outs.append(eng.flush()[0])
    eng.set_option("splits", 0)
    for o in outs[1:]:
        assert o.tobytes() == outs[0].tobytes()


def test_engines_agree_bitwise(eng, oracle):
    if eng.info("tc_available") != 1:
        pytest.skip("tensor-core engine not available")
    O = oracle
    G, X, y = make_problem(O, 22, 70000, 50, 3, n_flip=3, n_mono=1)
    eng.set_null_model(X, y)
    outs = []
    for which in (1, 2):
        eng.set_option("engine", which)
        eng.push_i8(G.T.copy(), af_of(G))
        outs.append(eng.flush()[0])
    eng.set_option("engine", 0)
    assert outs[0].tobytes() == outs[1].tobytes()


@pytest.mark.parametrize("which", ENGINES)
def test_loaded_synthetic_cohort(eng, oracle, which):
    """device generator == host twin (bit exact), and the zero-copy loaded path == oracle"""
    _set_engine(eng, which)
    O = oracle
    seed, N, M, ng, C = 20260925, 30011, 30, 6, 3
    X, y = O.synth_covariates(seed, N, C)
    eng.set_null_model(X, y)
    nm = O.fit_null_linear(X, y)
    vid = np.arange(ng * M, dtype=np.uint64)
    maf = O.synth_maf(seed, vid)
    t0, t1 = O.synth_thresholds(maf)
    eng.synth_load(O.synth_variant_key(seed, vid), t0, t1, ng, M)
    host = O.synth_genotypes(seed, vid, N, maf=maf)
    dev = eng.loaded_read(0, ng * M)
    assert np.array_equal(host, dev)
    res = eng.run_loaded()
    assert len(res) == ng
    for g in range(ng):
        Gg = host[g * M:(g + 1) * M].T
        ref, lam = _oracle_gene(O, Gg, X, nm)
        check_gene(res[g], ref, lam, ctx=f"loaded gene {g}")


SKATO_CASES = [(40, 700, 1, 1), (41, 1500, 2, 3), (42, 2000, 12, 3), (43, 5000, 50, 3), (44, 3000, 64, 2), (45, 900, 7, 1)]


@pytest.mark.parametrize("case", SKATO_CASES)
def test_skato_vs_oracle(eng, oracle, case):
    """SKAT-O (regression/SkatO.cpp) on the device vs the numpy + GSL 1.16 + reference-Davies oracle."""
    from oracle import skato_oracle as SO
    O = oracle
    seed, N, M, C = case
    # keep the carrier fraction well below 1: a constant CMC indicator is collinear with the
    # intercept and makes the burden score test numerically undefined (in the reference as well)
    hi = 0.35 if M <= 12 else 0.02
    G, X, y = make_problem(O, seed, N, M, C, maf=np.linspace(0.002, hi, M), n_flip=min(2, M - 1) if M > 1 else 0,
                           n_mono=1 if M > 5 else 0)
    eng.set_option("engine", 0)
    eng.set_option("skato", 1)
    try:
        eng.set_null_model(X, y)
        nm = O.fit_null_linear(X, y)
        af = af_of(G)
        eng.push_i8(G.T.copy(), af)
        r = eng.flush()[0]
    finally:
        eng.set_option("skato", 0)
    ref = SO.skato_gene(G.astype(float), af, X, nm["resid"])
    assert int(r["skato_ok"]) == int(ref["ok"])
    if ref["ok"]:
        assert rel(r["skato_Q"], ref["Q"]) <= 1e-6, (r["skato_Q"], ref["Q"])
        assert r["skato_rho"] == ref["rho"]
        # SKAT-O p <= 1e-3 rel (QAGS epsrel is 1.2e-4); the port reproduces GSL's nodes
        assert rel(r["skato_p"], ref["pvalue"]) <= 1e-5, (r["skato_p"], ref["pvalue"])
    # the SKAT / burden columns are unaffected by enabling SKAT-O
    ref2, lam = _oracle_gene(O, G, X, nm)
    check_gene(r, ref2, lam, ctx=f"skat with skato on {case}")


@pytest.mark.parametrize("case", [(50, 700, 8, 1, 0.02), (51, 3000, 30, 3, 0.01), (52, 2000, 64, 2, 0.3)])
def test_mean_imputed_and_dosage_genotypes(eng, oracle, case):
    """A0: DataConsolidator::imputeGenotypeToMean (src/DataConsolidator.cpp:217-245) fills missing
    calls with 2p -- not a hard call -- so such a gene takes the fp64 dosage path; same for dosages."""
    O = oracle
    seed, N, M, C, miss = case
    # rare variants for the wide genes: a constant CMC indicator makes the burden test undefined
    G, X, y = make_problem(O, seed, N, M, C, maf=np.linspace(0.004, 0.3 if M <= 8 else 0.03, M), n_flip=2, n_mono=1 if M > 8 else 0)
    eng.set_option("engine", 0)
    eng.set_null_model(X, y)
    nm = O.fit_null_linear(X, y)
    rng = np.random.default_rng(seed)
    Gd = G.astype(np.float64)
    af = np.zeros(M)
    for j in range(M):
        m = rng.random(N) < miss
        obs = Gd[~m, j]
        p = obs.sum() / (2 * len(obs)) if len(obs) else 0.0
        Gd[m, j] = 2.0 * p                                    # imputeGenotypeToMean
        af[j] = 0.5 * obs.sum() / N                            # GenotypeCounter::getAF divides by nSample incl. missing (F9)
    Gdos = np.clip(G + rng.normal(0, 0.05, G.shape), 0, 2)      # genuine dosages
    eng.push_f64(Gd, af)
    eng.push_f64(Gdos, None)
    eng.push_f64(G.astype(float), af_of(G))                    # a hard-call gene in the same flush
    res = eng.flush()
    ref, lam = O.gene(Gd, af, X, nm["resid"], nm["sigma2"])
    check_gene(res[0], ref, lam, ctx=f"imputed {case}")
    keep = [j for j in range(M) if Gdos[:, j].min() != Gdos[:, j].max()]
    af2 = np.zeros(M)
    af2[: len(keep)] = 0.5 * Gdos[:, keep].sum(axis=0) / N
    ref2, lam2 = O.gene(Gdos, af2, X, nm["resid"], nm["sigma2"])
    check_gene(res[1], ref2, lam2, ctx=f"dosage {case}")
    ref3, lam3 = _oracle_gene(O, G, X, nm)
    check_gene(res[2], ref3, lam3, ctx=f"hard calls {case}")


def test_int8_path_rejects_non_hard_calls(eng, oracle):
    O = oracle
    G, X, y = make_problem(O, 30, 500, 6, 1, maf=0.2)
    eng.set_null_model(X, y)
    Gb = G.T.copy()
    Gb[2, 3] = -1  # a missing code has no meaning in the packed hard-call format
    eng.push_i8(Gb, af_of(G))
    r = eng.flush()[0]
    assert int(r["status"]) == 5  # RVT_GENE_BADVALUE: reported, never silently computed


def test_errors_are_loud(eng):
    import rvtests_b200
    e2 = rvtests_b200.GeneEngine(0)
    with pytest.raises(rvtests_b200.RvtError):
        e2.push_i8(np.zeros((3, 10), dtype=np.int8))  # no null model yet
    X = np.ones((10, 2))
    X[:, 0] = 2.0
    with pytest.raises(rvtests_b200.RvtError):
        e2.set_null_model(X, np.arange(10.0))  # column 0 is not the intercept
    e2.close()


def test_full_size_properties(engine_cls, oracle):
    """BASELINE.json full size (N = 500 000 samples x M = 50 variants): size-independent properties
    plus two genes checked against the oracle.
      * the tcgen05 and the dp4a sweeps agree bit for bit, for every split count
      * recoding a variant as 2-g (ALT <-> REF) and telling the engine to flip it back changes nothing
      * 2 genes vs the CPU oracle (Q 1e-6, p 1e-4, NonRefSite exact)"""
    O = oracle
    from rvtests_b200 import synth
    seed, N, M, ng, C = 20260925, 500_000, 50, 6, 3
    X, y = synth.covariates(seed, N, C)
    keys, t0, t1 = synth.variant_params(seed, 0, ng * M)
    eng = engine_cls(0)
    eng.set_null_model(X, y)
    eng.synth_load(keys, t0, t1, ng, M)
    outs = {}
    for which in (1, 2):
        for splits in (0, 5):
            eng.set_option("engine", which)
            eng.set_option("splits", splits)
            outs[(which, splits)] = eng.run_loaded()
    base = outs[(2, 0)]
    for k, v in outs.items():
        assert v.tobytes() == base.tobytes(), k
    assert np.all(base["status"] == 0)
    # oracle on two genes (host twin regenerates the same genotypes)
    nm = O.fit_null_linear(X, y)
    for g in (0, ng - 1):
        Gg = eng.loaded_read(g * M, M)                      # (M, N) int8
        ref, lam = O.gene(Gg.T.astype(np.float64), 0.5 * Gg.sum(axis=1) / N, X, nm["resid"], nm["sigma2"])
        check_gene(base[g], ref, lam, ctx=f"full size gene {g}")
    # flip invariance through the host int8 path (gene 1)
    eng.set_option("engine", 0)
    eng.set_option("splits", 0)
    G1 = eng.loaded_read(M, M)
    af = 0.5 * G1.sum(axis=1) / N
    G1f = G1.copy()
    G1f[[3, 17, 40]] = 2 - G1f[[3, 17, 40]]                  # these rows are now ALT-major: the engine flips them back
    eng.push_i8(G1, af)
    eng.push_i8(G1f, af)                                     # same AF => same weights (the quirk path takes AF from the caller)
    r = eng.flush()
    for k in ("Q", "p_skat", "cmc_nonref", "cmc_p", "zeg_p", "lambda_max"):
        assert rel(r[0][k], r[1][k]) <= 1e-9, k
    assert r[0]["cmc_nonref"] == r[1]["cmc_nonref"] == base[1]["cmc_nonref"]
    eng.close()


WIDE_CASES = [
    # seed, N, M, C, n_mono, n_flip  -- genes wider than one 64-variant tensor-core tile (wide.cuh)
    (60, 3000, 65, 3, 1, 1),
    (61, 3000, 150, 3, 2, 3),
    (62, 5003, 300, 2, 0, 2),
    (63, 70000, 129, 3, 1, 2),
]


@pytest.mark.parametrize("case", WIDE_CASES)
def test_wide_genes_vs_oracle(eng, oracle, case):
    """The reference has no limit on the variants of a gene (MixtureChiSquare grows its lambda array,
    regression/MixtureChiSquare.h:26-52): M > 64 goes through tile pairs + a global-memory tail."""
    from rvtests_b200.synth import pack_bed
    if eng.info("tc_available") != 1:
        pytest.skip("wide genes need the tensor-core sweep")
    O = oracle
    seed, N, M, C, n_mono, n_flip = case
    G, X, y = make_problem(O, seed, N, M, C, maf=np.linspace(0.001, 0.02, M), n_mono=n_mono, n_flip=n_flip)
    Gs, _, _ = make_problem(O, seed + 100, N, 20, C, maf=np.linspace(0.01, 0.1, 20))
    eng.set_option("engine", 0)
    eng.set_null_model(X, y)
    nm = O.fit_null_linear(X, y)
    af = af_of(G)
    eng.push_i8(Gs.T.copy(), af_of(Gs))         # ordinary genes around the wide ones: record order is push order
    eng.push_i8(G.T.copy(), af)
    eng.push_bed(pack_bed(G.T), af)
    eng.push_i8(Gs.T.copy(), af_of(Gs))
    eng.push_f64(G.astype(float), af)
    res = eng.flush()
    assert len(res) == 5
    ref, lam = _oracle_gene(O, G, X, nm)
    refs, lams = _oracle_gene(O, Gs, X, nm)
    for k in (1, 2, 4):
        check_gene(res[k], ref, lam, ctx=f"wide[{k}] {case}")
    assert res[1].tobytes() == res[2].tobytes() == res[4].tobytes()
    for k in (0, 3):
        check_gene(res[k], refs, lams, ctx=f"ordinary gene beside a wide one [{k}] {case}")
    # (n_lambda is not compared: eigenvalues of duplicate rare variants are rounding noise around the 1e-30 cut)


def test_wide_gene_skato_vs_oracle(eng, oracle):
    from oracle import skato_oracle as SO
    if eng.info("tc_available") != 1:
        pytest.skip("wide genes need the tensor-core sweep")
    O = oracle
    seed, N, M, C = 64, 4000, 100, 3
    G, X, y = make_problem(O, seed, N, M, C, maf=np.linspace(0.002, 0.02, M), n_flip=2, n_mono=1)
    eng.set_option("engine", 0)
    eng.set_option("skato", 1)
    try:
        eng.set_null_model(X, y)
        nm = O.fit_null_linear(X, y)
        af = af_of(G)
        eng.push_i8(G.T.copy(), af)
        r = eng.flush()[0]
    finally:
        eng.set_option("skato", 0)
    ref = SO.skato_gene(G.astype(float), af, X, nm["resid"])
    assert int(r["skato_ok"]) == int(ref["ok"]) == 1
    assert rel(r["skato_Q"], ref["Q"]) <= 1e-6
    assert r["skato_rho"] == ref["rho"]
    assert rel(r["skato_p"], ref["pvalue"]) <= 1e-5, (r["skato_p"], ref["pvalue"])
    ref2, lam = _oracle_gene(O, G, X, nm)
    check_gene(r, ref2, lam, ctx="wide gene with skato on")


def test_too_wide_gene_is_refused(eng, oracle):
    import rvtests_b200
    O = oracle
    X, y = O.synth_covariates(5, 64, 1)
    eng.set_null_model(X, y)
    with pytest.raises(rvtests_b200.RvtError):
        eng.push_i8(np.zeros((2049, 64), dtype=np.int8))


@pytest.mark.parametrize("case", [(161, 3000, 100, 3, 0.01, 2, 1), (162, 2500, 200, 2, 0.02, 3, 2), (163, 700, 65, 1, 0.05, 1, 0)])
def test_wide_genes_with_missing_calls_vs_oracle(eng, oracle, case):
    """A gene of more than 64 variants WITH missing calls (PLINK code 01): the reference imputes to the mean
    (DataConsolidator::imputeGenotypeToMean, src/DataConsolidator.cpp:217-245) whatever the width.  Here every tile is split
    into H (hard calls + fill) and Mi (indicators), the gene is swept as the 2M rows [H ; Mi] (diagonal + pair units, exact
    integers) and the tail combines G = H + Mi diag(delta).  Flipped and monomorphic columns with missing calls, SKAT + burden
    + SKAT-O, an ordinary gene and a complete wide gene in the same flush."""
    from rvtests_b200.synth import pack_bed
    from oracle import skato_oracle as SO
    if eng.info("tc_available") != 1:
        pytest.skip("wide genes need the tensor-core sweep")
    O = oracle
    seed, N, M, C, miss, n_flip, n_mono = case
    G, X, y = make_problem(O, seed, N, M, C, maf=np.linspace(0.002, 0.03, M), n_mono=n_mono, n_flip=n_flip)
    Gs, _, _ = make_problem(O, seed + 100, N, 20, C, maf=np.linspace(0.01, 0.1, 20))
    rng = np.random.default_rng(seed)
    mask = rng.random((M, N)) < miss
    mask[5] = False                                              # a fully called variant among them
    mask[M - 1, : N // 3] = True                                 # a third of the last variant missing
    bed = pack_bed(G.T, mask)
    raw = O.bed_decode_fast(bed, N).T
    Gd = O.impute_mean(raw)
    af = 0.5 * np.where(raw >= 0, raw, 0.0).sum(axis=0) / N      # GenotypeCounter::getAF over nSample
    eng.set_option("engine", 0)
    eng.set_option("skato", 1)
    try:
        eng.set_null_model(X, y)
        nm = O.fit_null_linear(X, y)
        eng.push_i8(Gs.T.copy(), af_of(Gs))
        eng.push_bed(bed, af)
        eng.push_bed(pack_bed(G.T), af_of(G))                    # the same gene fully called
        eng.push_bed(bed, None)                                  # frequencies left to the engine
        res = eng.flush()
    finally:
        eng.set_option("skato", 0)
    assert len(res) == 4 and [int(r["status"]) for r in res] == [0, 0, 0, 0]
    ref, lam = O.gene(Gd, af, X, nm["resid"], nm["sigma2"])
    check_gene(res[1], ref, lam, ctx=f"wide+missing {case}")
    ref_c, lam_c = _oracle_gene(O, G, X, nm)
    check_gene(res[2], ref_c, lam_c, ctx=f"wide complete beside it {case}")
    refs, lams = _oracle_gene(O, Gs, X, nm)
    check_gene(res[0], refs, lams, ctx=f"ordinary gene beside it {case}")
    so = SO.skato_gene(Gd, af, X, nm["resid"])
    assert int(res[1]["skato_ok"]) == int(so["ok"]) == 1
    assert rel(res[1]["skato_Q"], so["Q"]) <= 1e-6 and res[1]["skato_rho"] == so["rho"]
    assert rel(res[1]["skato_p"], so["pvalue"]) <= 1e-5
    # no frequencies supplied: the weight of a kept column comes from ITS OWN imputed column sum (the index quirk F9 is a
    # property of caller-supplied frequencies) -- for the oracle, which applies the quirk, line the values up accordingly
    af_imp = Gd.sum(axis=0) / (2.0 * N)
    kept = [j for j in range(M) if Gd[:, j].min() != Gd[:, j].max()]
    af_q = af_imp.copy()
    af_q[: len(kept)] = af_imp[kept]
    ref_n, lam_n = O.gene(Gd, af_q, X, nm["resid"], nm["sigma2"])
    check_gene(res[3], ref_n, lam_n, ctx=f"wide+missing, engine frequencies {case}")


def test_push_f64_of_a_mean_imputed_matrix_takes_the_integer_paths(eng, oracle):
    """ModelFitter::fit() sees the genotype Matrix AFTER DataConsolidator::imputeGenotypeToMean (src/DataConsolidator.cpp:217-245):
    hard calls plus, per column, ONE fractional value 2 p^ at the missing entries.  rvt_gene_push_f64 recognises that pattern and
    routes the gene like a 2-bit push with code 01 -- augmented tensor-core sweep, wide operand tiles -- so the literal drop-in
    gets the same records as the 2-bit form, bit for bit; anything else (a real dosage) keeps the fp64 path."""
    import rvtests_b200
    from rvtests_b200.synth import pack_bed
    if eng.info("tc_available") != 1:
        pytest.skip("needs the tensor-core sweep")
    O = oracle
    N, C = 2600, 3
    X, y = O.synth_covariates(171, N, C)
    eng.set_option("engine", 0)
    eng.set_null_model(X, y)
    nm = O.fit_null_linear(X, y)
    rng = np.random.default_rng(171)
    for M in (40, 100):
        G, _, _ = make_problem(O, 172 + M, N, M, C, maf=np.linspace(0.003, 0.04, M), n_flip=2, n_mono=1)
        mask = rng.random((M, N)) < 0.015
        mask[3] = False
        bed = pack_bed(G.T, mask)
        raw = O.bed_decode_fast(bed, N).T
        Gd = O.impute_mean(raw)
        af = 0.5 * np.where(raw >= 0, raw, 0.0).sum(axis=0) / N
        eng.push_f64(Gd, af)
        eng.push_bed(bed, af)
        r = eng.flush()
        assert [int(v) for v in r["status"]] == [0, 0]
        assert r[0].tobytes() == r[1].tobytes(), f"M={M}: Matrix form and 2-bit form differ"
        if M <= 62:
            assert int(eng.info("last_aug")) == 2                      # both rode the augmented sweep
        ref, lam = O.gene(Gd, af, X, nm["resid"], nm["sigma2"])
        check_gene(r[0], ref, lam, ctx=f"imputed Matrix M={M}")
        # a real dosage in one entry: not the pattern -> the generic fp64 path (or refused beyond 64 variants)
        Gx = Gd.copy()
        Gx[7, 2] = 0.61
        if M <= 62:
            eng.push_f64(Gx, af)
            rx = eng.flush()
            refx, lamx = O.gene(Gx, af, X, nm["resid"], nm["sigma2"])
            check_gene(rx[0], refx, lamx, ctx=f"dosage Matrix M={M}")
        else:   # beyond 64 variants the matrix stays on the device for the dense statistics
            eng.push_f64(Gx, af)
            rx = eng.flush()
            refx, lamx = O.gene(Gx, af, X, nm["resid"], nm["sigma2"])
            check_gene(rx[0], refx, lamx, ctx=f"wide dosage Matrix M={M}")


@pytest.mark.parametrize("binary", [False, True])
def test_wide_genes_with_real_dosages_vs_oracle(engine_cls, oracle, binary):
    """A gene of more than 64 variants pushed as doubles with REAL dosages (BGEN): the matrix stays on the device and the
    statistics are dense fp64 (k_wide_dos_cols / _gram / _burden) into the same tail; quantitative and binary trait, flipped and
    monomorphic columns, SKAT + burden + SKAT-O, an ordinary gene and a hard-call wide gene in the same flush."""
    from oracle import binary_oracle as BIN
    from oracle import skato_oracle as SO
    import rvtests_b200
    O = oracle
    seed, N, M, C = 181, 2100, 150, 3
    # (rare enough that a fifth of the samples carries nothing: with a carrier in every sample the CMC indicator is the intercept)
    G, X, yq = make_problem(O, seed, N, M, C, maf=np.linspace(0.002, 0.012, M), n_flip=0, n_mono=1)
    Gs, _, _ = make_problem(O, seed + 1, N, 20, C, maf=np.linspace(0.01, 0.1, 20))
    rng = np.random.default_rng(seed)
    Gd = G.astype(np.float64)
    soft = (rng.random(Gd.shape) < 0.01) | ((G > 0) & (rng.random(Gd.shape) < 0.5))
    Gd[soft] = np.clip(Gd[soft] + rng.normal(scale=0.2, size=int(soft.sum())), 0.0, 2.0)     # imputation uncertainty
    poly = [j for j in range(M) if G[:, j].min() != G[:, j].max()]
    mono = [j for j in range(M) if j not in poly]
    Gd[:, mono] = G[:, mono]                                      # the monomorphic column stays monomorphic
    af = 0.5 * Gd.mean(axis=0)
    eng = engine_cls(0)
    if eng.info("tc_available") != 1:
        eng.close()
        pytest.skip("wide genes need the tensor-core sweep")
    try:
        if binary:
            y = (rng.random(N) < 1.0 / (1.0 + np.exp(0.6 - 0.4 * X[:, 1]))).astype(np.float64)
            eng.set_null_model(X, y, binary=True)
            nm = BIN.fit_null_logistic(X, y)
        else:
            eng.set_null_model(X, yq)
            nm = O.fit_null_linear(X, yq)
        eng.set_option("skato", 1)
        eng.push_i8(Gs.T.copy(), af_of(Gs))
        eng.push_f64(Gd, af)
        eng.push_i8(G.T.copy(), af_of(G))
        res = eng.flush()
    finally:
        eng.close()
    assert [int(v) for v in res["status"]] == [0, 0, 0]
    r = res[1]
    if binary:
        ref = BIN.gene(Gd, af, X, nm)
        so = SO.skato_gene(Gd, af, X, nm["resid"], vv=nm["v"])
        assert int(r["m_poly"]) == ref["m_poly"] and rel(r["Q"], ref["Q"]) <= 1e-6 and rel(r["p_skat"], ref["p_skat"]) <= 1e-4
        for pre in ("cmc", "zeg"):
            b = ref[pre]
            assert abs(r[pre + "_U"] - b["U"]) <= 1e-6 * max(abs(b["U"]), np.sqrt(b["V"])) and rel(r[pre + "_V"], b["V"]) <= 1e-6
        assert int(r["cmc_nonref"]) == ref["cmc"]["nonref"]
        refh = BIN.gene(G.astype(float), af_of(G), X, nm)
        assert rel(res[2]["Q"], refh["Q"]) <= 1e-6
    else:
        ref, lam = O.gene(Gd, af, X, nm["resid"], nm["sigma2"])
        check_gene(r, ref, lam, ctx="wide dosage gene")
        so = SO.skato_gene(Gd, af, X, nm["resid"])
        refh, lamh = _oracle_gene(O, G, X, nm)
        check_gene(res[2], refh, lamh, ctx="hard-call wide gene beside it")
    assert int(r["skato_ok"]) == int(so["ok"]) == 1
    assert rel(r["skato_Q"], so["Q"]) <= 1e-6 and r["skato_rho"] == so["rho"] and rel(r["skato_p"], so["pvalue"]) <= 1e-5

