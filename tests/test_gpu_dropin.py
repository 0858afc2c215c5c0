"""GPU: the drop-in, literally.  oracle/_ref/libdropin_ref.so holds the REFERENCE's own src/ModelManager.cpp (with the
two registration lines of rvtests_b200/host/ModelB200.h applied to a scratch copy at build time) and the gene loop of
src/Main.cpp:1221-1254 on a real DataConsolidator.  `--kernel skat,skato --burden cmc,zeggini` is created BY NAME through
ModelManager::create twice: once as the reference's SkatTest / SkatOTest / CMCTest / ZegginiTest, once (RVTESTS_B200=1) as
the B200 adapters -- true ModelFitter subclasses in ModelManager's std::vector<ModelFitter*> -- and the `.assoc` files
that ModelManager's own writers produced are compared column by column.  Missing calls (mean imputation), flipped and
monomorphic variants, a monomorphic-only gene (NA line) and the binary trait are all in the gene list."""
import numpy as np
import pytest

from util import make_problem

pytestmark = pytest.mark.gpu


def _genes(O, seed, N, C):
    rng = np.random.default_rng(seed)
    genes = []
    X = None
    for gi, (M, nm_, nf, miss) in enumerate([(8, 0, 1, 0.0), (30, 2, 2, 0.0), (1, 0, 0, 0.0), (50, 1, 3, 0.01), (12, 0, 0, 0.03),
                                              (3, 3, 0, 0.0), (64, 0, 4, 0.0),
                                              (70, 0, 1, 0.0), (90, 1, 2, 0.01)]):       # wider than a tile; the last with missing calls
        G, X, y = make_problem(O, seed + gi, N, M, C, maf=np.linspace(0.004, 0.06, M), n_mono=nm_, n_flip=nf)
        G = G.astype(np.float64)
        if miss > 0:
            G[rng.random(G.shape) < miss] = -9.0          # missing call, as GenotypeExtractor hands it over
        genes.append(G)
    X, y = O.synth_covariates(seed, N, C)
    return genes, X, y


def _num(x):
    return None if x == "NA" else float(x)


def _compare(ref, b2, tol):
    for model in ("Skat", "SkatO", "CMC", "Zeggini"):
        cr, hr, rr = ref[model]
        cb, hb, rb = b2[model]
        assert hr == hb, (model, hr, hb)                      # same header line
        assert len(rr) == len(rb) and len(rr) > 0, model
        for lr, lb in zip(rr, rb):
            assert lr[:5] == lb[:5], (model, lr, lb)           # Gene RANGE N_INFORMATIVE NumVar NumPolyVar
            assert len(lr) == len(lb), (model, lr, lb)
            for k in range(5, len(lr)):
                a, b = _num(lr[k]), _num(lb[k])
                assert (a is None) == (b is None), (model, hr[k], lr, lb)
                if a is None:
                    continue
                if hr[k] in ("NonRefSite", "rho"):
                    assert a == b, (model, hr[k], lr, lb)
                else:   # "%g" prints 6 digits; the reference's SKAT is float32 (regression/Skat.cpp:41-56)
                    t = tol[model] * (4.0 if (model == "Skat" and hr[k] == "Pvalue") else 1.0)
                    assert abs(a - b) <= t * max(abs(a), abs(b), 1e-300), (model, hr[k], lr, lb)


def test_dropin_quantitative(oracle, tmp_path):
    O = oracle
    if O.ref_dropin() is None:
        pytest.skip("oracle/_ref/libdropin_ref.so not built")
    genes, X, y = _genes(O, 301, 2500, 3)
    ref = O.dropin_run_gene_models(genes, X[:, 1:], y, str(tmp_path / "ref"), use_b200=False)
    b2 = O.dropin_run_gene_models(genes, X[:, 1:], y, str(tmp_path / "b200"), use_b200=True, batch=3)
    _compare(ref, b2, {"Skat": 5e-5, "SkatO": 2e-5, "CMC": 2e-5, "Zeggini": 2e-5})
    # and the stock path of the patched ModelManager is the reference's: identical text to the unpatched model layer
    stock = O.ref_run_gene_models(genes, X[:, 1:], y, str(tmp_path / "stock"), n_perm=0)
    for model in ("Skat", "SkatO", "CMC", "Zeggini"):
        assert stock[model][2] == ref[model][2], model


def test_dropin_binary(oracle, tmp_path):
    O = oracle
    if O.ref_dropin() is None:
        pytest.skip("oracle/_ref/libdropin_ref.so not built")
    genes, X, _ = _genes(O, 302, 2000, 2)
    rng = np.random.default_rng(302)
    y = (rng.random(len(X)) < 1.0 / (1.0 + np.exp(0.6 - 0.5 * X[:, 1]))).astype(np.float64)
    # (genes wider than a tile included: binary-trait statistics of any width come from the tiles, csrc/wide.cuh)
    ref = O.dropin_run_gene_models(genes, X[:, 1:], y, str(tmp_path / "ref"), use_b200=False, binary=True)
    b2 = O.dropin_run_gene_models(genes, X[:, 1:], y, str(tmp_path / "b200"), use_b200=True, binary=True, batch=4)
    _compare(ref, b2, {"Skat": 5e-5, "SkatO": 2e-5, "CMC": 2e-5, "Zeggini": 2e-5})


def _meta_problem(O, seed, N, nv, C):
    rng = np.random.default_rng(seed)
    maf = 10 ** rng.uniform(-2.3, np.log10(0.4), nv)
    u = rng.random((N, nv))
    G = (u < (1 - (1 - maf) ** 2)[None, :]).astype(np.float64) + (u < (maf * maf)[None, :])
    G[:, 7] = 0.0                                   # a monomorphic variant: NA statistics, never queued by MetaCov
    X, y = O.synth_covariates(seed, N, C)
    pos = np.cumsum(rng.integers(50, 900, nv)).astype(np.int32)
    return G, pos, X, y


def _compare_comments(cr, cb):
    """the '##' block: the summary header verbatim, the null-model estimates to the printed precision"""
    assert len(cr) == len(cb) and len(cr) > 0, (cr, cb)
    for a, b in zip(cr, cb):
        if a.startswith("## - "):
            fa, fb = a.split("\t"), b.split("\t")
            assert fa[0] == fb[0] and len(fa) == len(fb), (a, b)
            for xa, xb in zip(fa[1:], fb[1:]):
                if xa in ("Beta", "SD", "NA") or xb == "NA":
                    assert xa == xb, (a, b)
                else:
                    assert abs(float(xa) - float(xb)) <= 2e-5 * max(abs(float(xa)), 1e-300), (a, b)
        else:
            assert a == b, (a, b)


def _compare_meta(ref, b2):
    cr, hr, rr = ref["MetaScore"]
    cb, hb, rb = b2["MetaScore"]
    _compare_comments(cr, cb)
    _compare_comments(ref["MetaCov"][0], b2["MetaCov"][0])
    assert hr == hb and len(rr) == len(rb) and len(rr) > 0
    exact = {"AF", "INFORMATIVE_ALT_AC", "CALL_RATE", "N_REF", "N_HET", "N_ALT"}
    for lr, lb in zip(rr, rb):
        assert lr[:5] == lb[:5], (lr, lb)
        for k in range(5, len(hr)):
            fa, fb = lr[k].split(":"), lb[k].split(":")     # binary trait: all:case:control
            assert len(fa) == len(fb), (hr[k], lr, lb)
            for xa, xb in zip(fa, fb):
                a, b = _num(xa), _num(xb)
                assert (a is None) == (b is None), (hr[k], lr, lb)
                if a is None:
                    continue
                if hr[k] in exact:
                    assert a == b, (hr[k], lr, lb)
                else:
                    assert abs(a - b) <= 3e-5 * max(abs(a), abs(b), 1e-300), (hr[k], lr, lb)
    _, hcr, rcr = ref["MetaCov"]
    _, hcb, rcb = b2["MetaCov"]
    assert hcr == hcb and len(rcr) == len(rcb) and len(rcr) > 0
    for lr, lb in zip(rcr, rcb):
        assert lr[:5] == lb[:5], (lr, lb)             # CHROM START_POS END_POS NUM_MARKER MARKER_POS: the window logic
        pa, pb = lr[5].split(":"), lb[5].split(":")    # binary trait: band ':' covXZ / n ':' covZZ / n
        assert len(pa) == len(pb), (lr, lb)
        ca = np.array([float(x) for x in pa[0].split(",")])
        cb_ = np.array([float(x) for x in pb[0].split(",")])
        assert len(ca) == len(cb_)
        # the reference accumulates the products in float32 (FloatMatrixRef, src/Model.cpp:534-554); a binary trait works on
        # UNcentred genotypes, so g_i'W g_j and the projection term cancel to a small entry: wider floor there
        floor = 5e-2 if len(pa) > 1 else 1e-2
        assert np.all(np.abs(ca - cb_) <= 2e-4 * np.maximum(np.abs(ca), ca[0] * floor)), (lr[:4], ca, cb_)
        for xa, xb in zip(pa[1:], pb[1:]):
            va = np.array([float(x) for x in xa.split(",")])
            vb = np.array([float(x) for x in xb.split(",")])
            assert len(va) == len(vb) and np.all(np.abs(va - vb) <= 2e-4 * np.max(np.abs(va))), (lr[:4], xa, xb)


def test_dropin_meta_score_cov(oracle, tmp_path):
    """--meta score[se],cov[windowSize=..] created by name through the reference's ModelManager: MetaScoreTest / MetaCovTest vs
    MetaScoreTestB200 / MetaCovTestB200 (ModelFitter subclasses) in the reference's single-variant loop.  Intercept-only
    model: with covariates the reference's computeQuadraticForm multiplies non-conforming shapes."""
    O = oracle
    if O.ref_dropin() is None:
        pytest.skip("oracle/_ref/libdropin_ref.so not built")
    G, pos, X, y = _meta_problem(O, 311, 1800, 150, 1)
    ref = O.dropin_run_meta_models(G, pos, X[:, 1:], y, 4000, str(tmp_path / "ref"), use_b200=False, se=True)
    b2 = O.dropin_run_meta_models(G, pos, X[:, 1:], y, 4000, str(tmp_path / "b200"), use_b200=True, se=True, segment=64)
    _compare_meta(ref, b2)


def test_dropin_meta_binary_trait(oracle, tmp_path):
    """A case/control phenotype through the reference's ModelManager::setBinaryOutcome: MetaUnrelatedBinary /
    MetaCovUnrelatedBinary (src/Model.h:3669-3784, src/Model.cpp:695-778) vs the B200 adapters, which read isBinaryOutcome() at
    fit time -- all:case:control site columns, the ':covXZ:covZZ' tail of every MetaCov line.  Intercept-only model: with
    covariates the reference's score test solves a 1 x 1 matrix against a d x d identity (LogisticRegressionScoreTest.cpp:292)."""
    O = oracle
    if O.ref_dropin() is None:
        pytest.skip("oracle/_ref/libdropin_ref.so not built")
    G, pos, X, _y = _meta_problem(O, 312, 1500, 140, 1)
    rng = np.random.default_rng(9)
    eta = -0.5 + 0.4 * (G[:, 20] - G[:, 20].mean())
    y = (rng.random(len(eta)) < 1.0 / (1.0 + np.exp(-eta))).astype(np.float64)
    ref = O.dropin_run_meta_models(G, pos, X[:, 1:], y, 4000, str(tmp_path / "ref"), use_b200=False, se=True, binary=True)
    b2 = O.dropin_run_meta_models(G, pos, X[:, 1:], y, 4000, str(tmp_path / "b200"), use_b200=True, se=True, segment=64, binary=True)
    assert ":" in ref["MetaScore"][2][0][5]
    _compare_meta(ref, b2)
