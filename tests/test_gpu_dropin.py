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
                                              (3, 3, 0, 0.0), (64, 0, 4, 0.0)]):
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
    genes = [g for g in genes if g.shape[1] <= 64]
    ref = O.dropin_run_gene_models(genes, X[:, 1:], y, str(tmp_path / "ref"), use_b200=False, binary=True)
    b2 = O.dropin_run_gene_models(genes, X[:, 1:], y, str(tmp_path / "b200"), use_b200=True, binary=True, batch=4)
    _compare(ref, b2, {"Skat": 5e-5, "SkatO": 2e-5, "CMC": 2e-5, "Zeggini": 2e-5})
