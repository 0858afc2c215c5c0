"""GPU: genes with missing calls on the AUGMENTED tensor-core sweep (rvtests_b200/csrc/sweep_aug.cuh) -- the imputed
matrix of DataConsolidator::imputeGenotypeToMean (src/DataConsolidator.cpp:217-245) as H + M diag(delta), every sum an exact
integer product over the rows [H ; M] -- against the oracle (which decodes the same bytes with the reference's table and
imputes with the reference's rule), against the sparse CUDA-core kernel it replaces (option "aug" = 0), and with SKAT-O on."""
import numpy as np
import pytest

from util import af_of, check_gene, make_problem, rel

pytestmark = pytest.mark.gpu


def _problem(O, seed, N, M, C, miss, n_flip, n_mono, hi=0.05):
    from rvtests_b200.synth import pack_bed
    G, X, y = make_problem(O, seed, N, M, C, maf=np.linspace(0.004, hi, M), n_flip=n_flip, n_mono=n_mono)
    rng = np.random.default_rng(seed)
    mask = rng.random((M, N)) < miss
    if M > 2:
        mask[0] = False                                   # one fully called variant
        mask[1, : N // 3] = True                          # one variant with a third of its calls missing
    bed = pack_bed(G.T, mask)
    raw = O.bed_decode_fast(bed, N).T                     # (N, M) with -9
    Gd = O.impute_mean(raw)
    af = 0.5 * np.where(raw >= 0, raw, 0.0).sum(axis=0) / N   # GenotypeCounter::getAF divides by nSample incl. missing
    return G, X, y, bed, Gd, af


CASES = [(901, 700, 8, 1, 0.02, 1, 0), (902, 3001, 30, 3, 0.01, 2, 1), (903, 20011, 50, 3, 0.01, 3, 1), (904, 1500, 1, 2, 0.05, 0, 0),
         (905, 5000, 62, 4, 0.002, 4, 2), (906, 2500, 2, 3, 0.2, 1, 0), (907, 70000, 50, 3, 0.01, 2, 0), (908, 4000, 40, 6, 0.03, 2, 1)]


@pytest.mark.parametrize("case", CASES)
def test_missing_calls_on_the_augmented_sweep_vs_oracle(engine_cls, oracle, case):
    from rvtests_b200.synth import pack_bed
    O = oracle
    seed, N, M, C, miss, n_flip, n_mono = case
    G, X, y, bed, Gd, af = _problem(O, seed, N, M, C, miss, n_flip, n_mono)
    nm = O.fit_null_linear(X, y)
    eng = engine_cls(0)
    try:
        eng.set_null_model(X, y)
        eng.push_bed(bed, af)
        eng.push_bed(pack_bed(G.T), af_of(G))             # a fully called gene in the same flush
        eng.push_bed(bed, af)
        r = eng.flush()
        assert int(eng.info("last_aug")) == 2
        eng.set_option("aug", 0)                          # the sparse CUDA-core kernel
        eng.push_bed(bed, af)
        r_sparse = eng.flush()
        assert int(eng.info("last_aug")) == 0
    finally:
        eng.close()
    ref, lam = O.gene(Gd, af, X, nm["resid"], nm["sigma2"])
    check_gene(r[0], ref, lam, ctx=f"aug {case}")
    ref2, lam2 = O.gene(G.astype(float), af_of(G), X, nm["resid"], nm["sigma2"])
    check_gene(r[1], ref2, lam2, ctx=f"complete {case}")
    assert r[0].tobytes() == r[2].tobytes()               # exact integer sums: bit-reproducible
    # (the sparse kernel multiplies by the fp64 residual, the sweep by its 2^-30 fixed-point digits: ~1e-9 apart)
    for k in ("Q", "p_skat", "cmc_p", "zeg_p", "lambda_max", "cmc_U", "zeg_U"):
        assert abs(r[0][k] - r_sparse[0][k]) <= 1e-7 * max(abs(r_sparse[0][k]), 1e-300) + 1e-7 * abs(r_sparse[0]["cmc_V"]) ** 0.5, (k, r[0][k], r_sparse[0][k])
    assert int(r[0]["cmc_nonref"]) == int(r_sparse[0]["cmc_nonref"]) == ref.cmc_nonref


def test_augmented_sweep_with_skato_and_large_batch(engine_cls, oracle):
    """many genes with missing calls in one flush (several sweep units per CTA, both TMEM accumulators in turn), SKAT-O on"""
    from oracle import skato_oracle as SO
    O = oracle
    N, C = 9000, 3
    X, y = O.synth_covariates(950, N, C)
    nm = O.fit_null_linear(X, y)
    eng = engine_cls(0)
    probs = []
    try:
        eng.set_option("skato", 1)
        eng.set_null_model(X, y)
        for g in range(40):
            M = [50, 12, 33, 62, 5][g % 5]
            G, _, _, bed, Gd, af = _problem(O, 960 + g, N, M, C, 0.01 + 0.002 * (g % 7), 1 if M > 3 else 0, 1 if g % 4 == 0 and M > 5 else 0)
            probs.append((Gd, af))
            eng.push_bed(bed, af)
        r = eng.flush()
        assert int(eng.info("last_aug")) == 40
    finally:
        eng.close()
    for g in (0, 7, 13, 24, 39):
        Gd, af = probs[g]
        ref, lam = O.gene(Gd, af, X, nm["resid"], nm["sigma2"])
        check_gene(r[g], ref, lam, ctx=f"gene {g}")
        so = SO.skato_gene(Gd, af, X, nm["resid"])
        assert int(r[g]["skato_ok"]) == int(bool(so["ok"]))
        if so["ok"]:
            assert rel(r[g]["skato_Q"], so["Q"]) <= 1e-6 and r[g]["skato_rho"] == so["rho"] and rel(r[g]["skato_p"], so["pvalue"]) <= 1e-4
