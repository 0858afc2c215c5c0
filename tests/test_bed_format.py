"""PLINK 2-bit rows (rvt_gene_push_bed): the host packer against the oracle's restatement of the
reference decoder (libVcf/PlinkInputFile.cpp:23-47), mean imputation against
DataConsolidator::imputeGenotypeToMean, and -- on the GPU -- the 2-bit push against the int8 push
(identical bits) and against the oracle for genes with missing calls."""
import numpy as np
import pytest

from util import af_of, check_gene, make_problem


def test_pack_bed_roundtrip_through_reference_decoder(oracle):
    from rvtests_b200.synth import pack_bed
    O = oracle
    rng = np.random.default_rng(5)
    for N in (1, 3, 4, 5, 127, 130):
        Gt = rng.integers(0, 3, size=(6, N)).astype(np.int8)
        miss = rng.random((6, N)) < 0.1
        bed = pack_bed(Gt, miss)
        assert bed.shape == (6, (N + 3) // 4) and bed.dtype == np.uint8
        want = np.where(miss, -9.0, Gt.astype(float))
        assert np.array_equal(O.bed_decode(bed, N), want)
        assert np.array_equal(O.bed_decode_fast(bed, N), want)
    # the four codes, spelled out: sample 0 in the low bits (PlinkInputFile.h:206-209)
    assert pack_bed(np.array([[0, 1, 2, 0]]), np.array([[False, False, False, True]]))[0, 0] == (0 | 2 << 2 | 3 << 4 | 1 << 6)


def test_impute_mean_matches_known_dump(oracle):
    """test/correct.check1.GENE1.*.data convention: a missing call becomes 2 * p-hat"""
    G = np.array([[0, 1], [-9, 0], [1, -9], [0, 2], [0, 0]], dtype=float)
    out = oracle.impute_mean(G)
    assert out[1, 0] == pytest.approx(2 * (1 / 8))
    assert out[2, 1] == pytest.approx(2 * (3 / 8))
    all_missing = np.full((3, 1), -9.0)
    assert np.array_equal(oracle.impute_mean(all_missing), np.zeros((3, 1)))


@pytest.fixture(scope="module")
def eng(engine_cls):
    e = engine_cls(0)
    yield e
    e.close()


@pytest.mark.gpu
@pytest.mark.parametrize("case", [(60, 50, 3, 1), (61, 1001, 30, 3), (62, 70001, 50, 3), (63, 4096, 64, 2)])
def test_push_bed_equals_push_i8_bitwise(eng, oracle, case):
    from rvtests_b200.synth import pack_bed
    O = oracle
    seed, N, M, C = case
    maf = None if N >= 1000 else np.linspace(0.05, 0.4, M)
    G, X, y = make_problem(O, seed, N, M, C, maf=maf, n_flip=min(2, M - 1), n_mono=1 if M > 3 else 0)
    eng.set_null_model(X, y)
    af = af_of(G)
    bed = pack_bed(G.T)
    wide = np.zeros((M, bed.shape[1] + 5), dtype=np.uint8)      # a row stride larger than ceil(N/4)
    wide[:, : bed.shape[1]] = bed
    wide[:, bed.shape[1]:] = 0xFF
    eng.push_i8(G.T.copy(), af)
    eng.push_bed(bed, af)
    eng.push_bed(wide[:, : bed.shape[1]], af)
    r = eng.flush()
    assert r[0].tobytes() == r[1].tobytes() == r[2].tobytes()
    nm = O.fit_null_linear(X, y)
    ref, lam = O.gene(G.astype(float), af, X, nm["resid"], nm["sigma2"])
    check_gene(r[1], ref, lam, ctx=f"bed {case}")


@pytest.mark.gpu
@pytest.mark.parametrize("case", [(70, 700, 8, 1, 0.02), (71, 3001, 30, 3, 0.01)])
def test_push_bed_with_missing_calls_vs_oracle(eng, oracle, case):
    """01 codes -> imputeGenotypeToMean on the device -> fp64 path; the oracle decodes the same bytes with
    the reference's table and imputes with the reference's rule."""
    from rvtests_b200.synth import pack_bed
    O = oracle
    seed, N, M, C, miss = case
    G, X, y = make_problem(O, seed, N, M, C, maf=np.linspace(0.004, 0.3 if M <= 8 else 0.03, M), n_flip=2)
    rng = np.random.default_rng(seed)
    mask = rng.random((M, N)) < miss
    mask[0] = False                                              # one fully called variant
    bed = pack_bed(G.T, mask)
    eng.set_null_model(X, y)
    nm = O.fit_null_linear(X, y)
    raw = O.bed_decode_fast(bed, N).T                            # (N, M) with -9
    Gd = O.impute_mean(raw)
    # GenotypeCounter::getAF divides by nSample incl. missing (src/GenotypeCounter.h:46-52)
    af = 0.5 * np.where(raw >= 0, raw, 0.0).sum(axis=0) / N
    eng.push_bed(bed, af)
    eng.push_bed(pack_bed(G.T), af_of(G))                        # a fully called gene in the same flush
    r = eng.flush()
    ref, lam = O.gene(Gd, af, X, nm["resid"], nm["sigma2"])
    check_gene(r[0], ref, lam, ctx=f"bed+missing {case}")
    ref2, lam2 = O.gene(G.astype(float), af_of(G), X, nm["resid"], nm["sigma2"])
    check_gene(r[1], ref2, lam2, ctx=f"bed complete {case}")


@pytest.mark.gpu
def test_stream_batch_is_invisible_in_the_results(eng, oracle):
    """option stream_batch: kernels enqueued every B pushes (under the next H2D copies) instead of at
    flush -- same records, same order, also when the staging arena grows in between and when a gene
    with missing calls sits in an already-launched range"""
    from rvtests_b200.synth import pack_bed
    O = oracle
    N, C = 20011, 3
    rng = np.random.default_rng(81)
    Ms = [5, 50, 64, 1, 33, 62, 63, 17, 50, 8, 40]
    X, y = O.synth_covariates(81, N, C)
    eng.set_null_model(X, y)
    genes = []
    for g, M in enumerate(Ms):
        G, _, _ = make_problem(O, 810 + g, N, M, C, n_flip=1 if M > 1 else 0)
        mask = None
        if g == 2:
            mask = rng.random((M, N)) < 0.01
        genes.append((G, mask))

    def run(batch):
        eng.set_option("stream_batch", batch)
        try:
            for g, (G, mask) in enumerate(genes):
                if g % 3 == 1:
                    eng.push_i8(G.T.copy(), af_of(G))
                else:
                    eng.push_bed(pack_bed(G.T, mask), af_of(G))
            return eng.flush()
        finally:
            eng.set_option("stream_batch", 0)

    base = run(0)
    assert len(base) == len(Ms) and np.all(base["status"] == 0)
    hard = [g for g in range(len(Ms)) if g != 2]
    for batch in (1, 3, 4, 64):
        r = run(batch)
        # exact integer pipeline: identical bits whatever the batching
        assert r[hard].tobytes() == base[hard].tobytes(), batch
        # gene 2 took the fp64 dosage path (floating-point atomics: last-bit run-to-run noise)
        for k in ("Q", "p_skat", "cmc_p", "zeg_p", "lambda_max"):
            assert abs(r[2][k] - base[2][k]) <= 1e-10 * abs(base[2][k]), (batch, k)
        assert int(r[2]["cmc_nonref"]) == int(base[2]["cmc_nonref"])
