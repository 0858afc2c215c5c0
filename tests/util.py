"""Shared helpers for the parity tests: seeded problem construction + comparison with the oracle."""
import numpy as np


def make_problem(O, seed, N, M, C, maf=None, n_mono=0, n_flip=0):
    """Genotypes (N, M) int8 from the SURVEY 8(d) stream, covariates and a null trait.
    n_flip columns get MAF replaced by 1-maf' (ALT is the major allele => flip-to-minor);
    n_mono columns are made monomorphic (all 0, all 2, or all 1 in turn)."""
    vid = np.arange(M, dtype=np.uint64) + np.uint64(seed * 1000003)
    m = O.synth_maf(seed, vid) if maf is None else np.broadcast_to(np.asarray(maf, dtype=float), (M,)).copy()
    rng = np.random.default_rng(seed)
    cols = rng.permutation(M)
    for j in cols[:n_flip]:
        m[j] = 1.0 - rng.uniform(0.01, 0.3)
    G = O.synth_genotypes(seed, vid, N, maf=m).T.copy()  # (N, M)
    for t, j in enumerate(cols[n_flip:n_flip + n_mono]):
        G[:, j] = (0, 2, 1)[t % 3]
    X, y = O.synth_covariates(seed, N, C)
    return G.astype(np.int8), X, y


def af_of(G):
    """GenotypeCounter::getAF for complete hard calls (src/GenotypeCounter.h:46-52)."""
    return 0.5 * G.astype(np.float64).sum(axis=0) / G.shape[0]


def rel(a, b):
    a, b = float(a), float(b)
    if a == b:
        return 0.0
    return abs(a - b) / max(abs(a), abs(b), 1e-300)


def check_gene(res, ref, lam_ref, tol_q=1e-6, tol_p=1e-4, ctx=""):
    """res: one record of RESULT_DTYPE from the GPU; ref: oracle GeneOut."""
    assert int(res["m_poly"]) == ref.m_poly, f"{ctx} m_poly {res['m_poly']} vs {ref.m_poly}"
    if ref.status == 2:
        assert int(res["status"]) == 2, f"{ctx} status"
        return
    assert int(res["status"]) == 0, f"{ctx} status {res['status']}"
    # integer burden statistics: bit exact
    assert int(res["cmc_nonref"]) == ref.cmc_nonref, f"{ctx} NonRefSite {res['cmc_nonref']} vs {ref.cmc_nonref}"
    assert rel(res["Q"], ref.skat.Q) <= tol_q, f"{ctx} Q {res['Q']} vs {ref.skat.Q}"
    if ref.skat.n_lambda > 0 and ref.skat.Q > 0:
        assert rel(res["lambda_max"], lam_ref[0]) <= 1e-8, f"{ctx} lambda_max"
        assert int(res["davies_fault"]) == ref.skat.fault, f"{ctx} fault {res['davies_fault']} vs {ref.skat.fault}"
        assert rel(res["p_skat"], ref.skat.pvalue) <= tol_p, f"{ctx} p {res['p_skat']} vs {ref.skat.pvalue}"
        assert rel(res["p_liu"], ref.skat.p_liu) <= tol_p, f"{ctx} p_liu"
    for pre in ("cmc", "zeg"):
        ok = getattr(ref, pre + "_ok")
        assert int(res[pre + "_ok"]) == ok, f"{ctx} {pre}_ok"
        if ok:
            # U ~ N(0, V): judge the error on the scale of its standard deviation (U itself may be ~0)
            u_ref, v_ref = getattr(ref, pre + "_U"), getattr(ref, pre + "_V")
            assert abs(res[pre + "_U"] - u_ref) <= 1e-6 * max(abs(u_ref), np.sqrt(abs(v_ref))), f"{ctx} {pre}_U {res[pre + '_U']} vs {u_ref}"
            assert rel(res[pre + "_V"], getattr(ref, pre + "_V")) <= 1e-6, f"{ctx} {pre}_V"
            assert rel(res[pre + "_p"], getattr(ref, pre + "_p")) <= tol_p, f"{ctx} {pre}_p {res[pre + '_p']} vs {getattr(ref, pre + '_p')}"
