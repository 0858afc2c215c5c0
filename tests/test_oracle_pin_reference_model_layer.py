"""CPU: the oracle against the reference's own MODEL LAYER -- src/Model.cpp + the fitters of src/Model.h (SkatTest, SkatOTest,
CMCTest, ZegginiTest, MetaScoreTest, MetaCovTest), src/DataConsolidator.cpp (mean imputation, flip to the minor allele,
monomorphic drop, the AF table with its index quirk), the collapsing functions and the `.assoc` writers -- compiled
unmodified into oracle/_ref/libmodel_ref.so (oracle/Makefile, oracle/ref_model_shim.cpp; Eigen replaced by
oracle/eigen_standin) and driven like src/Main.cpp.  What is compared is the text the reference prints ("%g", 6 digits)."""
import numpy as np
import pytest

from util import make_problem, rel


def _g(x):
    return "%g" % x


def _close(txt, val, tol):
    """a printed "%g" field against a double: equal as text, or within tol (float32 fields, last-digit rounding)"""
    if txt == _g(val):
        return True
    return txt != "NA" and rel(float(txt), val) <= tol


def _af_ref(g):
    """GenotypeCounter::getAF: 0.5 * sum of called dosages / ALL samples (src/GenotypeCounter.h:46-52)"""
    return np.where(g < 0, 0.0, g).sum(axis=0) / 2.0 / g.shape[0]


@pytest.fixture(scope="module")
def ref(oracle):
    if oracle.ref_model() is None:
        pytest.skip("oracle/_ref/libmodel_ref.so not built (no /root/reference here)")
    return oracle


def test_gene_loop_skat_skato_cmc_zeggini(ref, tmp_path):
    from oracle import skato_oracle as SO
    O = ref
    N = 900
    specs = [(12, 30, None, 2, 3), (13, 7, 0.3, 0, 1), (14, 50, None, 3, 4), (15, 1, 0.2, 0, 0), (16, 6, None, 6, 0),
             (17, 12, 0.3, 1, 2)]
    genes = []
    X = y = None
    for seed, M, hi, n_mono, n_flip in specs:
        maf = None if hi is None else (np.linspace(0.004, hi, M) if M > 1 else np.array([hi]))
        G, Xk, yk = make_problem(O, seed, N, M, 3, maf=maf, n_mono=n_mono, n_flip=n_flip)
        if X is None:
            X, y = Xk, yk
        genes.append(G.astype(float))
    rng = np.random.default_rng(0)
    for k in (1, 5):  # missing calls -> mean imputation (A0) and the AF denominator that counts missing samples
        for _ in range(9):
            genes[k][rng.integers(N), rng.integers(genes[k].shape[1])] = -9.0
    out = O.ref_run_gene_models(genes, X[:, 1:], y, str(tmp_path / "g"))
    nm = O.fit_null_linear(X, y)
    assert out["Skat"][1][-2:] == ["Q", "Pvalue"] and out["SkatO"][1][-3:] == ["Q", "rho", "Pvalue"]
    assert out["CMC"][1][-2:] == ["NonRefSite", "Pvalue"] and out["Zeggini"][1][-1:] == ["Pvalue"]
    for k, g in enumerate(genes):
        gi = O.impute_mean(g) if (g < 0).any() else g
        o, lam = O.gene(gi, _af_ref(g), X, nm["resid"], nm["sigma2"])
        so = SO.skato_gene(gi, _af_ref(g), X, nm["resid"])
        sk, sko, cmc, zeg = (out[m][2][k] for m in ("Skat", "SkatO", "CMC", "Zeggini"))
        assert int(sk[4]) == o.m_poly, (k, sk)                       # NumPolyVar: flip + monomorphic drop (A1)
        if o.m_poly == 0:
            assert sk[-2:] == ["NA", "NA"] and sko[-3:] == ["NA", "NA", "NA"] and cmc[-2:] == ["NA", "NA"] and zeg[-1] == "NA"
            continue
        assert _close(sk[-2], o.skat.Q, 3e-6) and _close(sk[-1], o.skat.pvalue, 5e-4), (k, sk, o.skat.Q, o.skat.pvalue)
        assert _close(sko[-3], so["Q"], 1e-5) and float(sko[-2]) == so["rho"] and _close(sko[-1], so["pvalue"], 1e-5), (k, sko, so)
        assert int(cmc[-2]) == o.cmc_nonref and _close(cmc[-1], o.cmc_p, 1e-5), (k, cmc)
        assert _close(zeg[-1], o.zeg_p, 1e-5), (k, zeg)


def test_gene_loop_permutation_columns(ref, tmp_path):
    """skat[nPerm=..]: NumPerm ActualPerm Stat NumGreater NumEqual PermPvalue (src/Permutation.h) over two genes that
    continue one rand() stream.  The reference build runs in this process, so both sides restart the stream."""
    import ctypes as C
    O = ref
    N = 400
    genes = [make_problem(O, s, N, M, 2, maf=np.linspace(0.01, 0.3, M))[0].astype(float) for s, M in ((31, 8), (32, 5))]
    _, X, y = make_problem(O, 31, N, 8, 2)
    libc = C.CDLL(None)
    libc.srand(1)
    out = O.ref_run_gene_models(genes, X[:, 1:], y, str(tmp_path / "p"), n_perm=300, alpha=0.05)
    nm = O.fit_null_linear(X, y)
    hdr = out["Skat"][1]
    assert hdr[-8:] == ["Q", "Pvalue", "NumPerm", "ActualPerm", "Stat", "NumGreater", "NumEqual", "PermPvalue"]
    for k, g in enumerate(genes):
        o, _ = O.gene(g, _af_ref(g), X, nm["resid"], nm["sigma2"])
        pr = O.gene_perm(g, _af_ref(g), nm["resid"], o.skat.Q, n_perm=300, alpha=0.05, reseed=1 if k == 0 else 0)
        row = out["Skat"][2][k]
        assert int(row[-6]) == 300 and int(row[-5]) == pr["actual"], (k, row, pr["actual"])
        assert (int(row[-3]), int(row[-2])) == (pr["greater"], pr["equal"]), (k, row)
        assert _close(row[-1], pr["p"], 1e-6)


def test_single_variant_loop_meta_score_and_cov(ref, tmp_path):
    from oracle import meta_oracle as MO
    O = ref
    N, nv, window = 700, 14, 10000
    G, X, y = make_problem(O, 21, N, nv, 3, maf=np.r_[np.linspace(0.01, 0.45, nv - 2), [0.8, 0.97]], n_mono=1)
    pos = np.array([100, 200, 300, 5000, 5100, 5200, 15100, 90000, 90010, 90020, 200000, 200001, 300000, 300500], dtype=np.int32)
    out = O.ref_run_meta_models(G.astype(float), pos, X[:, 1:], y, window, str(tmp_path / "m"))
    nm = O.fit_null_linear(X, y)
    com, hdr, rows = out["MetaScore"]
    assert hdr == ["CHROM", "POS", "REF", "ALT", "N_INFORMATIVE", "AF", "INFORMATIVE_ALT_AC", "CALL_RATE", "HWE_PVALUE", "N_REF",
                   "N_HET", "N_ALT", "U_STAT", "SQRT_V_STAT", "ALT_EFFSIZE", "PVALUE"]
    # the ##NullModelEstimates block: beta and the DIAGONAL OF covB (printed under the name "SD"), sigma2
    lines = {l.split("\t")[0]: l.split("\t") for l in com if l.startswith("## - ")}
    covB = nm["xtx_inv"] * nm["sigma2"]
    for i, name in enumerate(["## - Intercept", "## - cov0", "## - cov1"]):
        assert _close(lines[name][1], nm["beta"][i], 1e-5) and _close(lines[name][2], covB[i, i], 1e-5), lines[name]
    assert _close(lines["## - Sigma2"][1], nm["sigma2"], 1e-5) and lines["## - Sigma2"][2] == "NA"
    assert len(rows) == nv
    for j, row in enumerate(rows):
        o = MO.meta_score(G[:, j].astype(float), X, nm["resid"], nm["sigma2"])
        assert row[5:12] == [_g(o["af"]), _g(o["ac"]), _g(o["call_rate"]), _g(o["hwe_p"]), str(o["n_ref"]), str(o["n_het"]),
                             str(o["n_alt"])], (j, row)
        if not o.get("ok"):
            assert row[12:] == ["NA"] * 4, (j, row)
            continue
        for txt, key in zip(row[12:], ("U", "sqrtV", "effect", "pvalue")):
            assert _close(txt, o[key], 1e-5), (j, key, txt, o[key])
    # covariance lines: one per polymorphic variant, in order, window by position.  With covariates the reference's
    # computeQuadraticForm (src/Model.h:3993-4010) multiplies the transposes of two 1 x C maps with the C x C inverse --
    # shapes that do not conform, i.e. out-of-contract Eigen usage whose value depends on Eigen's internals (asserts are
    # compiled out) and that no stand-in can pin.  The window logic does not depend on it; the values are pinned on an
    # intercept-only null model below, where every shape is 1 x 1.
    com, hdr, rows = out["MetaCov"]
    assert hdr == ["CHROM", "START_POS", "END_POS", "NUM_MARKER", "MARKER_POS", "COV"]
    want = [w for w in MO.meta_cov(G.astype(float), pos, ["1"] * nv, X, nm["sigma2"], window) if w is not None]
    assert len(rows) == len(want)
    for row, (ps, vals) in zip(rows, want):
        assert [int(p) for p in row[4].split(",")] == ps, row
        assert (int(row[1]), int(row[2]), int(row[3])) == (ps[0], ps[-1], len(ps))
        assert len(row[5].split(",")) == len(vals)

    X1 = X[:, :1]
    out1 = O.ref_run_meta_models(G.astype(float), pos, np.zeros((N, 0)), y, window, str(tmp_path / "m1"))
    nm1 = O.fit_null_linear(X1, y)
    rows = out1["MetaCov"][2]
    want = [w for w in MO.meta_cov(G.astype(float), pos, ["1"] * nv, X1, nm1["sigma2"], window) if w is not None]
    assert len(rows) == len(want) and len(rows) == nv - 1
    for row, (ps, vals) in zip(rows, want):
        assert [int(p) for p in row[4].split(",")] == ps, row
        got = np.array([float(v) for v in row[5].split(",")])
        # the reference holds genotypes and sigma2 in float here (MetaCov entries <= 1e-5 rel); off-diagonal
        # entries are judged on the scale of the head variant's variance
        assert np.max(np.abs(got - np.array(vals))) <= 1e-5 * vals[0], (row, vals)
    for j, row in enumerate(out1["MetaScore"][2]):
        o = MO.meta_score(G[:, j].astype(float), X1, nm1["resid"], nm1["sigma2"])
        if o.get("ok"):
            for txt, key in zip(row[12:], ("U", "sqrtV", "effect", "pvalue")):
                assert _close(txt, o[key], 1e-5), (j, key, txt, o[key])


def model_columns_match(ref_rows, got, tol_skat_q=3e-6, tol_skat_p=5e-4, tol=1e-5):
    """ref_rows: {model: row of the reference's .assoc text}; got: {model: the same model columns from another
    implementation, as text}.  Model columns only: Skat Q,Pvalue | SkatO Q,rho,Pvalue | CMC NonRefSite,Pvalue | Zeggini Pvalue."""
    width = {"Skat": 2, "SkatO": 3, "CMC": 2, "Zeggini": 1}
    for m, w in width.items():
        r, o = ref_rows[m][-w:], got[m][-w:]
        if "NA" in r:
            assert o == r, (m, r, o)
            continue
        for i, (a, b) in enumerate(zip(r, o)):
            if a == b:
                continue
            if (m, i) in (("CMC", 0), ("SkatO", 1)):        # NonRefSite and rho: exact
                assert float(a) == float(b), (m, r, o)
            else:
                t = (tol_skat_q if i == 0 else tol_skat_p) if m == "Skat" else tol
                assert rel(float(a), float(b)) <= t + 1e-6, (m, r, o)   # + the rounding of "%g" itself


def load_assoc_golden():
    import json
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_assoc_golden.npz"))
    text = json.loads(str(g["assoc"]))
    genes = [g[f"G{k}"] for k in range(len(text["Skat"]["rows"]))]
    return g["X"], g["y"], genes, text


def test_golden_assoc_text_vs_oracle(oracle):
    """tests/golden/ref_assoc_golden.npz: the `.assoc` lines the reference's model layer printed for five genes
    (tests/golden/make_golden_ref_assoc.py) against the oracle's numbers printed with "%g" -- runs without oracle/_ref.
    The same file is what tests/test_gpu_adapters.py holds the C++ adapters on the device against."""
    from oracle import skato_oracle as SO
    O = oracle
    X, y, genes, text = load_assoc_golden()
    nm = O.fit_null_linear(X, y)
    for k, G in enumerate(genes):
        o, _ = O.gene(G.astype(float), _af_ref(G.astype(float)), X, nm["resid"], nm["sigma2"])
        ref_rows = {m: text[m]["rows"][k] for m in text}
        if o.status == 2:
            got = {"Skat": ["NA", "NA"], "SkatO": ["NA"] * 3, "CMC": ["NA", "NA"], "Zeggini": ["NA"]}
        else:
            so = SO.skato_gene(G.astype(float), _af_ref(G.astype(float)), X, nm["resid"])
            got = {"Skat": [_g(o.skat.Q), _g(o.skat.pvalue)], "SkatO": [_g(so["Q"]), _g(so["rho"]), _g(so["pvalue"])],
                   "CMC": [str(o.cmc_nonref), _g(o.cmc_p)], "Zeggini": [_g(o.zeg_p)]}
        model_columns_match(ref_rows, got)
