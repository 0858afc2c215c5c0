"""GPU: the C++ ModelFitter-shaped adapters (rvtests_b200/host/rvt_fitters.h) driven like the
reference's gene loop (src/Main.cpp:1221-1254) produce the .assoc lines the reference's own
writeOutput would print ("%g" columns) for the oracle's numbers."""
import os
import struct
import subprocess

import numpy as np
import pytest

from util import af_of, make_problem

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_demo():
    exe = os.path.join(ROOT, "rvtests_b200", "host", "adapter_demo")
    src = os.path.join(ROOT, "rvtests_b200", "host", "adapter_demo.cpp")
    subprocess.run(["g++", "-std=c++11", "-O2", "-pthread", "-I", os.path.join(ROOT, "include"), src, "-o", exe,
                    "-L", os.path.join(ROOT, "rvtests_b200"), "-lrvtests_b200",
                    "-Wl,-rpath," + os.path.join(ROOT, "rvtests_b200")], check=True)
    return exe


def g(v):
    return "%g" % v


@pytest.mark.parametrize("batch", [1, 3, 100])
def test_adapters_match_reference_output_format(oracle, batch, tmp_path):
    from oracle import skato_oracle as SO
    import rvtests_b200
    rvtests_b200.load_library()
    O = oracle
    N, C = 1200, 3
    genes = []
    X = y = None
    for gi, (M, nm_, nf) in enumerate([(6, 0, 1), (1, 0, 0), (25, 2, 2), (4, 4, 0), (40, 1, 0)]):
        # rare variants: a constant CMC indicator (every sample a carrier) would make the burden test undefined
        G, X, y = make_problem(O, 77, N, M, C, maf=np.linspace(0.002, 0.03, M), n_mono=nm_, n_flip=nf)
        rng = np.random.default_rng(gi)
        G = G[:, rng.permutation(M)]
        genes.append(G)
    path = tmp_path / "problem.bin"
    with open(path, "wb") as f:
        f.write(struct.pack("iii", N, C - 1, len(genes)))
        f.write(np.ascontiguousarray(y).tobytes())
        f.write(np.asfortranarray(X[:, 1:]).tobytes(order="F"))
        for G in genes:
            f.write(struct.pack("i", G.shape[1]))
            f.write(np.asfortranarray(G.astype(np.float64)).tobytes(order="F"))
            f.write(af_of(G).tobytes())
    exe = build_demo()
    out = subprocess.run([exe, str(path), str(batch)], capture_output=True, text=True, check=True).stdout
    tables = {}
    cur = None
    for line in out.splitlines():
        if line.startswith("#"):
            cur = line[1:]
            tables[cur] = []
        else:
            tables[cur].append(line.split("\t"))
    assert tables["Skat"][0] == ["Range", "N_INFORMATIVE", "NumVar", "Q", "Pvalue"]
    assert tables["SkatO"][0][-3:] == ["Q", "rho", "Pvalue"]
    assert tables["CMC"][0][-2:] == ["NonRefSite", "Pvalue"]
    assert tables["Zeggini"][0][-1:] == ["Pvalue"]
    nm = O.fit_null_linear(X, y)
    for gi, G in enumerate(genes):
        ref, lam = O.gene(G.astype(float), af_of(G), X, nm["resid"], nm["sigma2"])
        site = ["gene%d" % gi, str(N), str(G.shape[1])]
        rs, ro, rc, rz = (tables[k][1 + gi] for k in ("Skat", "SkatO", "CMC", "Zeggini"))
        assert rs[:3] == site and ro[:3] == site and rc[:3] == site and rz[:3] == site
        if ref.status == 2:
            assert rs[3:] == ["NA", "NA"] and ro[3:] == ["NA", "NA", "NA"] and rc[3:] == ["NA", "NA"] and rz[3:] == ["NA"]
            continue
        assert rs[3:] == [g(ref.skat.Q), g(ref.skat.pvalue)]
        assert rc[3:] == [str(ref.cmc_nonref), g(ref.cmc_p)]
        assert rz[3:] == [g(ref.zeg_p)]
        so = SO.skato_gene(G.astype(float), af_of(G), X, nm["resid"])
        assert ro[3:] == [g(so["Q"]), g(so["rho"]), g(so["pvalue"])]


def test_skat_adapter_with_permutations(oracle, tmp_path):
    """`--kernel skat[nPerm=300,alpha=0.1]`: the eight columns SkatTest::writeOutput prints with permutations
    (src/Model.h:2722-2751, src/Permutation.h:118-139), against the oracle's rand()-driven loop."""
    import rvtests_b200
    rvtests_b200.load_library()
    O = oracle
    N, C, n_perm, alpha = 900, 2, 300, 0.1
    genes = []
    X = y = None
    for gi, (M, nm_, nf) in enumerate([(6, 0, 1), (3, 3, 0), (25, 2, 2)]):
        G, X, y = make_problem(O, 78, N, M, C, maf=np.linspace(0.004, 0.05, M), n_mono=nm_, n_flip=nf)
        genes.append(G)
    path = tmp_path / "problem.bin"
    with open(path, "wb") as f:
        f.write(struct.pack("iii", N, C - 1, len(genes)))
        f.write(np.ascontiguousarray(y).tobytes())
        f.write(np.asfortranarray(X[:, 1:]).tobytes(order="F"))
        for G in genes:
            f.write(struct.pack("i", G.shape[1]))
            f.write(np.asfortranarray(G.astype(np.float64)).tobytes(order="F"))
            f.write(af_of(G).tobytes())
    exe = build_demo()
    out = subprocess.run([exe, str(path), "2", str(n_perm), str(alpha)], capture_output=True, text=True, check=True).stdout
    rows = [l.split("\t") for l in out.split("#SkatO")[0].splitlines()[1:]]
    assert rows[0][3:] == ["Q", "Pvalue", "NumPerm", "ActualPerm", "Stat", "NumGreater", "NumEqual", "PermPvalue"]
    nm = O.fit_null_linear(X, y)
    first = True
    for gi, G in enumerate(genes):
        ref, lam = O.gene(G.astype(float), af_of(G), X, nm["resid"], nm["sigma2"])
        row = rows[1 + gi][3:]
        if ref.status == 2:
            assert row == ["NA"] * 8
            continue
        pr = O.gene_perm(G.astype(float), af_of(G), nm["resid"], ref.skat.Q, n_perm=n_perm, alpha=alpha, reseed=1 if first else 0)
        first = False
        assert row == [g(ref.skat.Q), g(ref.skat.pvalue), str(n_perm), str(pr["actual"]), g(ref.skat.Q), str(pr["greater"]),
                       str(pr["equal"]), g(pr["p"])]


def build_meta_demo():
    exe = os.path.join(ROOT, "rvtests_b200", "host", "meta_demo")
    src = os.path.join(ROOT, "rvtests_b200", "host", "meta_demo.cpp")
    subprocess.run(["g++", "-std=c++11", "-O2", "-pthread", "-I", os.path.join(ROOT, "include"), src, "-o", exe,
                    "-L", os.path.join(ROOT, "rvtests_b200"), "-lrvtests_b200", "-lz",
                    "-Wl,-rpath," + os.path.join(ROOT, "rvtests_b200")], check=True)
    return exe


@pytest.mark.parametrize("segment", [64, 100, 100000])
def test_meta_adapters_match_reference_output_format(oracle, segment, tmp_path):
    """MetaScoreTest / MetaCovTest adapters (host/rvt_meta_fitters.h) driven variant by variant: the .assoc lines the
    reference's writeOutput / printCovariance would print (src/Model.h:3282-3345, src/Model.cpp:942-1004) for the oracle's
    numbers -- whatever the segment size (windows must survive segment boundaries and close at a chromosome change)."""
    from oracle import meta_oracle as MO
    import rvtests_b200
    rvtests_b200.load_library()
    O = oracle
    seed, N, nv, C, window = 5, 1500, 230, 3, 900
    vid = np.arange(nv, dtype=np.uint64) + np.uint64(seed * 7919)
    rng = np.random.default_rng(seed)
    maf = 10 ** rng.uniform(np.log10(2.0 / N), np.log10(0.4), nv)
    G = O.synth_genotypes(seed, vid, N, maf=maf)                # (nv, N)
    for j in rng.integers(0, nv, 8):
        G[j] = 0                                                # monomorphic sites: NA statistics, no MetaCov line
    X, y = O.synth_covariates(seed, N, C)
    pos = np.cumsum(rng.integers(1, 40, nv)).astype(np.int32)
    chrom = np.ones(nv, dtype=np.int32)
    cut = int(nv * 0.6)
    chrom[cut:] = 2
    pos[cut:] -= pos[cut] - 7
    path = tmp_path / "meta.bin"
    with open(path, "wb") as f:
        f.write(struct.pack("iii", N, C - 1, nv))
        f.write(np.ascontiguousarray(y).tobytes())
        f.write(np.asfortranarray(X[:, 1:]).tobytes(order="F"))
        for v in range(nv):
            f.write(struct.pack("ii", int(chrom[v]), int(pos[v])))
            f.write(G[v].astype(np.float64).tobytes())
    exe = build_meta_demo()
    prefix = str(tmp_path / "out")
    out = subprocess.run([exe, str(path), str(segment), str(window), prefix], capture_output=True, text=True, check=True).stdout
    score_txt, cov_txt = out.split("#MetaCov\n")
    # the bgzipped + tabix-indexed files ModelManager would leave behind (rvt_bgzf.h; format checks in tests/test_bgzf_tabix.py)
    import gzip
    assert gzip.open(prefix + ".MetaScore.assoc.gz", "rt").read() == score_txt.split("#MetaScore\n", 1)[1]
    assert gzip.open(prefix + ".MetaCov.assoc.gz", "rt").read() == cov_txt
    for m in ("MetaScore", "MetaCov"):
        tbi = gzip.open(prefix + "." + m + ".assoc.gz.tbi", "rb").read()
        # two chromosomes + the column header line, which carries no '#' and which tabix therefore indexes as a sequence
        # named CHROM -- in the reference's own files too
        assert tbi[:4] == b"TBI\x01" and struct.unpack_from("<i", tbi, 4)[0] == 3
        assert tbi[36:].startswith(b"CHROM\x001\x002\x00")
        assert struct.unpack_from("<6i", tbi, 8) == (0, 1, 2, 0, ord("#"), 0)
    score_lines = score_txt.splitlines()[1:]
    nm = O.fit_null_linear(X, y)
    # ##NullModelEstimates block, then the column header
    assert score_lines[0] == "##NullModelEstimates" and score_lines[1] == "## - Name\tBeta\tSD"
    beta = np.linalg.solve(X.T @ X, X.T @ y)
    var = np.diag(nm["xtx_inv"]) * nm["sigma2"]
    assert score_lines[2] == "## - Intercept\t%s\t%s" % (g(beta[0]), g(var[0]))
    assert score_lines[2 + C] == "## - Sigma2\t%s\tNA" % g(nm["sigma2"])
    hdr = score_lines[3 + C].split("\t")
    assert hdr == ["CHROM", "POS", "REF", "ALT", "N_INFORMATIVE", "AF", "INFORMATIVE_ALT_AC", "CALL_RATE", "HWE_PVALUE", "N_REF",
                   "N_HET", "N_ALT", "U_STAT", "SQRT_V_STAT", "ALT_EFFSIZE", "ALT_EFFSIZE_SE", "PVALUE"]
    rows = [l.split("\t") for l in score_lines[4 + C:]]
    assert len(rows) == nv
    for v in range(nv):
        ref = MO.meta_score(G[v].astype(np.float64), X, nm["resid"], nm["sigma2"])
        exp = [str(chrom[v]), str(pos[v]), "A", "G", str(N), g(ref["af"]), g(ref["ac"]), "1", g(ref["hwe_p"]), str(ref["n_ref"]),
               str(ref["n_het"]), str(ref["n_alt"])]
        exp += [g(ref["U"]), g(ref["sqrtV"]), g(ref["effect"]), g(ref["effect_se"]), g(ref["pvalue"])] if ref["ok"] else ["NA"] * 5
        assert rows[v] == exp, (v, rows[v], exp)
    # MetaCov: one line per polymorphic variant, in order
    cov_lines = cov_txt.splitlines()
    assert cov_lines[0] == "CHROM\tSTART_POS\tEND_POS\tNUM_MARKER\tMARKER_POS\tCOV"
    ref_cov = [(v, rc) for v, rc in enumerate(MO.meta_cov(G.T, pos, chrom, X, nm["sigma2"], window)) if rc is not None]
    assert len(cov_lines) - 1 == len(ref_cov)
    for line, (v, (ps, vals)) in zip(cov_lines[1:], ref_cov):
        c = line.split("\t")
        assert c[:5] == [str(chrom[v]), str(pos[v]), str(ps[-1]), str(len(ps)), ",".join(map(str, ps))], (v, c[:5])
        got = np.array([float(x) for x in c[5].split(",")])
        assert len(got) == len(vals)
        assert np.all(np.abs(got - np.array(vals)) <= 2e-5 * np.maximum(np.abs(vals), 1e-12) + 1e-12)   # printed with 6 digits


def test_adapters_binary_outcome(oracle, tmp_path):
    """setBinaryOutcome(): the adapters switch to the logistic null model; Skat / CMC / Zeggini lines against the binary
    oracle; with enableSkatOBinary(false) (adapter_demo argv[6] = 0) SkatO prints NA -- the default prints the type "D"
    columns (tests/test_gpu_zz_fp64_skato.py::test_adapter_prints_skato_for_a_binary_trait)."""
    from oracle import binary_oracle as BIN
    import rvtests_b200
    rvtests_b200.load_library()
    O = oracle
    N, C = 1500, 2
    genes = []
    X = None
    for gi, (M, nm_, nf) in enumerate([(8, 0, 1), (30, 2, 2)]):
        G, X, _ = make_problem(O, 79, N, M, C, maf=np.linspace(0.004, 0.05, M), n_mono=nm_, n_flip=nf)
        genes.append(G)
    rng = np.random.default_rng(79)
    y = (rng.random(N) < 1.0 / (1.0 + np.exp(0.5 - 0.6 * X[:, 1]))).astype(np.float64)
    path = tmp_path / "problem.bin"
    with open(path, "wb") as f:
        f.write(struct.pack("iii", N, C - 1, len(genes)))
        f.write(np.ascontiguousarray(y).tobytes())
        f.write(np.asfortranarray(X[:, 1:]).tobytes(order="F"))
        for G in genes:
            f.write(struct.pack("i", G.shape[1]))
            f.write(np.asfortranarray(G.astype(np.float64)).tobytes(order="F"))
            f.write(af_of(G).tobytes())
    exe = build_demo()
    out = subprocess.run([exe, str(path), "8", "0", "0.05", "1", "0"], capture_output=True, text=True, check=True).stdout
    tables, cur = {}, None
    for line in out.splitlines():
        if line.startswith("#"):
            cur = line[1:]
            tables[cur] = []
        else:
            tables[cur].append(line.split("\t"))
    nm = BIN.fit_null_logistic(X, y)

    def close(txt, val, tol=2e-5):
        return txt != "NA" and abs(float(txt) - val) <= tol * max(abs(val), 1e-300)

    for gi, G in enumerate(genes):
        ref = BIN.gene(G.astype(float), af_of(G), X, nm)
        rs, ro, rc, rz = (tables[k][1 + gi][3:] for k in ("Skat", "SkatO", "CMC", "Zeggini"))
        ctx = (gi, rs, ro, rc, rz, ref["Q"], ref["p_skat"], ref["cmc"], ref["zeg"])
        # "%g" prints 6 digits; the statistics agree to ~1e-9, so compare the printed numbers numerically
        assert close(rs[0], ref["Q"]) and close(rs[1], ref["p_skat"]), ctx
        assert ro == ["NA", "NA", "NA"], ctx
        assert rc[0] == str(ref["cmc"]["nonref"]) and close(rc[1], ref["cmc"]["p"]), ctx
        assert close(rz[0], ref["zeg"]["p"]), ctx


def test_adapters_vs_the_reference_model_layer_text(tmp_path):
    """The C++ adapters on the device against the `.assoc` lines that the REFERENCE's own model layer printed for the same
    five genes (tests/golden/ref_assoc_golden.npz: src/Model.cpp + Model.h fitters + DataConsolidator.cpp compiled
    unmodified in the build container, tests/golden/make_golden_ref_assoc.py).  No oracle in between."""
    import rvtests_b200
    from test_oracle_pin_reference_model_layer import load_assoc_golden, model_columns_match
    rvtests_b200.load_library()
    X, y, genes, text = load_assoc_golden()
    N = len(y)
    path = tmp_path / "problem.bin"
    with open(path, "wb") as f:
        f.write(struct.pack("iii", N, X.shape[1] - 1, len(genes)))
        f.write(np.ascontiguousarray(y).tobytes())
        f.write(np.asfortranarray(X[:, 1:]).tobytes(order="F"))
        for G in genes:
            f.write(struct.pack("i", G.shape[1]))
            f.write(np.asfortranarray(G.astype(np.float64)).tobytes(order="F"))
            f.write(af_of(G).tobytes())
    exe = build_demo()
    out = subprocess.run([exe, str(path), "3"], capture_output=True, text=True, check=True).stdout
    tables, cur = {}, None
    for line in out.splitlines():
        if line.startswith("#"):
            cur = line[1:]
            tables[cur] = []
        else:
            tables[cur].append(line.split("\t"))
    for m in ("Skat", "SkatO", "CMC", "Zeggini"):
        w = {"Skat": 2, "SkatO": 3, "CMC": 2, "Zeggini": 1}[m]
        assert tables[m][0][-w:] == text[m]["header"][-w:], (m, tables[m][0], text[m]["header"])
    for k in range(len(genes)):
        model_columns_match({m: text[m]["rows"][k] for m in text}, {m: tables[m][1 + k] for m in text})
