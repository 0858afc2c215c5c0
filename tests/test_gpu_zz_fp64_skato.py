"""GPU (-m gpu): SKAT-O on genes that take the fp64 path (pre-digested statistics handed to k_finalize<true>):
  * quantitative trait, genes with missing calls (mean-imputed on the device, A0) and dosage genes -- SkatO::Fit type "C";
  * binary trait (option "skato_binary") -- SkatO::Fit type "D" (src/Model.h:2833-2841, 2854-2858; SkatO.cpp:72-91,
    133-134, 150-158): the same tail on the p(1-p)-weighted statistics with s2 = 1.
The oracle (oracle/skato_oracle.py) is pinned on the reference's own SkatO.cpp for both types
(tests/test_oracle_pin_reference_skat.py::test_live_reference_build, ::test_live_binary_skato).
This file sorts last on purpose: it was written after the round's GPU budget was spent, so the
rest of the suite runs before it."""
import numpy as np
import pytest

from util import af_of, check_gene, make_problem, rel

pytestmark = pytest.mark.gpu


def _check_skato(r, ref, ctx):
    assert int(r["skato_ok"]) == int(ref["ok"]), ctx
    if ref["ok"]:
        assert rel(r["skato_Q"], ref["Q"]) <= 1e-6, (ctx, r["skato_Q"], ref["Q"])
        assert r["skato_rho"] == ref["rho"], (ctx, r["skato_rho"], ref["rho"])
        assert rel(r["skato_p"], ref["pvalue"]) <= 1e-5, (ctx, r["skato_p"], ref["pvalue"])


def test_vcf_text_to_engine(engine_cls, oracle, vcfpack):
    """SURVEY 8(f) N2 end to end: VCF records -> rvt_vcf_pack.h (2-bit rows + AF) -> rvt_gene_push_bed -> device, against the
    oracle run on the matrix the REFERENCE's parser semantics give (decode + imputeGenotypeToMean).  One gene with and one
    without missing calls."""
    from test_vcf_pack import _header
    O = oracle
    N, C = 1501, 3
    X, y = O.synth_covariates(91, N, C)
    nm = O.fit_null_linear(X, y)
    rng = np.random.default_rng(91)
    code = np.array(["0/0", "0/1", "1/1", "./.", "1|0", "0/2"])
    eng = engine_cls(0)
    try:
        eng.set_null_model(X, y)
        assert vcfpack.header(_header(N)) == N
        vcfpack.set_range("7:100-200")
        refs = []
        for gene, with_missing in enumerate((False, True)):
            vcfpack.clear()
            M = 12
            maf = np.linspace(0.01, 0.2, M)
            for j in range(M + 3):
                g = rng.binomial(2, maf[j % M], size=N)
                if with_missing:
                    g = np.where(rng.random(N) < 0.01, 3, g)
                    g = np.where(rng.random(N) < 0.003, 5, g)      # multi-allelic call -> missing
                g = np.where((g == 1) & (rng.random(N) < 0.5), 4, g)  # phased het, same value
                pos = 100 + 5 * j if j < M else 300 + j               # the last three lie outside the range
                rec = "\t".join(["7", str(pos), ".", "A", "G", "50", "PASS", ".", "GT:GQ"] + [c + ":9" for c in code[g]])
                assert vcfpack.add(rec) == (1 if j < M else 0)
            rows, af, counts, names = vcfpack.gene()
            assert rows.shape == (M, (N + 3) // 4) and (counts[:, 3].sum() > 0) == with_missing
            raw = O.bed_decode_fast(rows, N).T
            refs.append((O.impute_mean(raw), af.copy()))
            eng.push_bed(rows, af)
        res = eng.flush()
    finally:
        eng.close()
    for k, (Gd, af) in enumerate(refs):
        ref, lam = O.gene(Gd, af, X, nm["resid"], nm["sigma2"])
        check_gene(res[k], ref, lam, ctx=f"vcf gene {k}")


@pytest.mark.parametrize("case", [(175, 700, 8, 1, 0.02), (171, 3001, 30, 3, 0.01), (172, 1200, 1, 2, 0.03)])
def test_skato_on_genes_with_missing_calls(engine_cls, oracle, case):
    from oracle import skato_oracle as SO
    from rvtests_b200.synth import pack_bed
    O = oracle
    seed, N, M, C, miss = case
    G, X, y = make_problem(O, seed, N, M, C, maf=np.linspace(0.004, 0.3 if M <= 8 else 0.03, M), n_flip=2 if M > 2 else 0)
    rng = np.random.default_rng(seed)
    mask = rng.random((M, N)) < miss
    bed = pack_bed(G.T, mask)
    raw = O.bed_decode_fast(bed, N).T
    Gd = O.impute_mean(raw)
    af = 0.5 * np.where(raw >= 0, raw, 0.0).sum(axis=0) / N
    nm = O.fit_null_linear(X, y)
    eng = engine_cls(0)
    try:
        eng.set_option("skato", 1)
        eng.set_null_model(X, y)
        eng.push_bed(bed, af)                       # missing calls -> fp64 path
        eng.push_f64(Gd, af)                        # the same gene as dosages -> fp64 path
        eng.push_bed(pack_bed(G.T), af_of(G))       # complete gene -> integer sweep, same flush
        r = eng.flush()
    finally:
        eng.close()
    ref = SO.skato_gene(Gd, af, X, nm["resid"])
    _check_skato(r[0], ref, f"bed+missing {case}")
    _check_skato(r[1], ref, f"dosage {case}")
    _check_skato(r[2], SO.skato_gene(G.astype(float), af_of(G), X, nm["resid"]), f"complete {case}")
    if M > 1:                                       # the SKAT / burden columns are unaffected by enabling SKAT-O
        refs, lam = O.gene(Gd, af, X, nm["resid"], nm["sigma2"])
        check_gene(r[0], refs, lam, ctx=f"skat columns with skato on, bed+missing {case}")
        check_gene(r[1], refs, lam, ctx=f"skat columns with skato on, dosage {case}")


@pytest.mark.parametrize("case", [(180, 900, 12, 1, 0.3), (181, 4000, 40, 3, 0.05), (182, 2500, 64, 2, 0.1), (183, 1500, 1, 2, 0.2),
                                  (184, 1800, 2, 3, 0.2)])
def test_binary_trait_skato_vs_oracle(engine_cls, oracle, case):
    from oracle import binary_oracle as BIN
    from oracle import skato_oracle as SO
    from rvtests_b200.synth import pack_bed
    O = oracle
    seed, N, M, C, hi = case
    G, X, _ = make_problem(O, seed, N, M, C, maf=np.linspace(0.004, hi, M), n_flip=2 if M > 2 else 0, n_mono=1 if M > 12 else 0)
    rng = np.random.default_rng(seed)
    eta = -0.8 + (X[:, 1:] @ np.full(C - 1, 0.5) if C > 1 else 0.0)
    y = (rng.random(N) < 1.0 / (1.0 + np.exp(-eta))).astype(np.float64)
    nm = BIN.fit_null_logistic(X, y)
    af = af_of(G)
    ref_skat = BIN.gene(G.astype(float), af, X, nm)
    ref = SO.skato_gene(G.astype(float), af, X, nm["resid"], vv=nm["v"])
    eng = engine_cls(0)
    try:
        eng.set_option("skato", 1)
        eng.set_option("skato_binary", 0)
        eng.set_null_model(X, y, binary=True)
        eng.push_i8(G.T.copy(), af)
        off = eng.flush()[0]
        assert int(off["skato_ok"]) == 0            # switched off: NA
        eng.set_option("skato_binary", 1)           # (the default)
        eng.push_i8(G.T.copy(), af)
        eng.push_bed(pack_bed(G.T), af)
        res = eng.flush()
    finally:
        eng.close()
    for k in range(2):
        _check_skato(res[k], ref, f"binary {case} push {k}")
        assert rel(res[k]["Q"], ref_skat["Q"]) <= 1e-6 and rel(res[k]["p_skat"], ref_skat["p_skat"]) <= 1e-4
        assert rel(off["Q"], res[k]["Q"]) <= 1e-12


def test_adapter_prints_skato_for_a_binary_trait(oracle, tmp_path):
    """SkatOTest adapter with setBinaryOutcome() (SKAT-O type "D" is on by default): the Q / rho / Pvalue columns against
    the oracle; with enableSkatOBinary(false) the line is NA (tests/test_gpu_adapters.py)."""
    import struct
    import subprocess
    from oracle import binary_oracle as BIN
    from oracle import skato_oracle as SO
    from test_gpu_adapters import build_demo
    import rvtests_b200
    rvtests_b200.load_library()
    O = oracle
    N, C = 1500, 2
    genes = []
    X = None
    for gi, (M, nm_, nf) in enumerate([(8, 0, 1), (30, 2, 2), (1, 0, 0)]):
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
    out = subprocess.run([exe, str(path), "8", "0", "0.05", "1"], capture_output=True, text=True, check=True).stdout
    tables, cur = {}, None
    for line in out.splitlines():
        if line.startswith("#"):
            cur = line[1:]
            tables[cur] = []
        else:
            tables[cur].append(line.split("\t"))
    nm = BIN.fit_null_logistic(X, y)
    for gi, G in enumerate(genes):
        ref = SO.skato_gene(G.astype(float), af_of(G), X, nm["resid"], vv=nm["v"])
        ro = tables["SkatO"][1 + gi][3:]
        assert ref["ok"] and "NA" not in ro, (gi, ro)
        # "%g" prints 6 significant digits
        assert rel(float(ro[0]), ref["Q"]) <= 2e-5 and float(ro[1]) == ref["rho"] and rel(float(ro[2]), ref["pvalue"]) <= 3e-5, (gi, ro, ref)


@pytest.mark.parametrize("case", [(190, 5000, 24, 3), (191, 777, 1, 1), (192, 20011, 64, 2)])
def test_zero_copy_entry_points(engine_cls, oracle, case):
    """caller-owned device memory end to end: rvt_set_null_model_dev (X, y in HBM) + rvt_gene_push_dev_i8 ([M][ld] int8 block in
    HBM, engine counts the rows itself) + rvt_flush_dev (records left in HBM) against the oracle and against the host entry
    points on the same context"""
    import torch
    import rvtests_b200
    O = oracle
    seed, N, M, C = case
    G, X, y = make_problem(O, seed, N, M, C, n_flip=min(2, M - 1), n_mono=1 if M > 3 else 0)
    af = af_of(G)
    nm = O.fit_null_linear(X, y)
    ld = (N + 15) // 16 * 16 + 32
    block = np.zeros((M, ld), dtype=np.int8)
    block[:, :N] = G.T
    dev = torch.device("cuda", 0)
    dG = torch.from_numpy(block).to(dev)
    dX = torch.from_numpy(np.asfortranarray(X).T.copy()).to(dev)      # column-major N x C = C contiguous columns
    dy = torch.from_numpy(np.ascontiguousarray(y)).to(dev)
    rec = rvtests_b200.engine.RESULT_DTYPE
    d_out = torch.zeros(2 * rec.itemsize, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    eng = engine_cls(0)
    try:
        eng.set_null_model_dev(N, C, dX.data_ptr(), dy.data_ptr())
        got = eng.get_null_model()
        assert np.max(np.abs(got["resid"] - nm["resid"])) <= 1e-9 * max(1.0, np.max(np.abs(nm["resid"])))
        assert rel(got["sigma2"], nm["sigma2"]) <= 1e-12
        eng.push_dev_i8(dG.data_ptr(), M, ld, af)
        eng.push_i8(G.T.copy(), af)
        n = eng.flush_dev(d_out.data_ptr(), 2)
        assert n == 2
        torch.cuda.synchronize()
        res = np.frombuffer(d_out.cpu().numpy().tobytes(), dtype=rec)
    finally:
        eng.close()
    ref, lam = O.gene(G.astype(float), af, X, nm["resid"], nm["sigma2"])
    check_gene(res[0], ref, lam, ctx=f"device block {case}")
    check_gene(res[1], ref, lam, ctx=f"host block, records on the device {case}")
    for k in ("Q", "cmc_nonref", "cmc_U", "zeg_U", "m_poly"):
        assert res[0][k] == res[1][k], k       # same exact integer sums whichever way the block arrived


def test_ingest_demo_vcf_and_bed(oracle, tmp_path):
    """rvtests_b200/host/ingest_demo.cpp: setFile + plain-text VCF (rvt_vcf_pack.h) and setFile + PLINK fileset
    (rvt_bed_file.h) through the C ABI in C++, one flush each, against the oracle (intercept-only null model)"""
    import os
    import subprocess
    from rvtests_b200.synth import pack_bed
    from test_vcf_pack import _header
    O = oracle
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "rvtests_b200", "host", "ingest_demo")
    subprocess.run(["g++", "-std=c++11", "-O2", "-I", os.path.join(root, "include"), "-I", os.path.join(root, "rvtests_b200", "host"),
                    os.path.join(root, "rvtests_b200", "host", "ingest_demo.cpp"), "-o", exe, "-L", os.path.join(root, "rvtests_b200"),
                    "-lrvtests_b200", "-lz", "-ldl", "-Wl,-rpath," + os.path.join(root, "rvtests_b200")], check=True)
    N, M = 1203, 30
    rng = np.random.default_rng(404)
    maf = np.linspace(0.01, 0.2, M)
    G = rng.binomial(2, maf[None, :], size=(N, M)).astype(np.int8)
    miss = rng.random((N, M)) < 0.004
    miss[:, :20] = False                                   # set A (first 10 variants) and B (next 10) complete, C with missing calls
    y = rng.normal(size=N) + 0.3 * G[:, 3]
    pos = 100 + 10 * np.arange(M)
    sets = {"A": "1:100-190", "B": "1:200-290", "C": "1:300-390", "EMPTY": "2:1-5"}
    sf = tmp_path / "sets.txt"
    sf.write_text("".join(f"{k} {v}\n" for k, v in sets.items()))
    # --- VCF
    code = np.array(["0/0", "0/1", "1/1"])
    with open(tmp_path / "g.vcf", "w") as f:
        f.write("##fileformat=VCFv4.1\n" + _header(N) + "\n")
        for j in range(M):
            gt = np.where(miss[:, j], "./.", code[G[:, j]])
            f.write("\t".join(["1", str(pos[j]), ".", "A", "G", "9", "PASS", ".", "GT"] + list(gt)) + "\n")
    with open(tmp_path / "ph.txt", "w") as f:
        for i in range(N):
            f.write(f"P{i + 1} {float(y[i])!r}\n")
    # --- PLINK fileset (same calls)
    prefix = str(tmp_path / "fs")
    with open(prefix + ".bed", "wb") as f:
        f.write(bytes([0x6C, 0x1B, 0x01]))
        f.write(pack_bed(G.T, miss.T).tobytes())
    with open(prefix + ".bim", "w") as f:
        for j in range(M):
            f.write(f"1\trs{j}\t0\t{pos[j]}\tA\tG\n")
    with open(prefix + ".fam", "w") as f:
        for i in range(N):
            f.write(f"F{i} P{i + 1} 0 0 1 {float(y[i])!r}\n")
    X = np.ones((N, 1))
    nm = O.fit_null_linear(X, y)
    raw = np.where(miss, -9.0, G.astype(float))
    want = {}
    for k, (a, b) in {"A": (0, 10), "B": (10, 20), "C": (20, 30)}.items():
        Gd = O.impute_mean(raw[:, a:b])
        af = 0.5 * np.where(raw[:, a:b] >= 0, raw[:, a:b], 0.0).sum(axis=0) / N
        want[k] = O.gene(Gd, af, X, nm["resid"], nm["sigma2"])[0]
    # --- the same VCF bgzipped + tabix-indexed (rvt_bgzf.h writes it, the demo reads it back by region queries)
    from test_bgzf_tabix import load_bgzf_check
    bzl = load_bgzf_check()
    text = open(tmp_path / "g.vcf", "rb").read()
    gz = str(tmp_path / "g.vcf.gz")
    assert bzl.bz_write_indexed(gz.encode(), text, len(text), 1 << 16) == 0
    for argv in (["vcf", str(tmp_path / "g.vcf"), str(sf), str(tmp_path / "ph.txt")], ["bed", prefix, str(sf)],
                 ["vcfgz", gz, str(sf), str(tmp_path / "ph.txt")]):
        out = subprocess.run([exe] + argv, capture_output=True, text=True, check=True).stdout.splitlines()
        assert out[0].split("\t") == ["Set", "NumPolyVar", "Q", "Pvalue", "NonRefSite", "CMC_P", "Zeggini_P"]
        rows = [l.split("\t") for l in out[1:]]
        assert [r[0] for r in rows] == ["A", "B", "C"], argv[0]                 # the empty set is not pushed
        for r in rows:
            ref = want[r[0]]
            assert int(r[4]) == ref.cmc_nonref
            # "%g": 6 significant digits
            assert rel(float(r[2]), ref.skat.Q) <= 2e-5 and rel(float(r[3]), ref.skat.pvalue) <= 2e-4, (argv[0], r, ref.skat.Q, ref.skat.pvalue)


def _write_bgen(path, ids, chrom, pos, P0, P1, miss, bits=16):
    """BGEN v1.2 (layout 2, zlib, unphased diploid bi-allelic): P0 / P1 = integer numerators of p(hom ref) / p(het), (M, N)"""
    import struct
    import zlib
    N, M = len(ids), len(pos)
    sblock = struct.pack("<II", 8 + sum(2 + len(s) for s in ids), N) + b"".join(struct.pack("<H", len(s)) + s.encode() for s in ids)
    header = struct.pack("<III4sI", 20, M, N, b"bgen", 1 | (2 << 2) | (1 << 31))
    out = struct.pack("<I", len(header) + len(sblock)) + header + sblock
    dt = {8: "<u1", 16: "<u2", 32: "<u4"}[bits]
    for j in range(M):
        def s2(x):
            return struct.pack("<H", len(x)) + x.encode()
        ident = s2(f"v{j}") + s2(f"rs{j}") + s2(chrom) + struct.pack("<IH", int(pos[j]), 2) + struct.pack("<I", 1) + b"A" + struct.pack("<I", 1) + b"G"
        pm = (2 | (miss[j].astype(np.uint8) << 7)).astype(np.uint8).tobytes()
        probs = np.stack([P0[j], P1[j]], axis=1).astype(dt).tobytes()
        payload = struct.pack("<IHBB", N, 2, 2, 2) + pm + struct.pack("<BB", 0, bits) + probs
        comp = zlib.compress(payload)
        out += ident + struct.pack("<II", len(comp) + 4, len(payload)) + comp
    open(path, "wb").write(out)


def test_ingest_demo_bgen_dosages(oracle, tmp_path):
    """BGEN (rvt_bgen.h) -> dosages -> mean imputation -> rvt_gene_push_f64 in C++ (ingest_demo bgen), against the oracle on
    the dosages the reference's float arithmetic gives (BitReader: numerator * float(1 / (2^B - 1)); remainder in float;
    p(het) + 2 p(hom alt) in double)."""
    import os
    import subprocess
    O = oracle
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "rvtests_b200", "host", "ingest_demo")
    subprocess.run(["g++", "-std=c++11", "-O2", "-I", os.path.join(root, "include"), "-I", os.path.join(root, "rvtests_b200", "host"),
                    os.path.join(root, "rvtests_b200", "host", "ingest_demo.cpp"), "-o", exe, "-L", os.path.join(root, "rvtests_b200"),
                    "-lrvtests_b200", "-lz", "-ldl", "-Wl,-rpath," + os.path.join(root, "rvtests_b200")], check=True)
    N, M, bits = 907, 24, 16
    rng = np.random.default_rng(808)
    top = (1 << bits) - 1
    maf = np.linspace(0.02, 0.3, M)
    g = rng.binomial(2, maf[:, None], size=(M, N))
    # imputed-looking probabilities around the call
    noise = rng.dirichlet([0.3, 0.3, 0.3], size=(M, N))
    pr = 0.85 * np.eye(3)[g] + 0.15 * noise
    P0 = np.floor(pr[..., 0] * top).astype(np.int64)
    P1 = np.minimum(np.floor(pr[..., 1] * top).astype(np.int64), top - P0)
    miss = rng.random((M, N)) < 0.01
    miss[:8] = False
    ids = [f"S{i}" for i in range(N)]
    pos = 1000 + 7 * np.arange(M)
    path = str(tmp_path / "d.bgen")
    _write_bgen(path, ids, "7", pos, P0, P1, miss, bits)
    scale = np.float32(1.0 / np.float32(top))                # BitReader: float scale = 1 / (2^B - 1)
    p0 = P0.astype(np.float32) * scale
    p1 = P1.astype(np.float32) * scale
    p2 = (np.float32(1.0) - p0) - p1
    D = p1.astype(np.float64) + p2.astype(np.float64) * 2.0
    D[miss] = -9.0
    y = rng.normal(size=N) + 0.4 * D[2].clip(0)
    with open(tmp_path / "ph.txt", "w") as f:
        for i in range(N):
            f.write(f"{ids[i]} {float(y[i])!r}\n")
    sf = tmp_path / "sets.txt"
    sf.write_text(f"A 7:{pos[0]}-{pos[7]}\nB 7:{pos[8]}-{pos[15]},7:{pos[20]}-{pos[23]}\nNONE 8:1-100\n")
    out = subprocess.run([exe, "bgen", path, str(sf), str(tmp_path / "ph.txt")], capture_output=True, text=True, check=True).stdout.splitlines()
    rows = [l.split("\t") for l in out[1:]]
    assert [r[0] for r in rows] == ["A", "B"]
    X = np.ones((N, 1))
    nm = O.fit_null_linear(X, y)
    for r, cols in zip(rows, (list(range(0, 8)), list(range(8, 16)) + list(range(20, 24)))):
        raw = D[cols].T                                        # (N, M)
        Gd = O.impute_mean_literal(raw)
        af = 0.5 * np.where(raw >= 0, raw, 0.0).sum(axis=0) / N
        ref = O.gene(Gd, af, X, nm["resid"], nm["sigma2"])[0]
        assert int(r[1]) == ref.m_poly
        assert rel(float(r[2]), ref.skat.Q) <= 2e-5 and rel(float(r[3]), ref.skat.pvalue) <= 2e-4, (r, ref.skat.Q, ref.skat.pvalue)
        assert rel(float(r[5]), ref.cmc_p) <= 2e-4 and rel(float(r[6]), ref.zeg_p) <= 2e-4
