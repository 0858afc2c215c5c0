"""CPU: genotype ingestion -- the product's host-side VCF packer (rvtests_b200/host/rvt_vcf_pack.h: VCF text
-> PLINK 2-bit rows + AF for rvt_gene_push_bed) against
  * the REFERENCE's own VCF record parser compiled unmodified (oracle/_ref/libvcf_ref.so: libVcf/VCFRecord, VCFIndividual,
    VCFValue::getGenotype ...), live when oracle/_ref is built,
  * the committed golden vectors that build produced (tests/golden/vcf_golden.json, make_golden_vcf.py),
  * the Python restatement (oracle.vcf_gt / vcf_record_genotypes), itself held against both."""
import itertools
import json
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden", "vcf_golden.json")
ALPHABET = "012.|/-9a"


def _all_gt_strings(maxlen=4):
    for n in range(maxlen + 1):
        for t in itertools.product(ALPHABET, repeat=n):
            yield "".join(t)


def test_gt_grammar_exhaustive(vcfpack, oracle, capfd):
    """every string of up to 4 characters over '012.|/-9a': product == restatement == reference build"""
    ref = oracle.ref_vcf()
    n = 0
    for s in _all_gt_strings():
        want = oracle.vcf_gt(s)
        assert vcfpack.gt(s) == want, s
        if ref is not None:
            assert ref.ref_vcf_gt(s.encode(), len(s)) == want, s
        n += 1
    capfd.readouterr()   # the reference REPORTs malformed calls on stderr
    assert n == sum(len(ALPHABET) ** k for k in range(5))
    # the quirks, spelled out (libVcf/VCFValue.h:74-116)
    assert [oracle.vcf_gt(s) for s in ("0/1", "1|1", "0", "1", "./.", "0/2", "2/1", "0/", "1/.", "0/-", "0/1/1", "")] == \
        [1, 2, 0, 1, -9, -9, -9, -9, -9, 0, -9, -9]


def _random_record(rng, n, pos):
    fmt_pool = [["GT"], ["GT", "GD", "GQ"], ["GD", "GT"], ["DS", "GQ", "GT"], ["GTX", "GT"], ["GD", "GQ"], ["PGT", "GT", "GL"]]
    fmt = fmt_pool[int(rng.integers(len(fmt_pool)))]
    gts = ["0/0", "0/1", "1/0", "1/1", "0|1", "1|1", "./.", ".", "0", "1", "0/2", "2/1", "1/.", "./1", "0/1/1", "0/", "00", "1-1"]
    w = np.array([40, 10, 6, 4, 4, 2, 4, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1], dtype=float)
    cols = []
    for _ in range(n):
        sub = []
        for k in fmt:
            if k in ("GT", "GTX", "PGT"):
                sub.append(gts[int(rng.choice(len(gts), p=w / w.sum()))])
            else:
                sub.append(str(int(rng.integers(0, 99))))
        keep = len(sub) if rng.random() > 0.08 else int(rng.integers(1, len(sub) + 1))   # truncated sample column
        cols.append(":".join(sub[:keep]))
    chrom = "1" if rng.random() < 0.7 else "X"
    return "\t".join([chrom, str(pos), f"r{pos}", "A", "G", "100", "PASS", "DP=10;AF=0.1", ":".join(fmt)] + cols)


def _header(n):
    return "\t".join(["#CHROM", "POS", "ID", "REF", "ALT", "QUAL", "FILTER", "INFO", "FORMAT"] + [f"P{i + 1}" for i in range(n)])


def _decode(oracle, rows, n):
    return oracle.bed_decode_fast(rows, n) if len(rows) else np.zeros((0, n))


@pytest.mark.parametrize("n", [1, 3, 4, 5, 9, 64, 131])
def test_records_vs_reference_parser(vcfpack, oracle, n, capfd):
    rng = np.random.default_rng(1000 + n)
    hdr = _header(n)
    assert vcfpack.header(hdr) == n
    vcfpack.set_range("")
    vcfpack.clear()
    want, afs, sites = [], [], []
    for k in range(60):
        rec = _random_record(rng, n, 10 * (k + 1))
        exp = oracle.vcf_record_genotypes(hdr, rec)
        if oracle.ref_vcf() is not None:
            r = oracle.ref_vcf_genotypes(hdr, rec)
            assert r is not None and r[0] == exp[0] and r[1] == exp[1]
            assert np.array_equal(r[2], exp[2]), rec
        assert vcfpack.add(rec + ("\n" if k % 3 == 0 else "")) == 1
        want.append(exp[2])
        # GenotypeCounter::add / getAF (src/GenotypeCounter.h:14-52): missing calls stay in the denominator
        afs.append(0.5 * exp[2][exp[2] >= 0].sum() / n)
        sites.append(f"{exp[0]}:{exp[1]}")
    capfd.readouterr()
    rows, af, counts, names = vcfpack.gene()
    want = np.array(want)
    assert rows.shape == (60, (n + 3) // 4)
    got = _decode(oracle, rows, n)
    assert np.array_equal(got, want.astype(float))
    assert np.array_equal(af, np.array(afs))
    assert names == sites
    for k, key in enumerate((0, 1, 2, -9)):
        assert np.array_equal(counts[:, k], (want == key).sum(axis=1))
    # padding bits of the last byte are zero (the engine's staging kernel reads whole bytes)
    if n % 4:
        assert np.all(rows[:, -1] >> (2 * (n % 4)) == 0)


def test_golden_vectors(vcfpack, oracle):
    """tests/golden/vcf_golden.json: records + what the reference build returned for them (runs without oracle/_ref)"""
    g = json.load(open(GOLD))
    hdr = g["header"]
    n = len(hdr.split("\t")) - 9
    assert vcfpack.header(hdr) == n
    vcfpack.set_range("")
    vcfpack.clear()
    for rec, exp in zip(g["records"], g["genotypes"]):
        assert vcfpack.add(rec) == 1
        assert list(oracle.vcf_record_genotypes(hdr, rec)[2]) == exp
    rows, af, counts, names = vcfpack.gene()
    assert np.array_equal(_decode(oracle, rows, n), np.array(g["genotypes"], dtype=float))
    assert names == g["sites"]
    for s, exp in g["gt_strings"].items():
        assert vcfpack.gt(s) == exp and oracle.vcf_gt(s) == exp, s


def test_sample_subset_ranges_and_malformed_lines(vcfpack, oracle):
    n = 7
    hdr = _header(n)
    rng = np.random.default_rng(5)
    recs = [_random_record(rng, n, p) for p in (5, 10, 20, 30, 31, 400)]
    # keep-list: a SET of names, output in VCF column order (VCFRecord::includePeople)
    assert vcfpack.header(hdr, keep=["P6", "P2", "P3"]) == 3
    assert vcfpack.sample_names() == ["P2", "P3", "P6"]
    assert vcfpack.header(hdr, keep=["P6", "nobody"]) < 0
    assert vcfpack.header(hdr, keep=["P6", "P2", "P3"]) == 3
    chrom = [r.split("\t")[0] for r in recs]
    vcfpack.set_range("1:10-30,X:10-30")
    vcfpack.clear()
    taken = [vcfpack.add(r) for r in recs]
    assert taken == [0, 1, 1, 1, 0, 0]
    rows, af, counts, names = vcfpack.gene()
    full = np.array([oracle.vcf_record_genotypes(hdr, r)[2] for r in recs[1:4]])
    assert np.array_equal(_decode(oracle, rows, 3), full[:, [1, 2, 5]].astype(float))
    assert names == [f"{c}:{p}" for c, p in zip(chrom[1:4], (10, 20, 30))]
    assert np.array_equal(af, np.array([0.5 * r[r >= 0].sum() / 3 for r in full[:, [1, 2, 5]]]))
    # other range spellings (parseRangeFormat): open-ended forms run to 1 << 29, pieces that do not conform are skipped
    assert vcfpack.set_range("1:5") == 1 and vcfpack.set_range("1:5-") == 1 and vcfpack.set_range("1:5-9,2:1-2") == 2
    assert vcfpack.set_range("1") == 0 and vcfpack.set_range("1:30-10") == 0 and vcfpack.set_range("1:x-3,1:4-5") == 1
    vcfpack.set_range("")
    # comment / meta lines are skipped; a record with a different sample count is an error and leaves no row behind
    vcfpack.clear()
    assert vcfpack.add("##fileformat=VCFv4.0") == 0 and vcfpack.add(hdr) == 0 and vcfpack.add("") == 0
    short = "\t".join(recs[0].split("\t")[:-1])
    assert vcfpack.add(short) < 0 and oracle.vcf_record_genotypes(hdr, short) is None
    assert vcfpack.add(recs[0] + "\t0/1") < 0
    assert vcfpack.add("1\t10\tr\tA\tG") < 0
    assert vcfpack.gene()[0].shape[0] == 0
    assert vcfpack.add(recs[0]) == 1 and vcfpack.gene()[0].shape[0] == 1
    # (the reference's parseIndividual asserts / returns -1 on such a record, libVcf/VCFRecord.h:129-201: not driven here)


def _random_dosage_record(rng, n, pos):
    fmt_pool = [["GT", "DS"], ["DS"], ["GT", "GQ", "DS"], ["DSX", "DS"], ["GT", "GQ"]]
    fmt = fmt_pool[int(rng.integers(len(fmt_pool)))]
    # (no EMPTY value: a sample column that ends in ':' makes VCFIndividual::parse call parseTill past the end, which
    # returns -1 without touching the VCFValue, and the loop then writes a NUL at that value's STALE end offset from the
    # previous record -- libVcf/VCFIndividual.h:39-53, VCFValue.h:262-263; the reference's result is then state-dependent.
    # The packer reads such a subfield as the empty string.)
    vals = ["0", "0.013", "1", "1.5", "2", "1.999", "0.666667", "1.333334", ".", "2.5", "-1", "1e-1", "0.5x", "nan"]
    cols = []
    for _ in range(n):
        sub = []
        for k in fmt:
            if k in ("DS", "DSX"):
                sub.append(vals[int(rng.integers(len(vals)))] if rng.random() < 0.4 else "%.3f" % rng.uniform(0, 2))
            elif k == "GT":
                sub.append("0/1")
            else:
                sub.append("30")
        keep = len(sub) if rng.random() > 0.1 else int(rng.integers(1, len(sub) + 1))
        cols.append(":".join(sub[:keep]))
    return "\t".join(["2", str(pos), ".", "C", "T", "9", "PASS", ".", ":".join(fmt)] + cols)


@pytest.mark.parametrize("n", [1, 6, 50])
def test_dosage_mode_vs_reference_parser(vcfpack, oracle, n):
    """--dosage DS: toDouble() of the tagged subfield (atof: '.', '' and a truncated column are 0.0, not missing; a record
    without the key is all-missing), GenotypeCounter thresholds and AF, and the literal mean imputation of negative entries"""
    rng = np.random.default_rng(2000 + n)
    hdr = _header(n)
    vcfpack.set_dosage_tag("DS")
    try:
        assert vcfpack.header(hdr) == n
        vcfpack.set_range("")
        vcfpack.clear()
        want = []
        for k in range(40):
            rec = _random_dosage_record(rng, n, 3 * k + 1)
            exp = oracle.vcf_record_dosages(hdr, rec, "DS")
            if oracle.ref_vcf() is not None:
                r = oracle.ref_vcf_dosages(hdr, rec, "DS")
                assert np.array_equal(r, exp, equal_nan=True), rec
            assert vcfpack.add(rec) == 1
            want.append(exp)
        want = np.array(want).T                                   # (N, M)
        G, af, counts = vcfpack.dosage_gene(raw=True)
        assert np.array_equal(G, want, equal_nan=True)
        for j in range(want.shape[1]):
            col = want[:, j]
            if np.isnan(col).any():
                continue                                          # ("nan" compares false everywhere: falls in the last branch)
            r0, r1, r2, rm, a = oracle.genotype_counter(col)
            assert list(counts[j]) == [r0, r1, r2, rm], j
            assert af[j] == a, j
        Gi, _, _ = vcfpack.dosage_gene(raw=False)
        ok = ~np.isnan(want).any(axis=0)
        assert np.array_equal(Gi[:, ok], oracle.impute_mean_literal(want[:, ok]))
        # integer hard-call matrices: the literal rule equals the restatement the model-layer pin uses
        H = rng.integers(-1, 3, size=(30, 5)).astype(float)
        H[H < 0] = -9
        assert np.array_equal(oracle.impute_mean_literal(H), oracle.impute_mean(H))
    finally:
        vcfpack.set_dosage_tag("")


@pytest.mark.parametrize("flt", [((3, 0), (-1, -1)), ((-1, -1), (3, 0)), ((2, 40), (10, 90)), ((0, 50), (-1, -1)), ((0, 0), (0, 0))])
def test_depth_and_quality_filters(vcfpack, oracle, flt, capfd):
    """--indvDepthMin/Max, --indvQualMin/Max (the reference's own test/Makefile check3 / check4 use the two minima): calls
    failing the GD / GQ bounds become missing; a record without the key reads 0 for everyone"""
    if oracle.ref_vcf() is None:
        pytest.skip("oracle/_ref/libvcf_ref.so not built (no /root/reference here)")
    gd, gq = flt
    n = 9
    hdr = _header(n)
    rng = np.random.default_rng(77)
    vcfpack.set_filters(gd, gq)
    try:
        assert vcfpack.header(hdr) == n
        vcfpack.set_range("")
        vcfpack.clear()
        want, plain = [], []
        for k in range(50):
            rec = _random_record(rng, n, k + 1)
            want.append(oracle.ref_vcf_genotypes_filtered(hdr, rec, gd, gq))
            plain.append(oracle.ref_vcf_genotypes(hdr, rec)[2])
            assert vcfpack.add(rec) == 1
        capfd.readouterr()
        rows, af, counts, _ = vcfpack.gene()
        want = np.array(want)
        assert np.array_equal(_decode(oracle, rows, n), want.astype(float))
        assert np.array_equal(counts[:, 3], (want == -9).sum(axis=1))
        if gd[0] > 0 or gq[0] > 0:
            assert (want == -9).sum() > (np.array(plain) == -9).sum()      # the filter did remove calls
        if flt == ((0, 0), (0, 0)):
            assert np.array_equal(want, np.array(plain))                    # switched on without bounds: nothing changes
    finally:
        vcfpack.set_filters()


def test_male_hemizygous_grammar_and_par_regions(vcfpack, oracle, capfd):
    """VCFValue::getMaleNonParGenotype02 on every short string, ParRegion::isHemiRegion for the built-in builds and custom
    --xLabel / --xParRegion strings: product == reference build"""
    if oracle.ref_vcf() is None:
        pytest.skip("oracle/_ref/libvcf_ref.so not built (no /root/reference here)")
    ref = oracle.ref_vcf()
    for s in _all_gt_strings():
        assert vcfpack.gt_male02(s) == ref.ref_vcf_gt_male02(s.encode(), len(s)), s
    capfd.readouterr()
    assert [vcfpack.gt_male02(s) for s in ("0", "1", "0/0", "1|1", "0/1", "2", "./.", "1", "1/", "")] == [0, 2, 0, 2, -9, -9, -9, 2, -9, -9]
    probes = [1, 10000, 10001, 60000, 60001, 2699520, 2699521, 2709520, 2781479, 2781480, 154584238, 154931043, 154931044,
              155260560, 155260561, 155701383, 156030895, 156030896, 2_000_000_000]
    for xl in ("", "X", "chrX,X,chr23,23", "Z"):
        for pr in ("", "hg19", "HG38", "b36", "grch37", "100-200", "100-200,5000-", "-100,300-400", "7-9-11,50-60", "junk"):
            for chrom in ("X", "23", "chrX", "1", "Z"):
                for pos in probes + [99, 100, 200, 201, 300, 400, 401, 5000, 55]:
                    assert vcfpack.par_is_hemi(xl, pr, chrom, pos) == ref.ref_par_is_hemi(xl.encode(), pr.encode(), chrom.encode(), pos), \
                        (xl, pr, chrom, pos)


@pytest.mark.parametrize("dosage", [False, True])
def test_records_on_x_with_sex(vcfpack, oracle, dosage, capfd):
    """males 0 / 2 (dosage x 2) outside the pseudo-autosomal regions, females as usual, unknown sex missing; autosomes and
    the PAR untouched"""
    if oracle.ref_vcf() is None:
        pytest.skip("oracle/_ref/libvcf_ref.so not built (no /root/reference here)")
    n = 10
    hdr = _header(n)
    rng = np.random.default_rng(9)
    sex = np.array([1, 2, 1, 2, 0, -9, 1, 2, 3, 1])
    tag = "DS" if dosage else ""
    vcfpack.set_dosage_tag(tag)
    try:
        assert vcfpack.header(hdr) == n
        assert vcfpack.set_sex(sex[:3]) != 0 and vcfpack.set_sex(sex) == 0
        vcfpack.set_range("")
        vcfpack.clear()
        want = []
        for k, pos in enumerate([5, 60001, 100000, 2699520, 2699521, 3000000, 154931043, 154931044, 155260561] * 4):
            rec = (_random_dosage_record if dosage else _random_record)(rng, n, pos)
            f = rec.split("\t")
            f[0] = ["X", "23", "1", "chrX"][k % 4]
            rec = "\t".join(f)
            want.append(oracle.ref_vcf_genotypes_sex(hdr, rec, sex, dosage_tag=tag))
            assert vcfpack.add(rec) == 1
        capfd.readouterr()
        want = np.array(want)
        if dosage:
            G, _, _ = vcfpack.dosage_gene(raw=True)
            assert np.array_equal(G.T, want, equal_nan=True)
        else:
            rows, _, _, _ = vcfpack.gene()
            assert np.array_equal(_decode(oracle, rows, n), want)
            hemi_rows = [k for k in range(len(want)) if k % 4 in (0, 1) and k % 9 in (0, 4, 5, 6, 8)]
            assert np.all(want[hemi_rows][:, sex == 1] != 1)           # no heterozygous male there
            assert np.all(want[hemi_rows][:, (sex != 1) & (sex != 2)] == -9)
    finally:
        vcfpack.set_dosage_tag("")
        vcfpack.set_sex(None)


def test_multi_allelic_grammar_exhaustive(vcfpack, oracle, capfd):
    """VCFValue::countAltAllele / countMaleNonParAltAllele2 for alt = 1, 2, 3 on every non-empty string of up to 4 characters
    (the reference reads past its buffer on the empty value, and asserts on a first allele above '9' in the male form)"""
    if oracle.ref_vcf() is None:
        pytest.skip("oracle/_ref/libvcf_ref.so not built (no /root/reference here)")
    ref = oracle.ref_vcf()
    for s in _all_gt_strings():
        if not s:
            continue
        for alt in (1, 2, 3):
            assert vcfpack.count_alt(s, alt) == ref.ref_vcf_count_alt(s.encode(), len(s), alt), (s, alt)
            if s[0] not in "a|":        # (bytes above '9': the reference asserts 0 <= g <= 9)
                assert vcfpack.count_male_alt2(s, alt) == ref.ref_vcf_count_male_alt2(s.encode(), len(s), alt), (s, alt)
    capfd.readouterr()
    assert [vcfpack.count_alt(s, 2) for s in ("0/2", "2/2", "2|1", "1/1", "2", "./2", "0/", "")] == [1, 2, 1, 0, 1, -9, -9, -9]


@pytest.mark.parametrize("with_sex", [False, True])
def test_multi_allelic_records(vcfpack, oracle, with_sex, capfd):
    """--multipleAllele: a record with K ALT alleles becomes K rows (copies of allele a), named chrom:posREF/ALTa"""
    if oracle.ref_vcf() is None:
        pytest.skip("oracle/_ref/libvcf_ref.so not built (no /root/reference here)")
    n = 8
    hdr = _header(n)
    rng = np.random.default_rng(31)
    sex = np.array([1, 2, 1, 2, 0, 1, 2, 1]) if with_sex else None
    gts = ["0/0", "0/1", "1/1", "0/2", "1/2", "2/2", "2|0", "0", "1", "2", "./.", "3/1", "0/3", ".", "1/", "0/1/2"]
    vcfpack.set_multi(True)
    try:
        assert vcfpack.header(hdr) == n
        assert vcfpack.set_sex(sex) == 0
        vcfpack.set_range("")
        vcfpack.clear()
        want, names = [], []
        for k in range(40):
            alts = ["G", "G,T", "G,T,<DEL>", "G,,T"][k % 4]          # "G,,T": the empty token is an allele of its own (stringTokenize keeps it)
            chrom, pos = ["X", "5"][k % 2], [70000, 5_000_000][(k // 2) % 2]
            cols = [gts[int(rng.integers(len(gts)))] + ":7" for _ in range(n)]
            rec = "\t".join([chrom, str(pos), ".", "AC", alts, "9", "PASS", ".", "GT:GQ"] + cols)
            na = alts.count(",") + 1
            assert vcfpack.add(rec) == na
            for a in range(1, na + 1):
                g, n_alt = oracle.ref_vcf_genotypes_alt(hdr, rec, a, sex)
                assert n_alt == na
                want.append(g)
                names.append(f"{chrom}:{pos}AC/{alts.split(',')[a - 1]}")
        capfd.readouterr()
        rows, af, counts, got_names = vcfpack.gene()
        want = np.array(want)
        assert np.array_equal(_decode(oracle, rows, n), want.astype(float))
        assert got_names == names
        assert np.array_equal(af, np.array([0.5 * r[r >= 0].sum() / n for r in want]))
        # a malformed record in this mode leaves nothing behind either
        m0 = rows.shape[0]
        assert vcfpack.add("\t".join(["5", "9", ".", "A", "C,G", "9", "PASS", ".", "GT"] + ["0/1"] * (n - 1))) < 0
        assert vcfpack.gene()[0].shape[0] == m0
    finally:
        vcfpack.set_multi(False)
        vcfpack.set_sex(None)


def test_maf_cut(vcfpack, oracle, capfd):
    """--freqLower / --freqUpper (src/VCFGenotypeExtractor.cpp:98-110): rows outside [lo, hi] in MAF are undone"""
    n = 40
    hdr = _header(n)
    rng = np.random.default_rng(3)
    recs = [_random_record(rng, n, k + 1) for k in range(80)]
    gen = [oracle.vcf_record_genotypes(hdr, r)[2] for r in recs]
    maf = []
    for g in gen:
        a = 0.5 * g[g >= 0].sum() / n
        maf.append(1 - a if a > 0.5 else a)
    maf = np.array(maf)
    assert vcfpack.header(hdr) == n
    vcfpack.set_range("")
    for lo, hi in ((0.0, 0.0), (0.1, 0.0), (0.0, 0.12), (0.08, 0.15)):
        vcfpack.set_freq(lo, hi)
        try:
            vcfpack.clear()
            kept = [vcfpack.add(r) for r in recs]
        finally:
            vcfpack.set_freq()
        want = ~(((lo > 0) & (lo > maf)) | ((hi > 0) & (hi < maf)))
        assert kept == [int(w) for w in want], (lo, hi)
        rows, af, counts, names = vcfpack.gene()
        assert rows.shape[0] == want.sum() == len(af) == len(names)
        assert np.array_equal(_decode(oracle, rows, n), np.array(gen)[want].astype(float))
    capfd.readouterr()
    assert 0 < want.sum() < len(recs)


def test_range_format_vs_reference(vcfpack, oracle):
    """parseRangeFormat (base/RangeList.cpp:78-125) on every string of up to 6 characters over '1X:-0 9a' plus hand-made ones"""
    import ctypes as C
    if oracle.ref_vcf() is None:
        pytest.skip("oracle/_ref/libvcf_ref.so not built (no /root/reference here)")
    ref = oracle.ref_vcf()
    cases = ["1:100-200", "X:150", "chr7:5-", "MT", "", ":", "1:", "1:-5", "1:5-3", "1:5-5", "1: 7-9", "1:7- 9", "1:12x-20y", "1:2147483647",
             "1:2147483648", "1:99999999999", "1:5--6", "a:b-c", "1:1-2147483647", "22:0-0", "1:007-010"]
    for n in range(1, 7):
        for t in itertools.product("1X:-0 9a", repeat=n):
            cases.append("".join(t))
    chrom = C.create_string_buffer(64)
    b, e = C.c_uint(0), C.c_uint(0)
    n_ok = 0
    for s in cases:
        rc = ref.ref_parse_range(s.encode(), chrom, C.byref(b), C.byref(e))
        got = vcfpack.parse_range(s)
        if rc != 0:
            assert got is None, s
        else:
            assert got == (chrom.value.decode(), b.value, e.value), (s, got)
            n_ok += 1
    assert n_ok > 1000


def test_gene_file_and_set_file(vcfpack, tmp_path):
    """--geneFile (refFlat) and --setFile readers (src/Main.cpp:91-122, 138-173), on the reference's own test/gene.txt layout"""
    gf = tmp_path / "gene.txt"
    gf.write_text("GENE1\ttran1\t1\t-\t10\t30\t10\t30\t1\t10\t30\n"
                  "GENE2\ttran2\tchr1\t-\t40\t60\n"
                  "GENE3 tran3 X + 50 110\n"
                  "GENE1\ttran1b\t1\t-\t100\t130\n"
                  "short\tline\n"
                  "GENE9\ttran9\t2\t+\t1\t2\n")
    L = vcfpack.L
    assert L.vp_load_gene_file(str(gf).encode(), b"") == 3            # reading stops at the short line
    assert [L.vp_map_name(i).decode() for i in range(3)] == ["GENE1", "GENE2", "GENE3"]
    assert L.vp_map_nranges(0) == 2
    assert [L.vp_map_contains(0, b"1", p) for p in (9, 10, 30, 31, 100, 131)] == [0, 1, 1, 0, 1, 0]
    assert L.vp_map_contains(1, b"1", 50) == 1 and L.vp_map_contains(1, b"chr1", 50) == 0     # chopChr
    assert L.vp_map_contains(2, b"X", 110) == 1
    assert L.vp_load_gene_file(str(gf).encode(), b"GENE3,GENE2") == 2
    assert [L.vp_map_name(i).decode() for i in range(2)] == ["GENE2", "GENE3"]
    assert L.vp_load_gene_file(str(tmp_path / "nope").encode(), b"") < 0
    sf = tmp_path / "setFile"
    sf.write_text("set1 1:1-3\nset2\t1:10-20,2:5-6,junk\nlonely\n\nset1 X:7\nset3  1:1-2\n")
    assert L.vp_load_range_file(str(sf).encode(), b"") == 2           # 'lonely', the blank line and the empty column are skipped
    assert [L.vp_map_name(i).decode() for i in range(2)] == ["set1", "set2"]
    assert L.vp_map_nranges(0) == 2 and L.vp_map_nranges(1) == 2
    assert [L.vp_map_contains(0, b"1", p) for p in (0, 1, 3, 4)] == [0, 1, 1, 0]
    assert L.vp_map_contains(0, b"X", 7) == 1 and L.vp_map_contains(0, b"X", 1 << 29) == 1 and L.vp_map_contains(0, b"X", 6) == 0
    assert L.vp_map_contains(1, b"2", 6) == 1
    assert L.vp_load_range_file(str(sf).encode(), b"set2") == 1
