"""CPU: the BGEN reader (rvtests_b200/host/rvt_bgen.h) against the golden outputs of the reference's own reader
tests (libBgen/test/*.vcf.correct, written by its testBGenFile): identifiers, alleles, phasing and EVERY probability of every
sample as printed with %g -- layouts 1 (v1.1, zlib) and 2 (v1.2: zlib, zstd; 1 / 8 / 16 / 31 bits; haploid .. tetraploid,
multi-allelic, phased).  The dosage is BGenGenotypeExtractor::getGenotype's (src/BGenGenotypeExtractor.cpp:413-472)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "bgen")
REFT = "/root/reference/libBgen/test"


@pytest.fixture(scope="module")
def bg():
    d = os.path.join(ROOT, "tests", "hostcheck")
    so = os.path.join(d, "libbgencheck.so")
    src = os.path.join(d, "bgen_check.cpp")
    hdr = os.path.join(ROOT, "rvtests_b200", "host", "rvt_bgen.h")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++11", "-Wall", "-Werror", "-fPIC", "-shared", "-I", os.path.dirname(hdr), "-o", so, src,
                               "-lz", "-ldl"])
    L = C.CDLL(so)
    L.bg_dump.restype = C.c_long
    L.bg_dump.argtypes = [C.c_char_p, C.c_char_p, C.c_uint, C.c_uint, C.c_char_p, C.c_long] + [C.c_void_p] * 4
    return L


def _dump(bg, path, region=None):
    out = C.create_string_buffer(64 << 20)
    lay, comp, ns, nm = C.c_int(), C.c_int(), C.c_uint(), C.c_uint()
    chrom, beg, end = region if region else ("", 0, 0)
    n = bg.bg_dump(path.encode(), chrom.encode(), beg, end, out, len(out), C.byref(lay), C.byref(comp), C.byref(ns), C.byref(nm))
    assert n >= 0, (n, out.value)
    lines = out.value.decode().split("\n")[:-1]
    return dict(layout=lay.value, compression=comp.value, n_sample=ns.value, n_marker=nm.value, samples=lines[0].split("\t") if lines[0] else [],
                variants=[l.split("\t") for l in lines[1:]])


def _check_against_vcf(d, vcf_path):
    vcf = [l.rstrip("\n").split("\t") for l in open(vcf_path) if not l.startswith("##")]
    header, recs = vcf[0], vcf[1:]
    assert d["samples"] == header[9:]
    assert len(recs) == len(d["variants"]) == d["n_marker"]
    n_prob = 0
    for v, r in zip(d["variants"], recs):
        chrom, pos, rsid, varid, alleles, ph = v[:6]
        assert [chrom, pos] == r[:2]
        assert r[2] == (rsid + ("," + varid if varid else ""))
        assert alleles.split(",") == [r[3]] + r[4].split(",")
        assert r[8] == ("GT:HP" if ph == "P" else "GT:GP")
        for mine, ref in zip(v[6:], r[9:]):
            probs, dos = mine.split("|")
            gt, gp = ref.split(":")
            assert probs == gp, (chrom, pos, mine, ref)
            n_prob += probs.count(",") + 1
            p = gp.split(",")
            if "." in p:
                assert float(dos) == -9.0
            elif len(alleles.split(",")) == 2 and len(p) == 3:       # diploid, bi-allelic: p(het) + 2 p(hom alt)
                assert abs(float(dos) - (float(p[1]) + 2 * float(p[2]))) <= 1e-5
                want_gt = "0/0" if float(p[0]) > max(float(p[1]), float(p[2])) else ("0/1" if float(p[1]) > max(float(p[0]), float(p[2])) else "1/1")
                if abs(float(p[0]) - float(p[1])) > 1e-5 and abs(float(p[1]) - float(p[2])) > 1e-5 and abs(float(p[0]) - float(p[2])) > 1e-5:
                    assert gt == want_gt
    return n_prob


@pytest.mark.parametrize("name,correct", [("complex.bgen", "complex.bgen.vcf.correct"), ("complex.1bits.bgen", "complex.1bits.bgen.vcf.correct"),
                                          ("complex.31bits.bgen", "complex.1bits.bgen.vcf.correct")])
def test_complex_fixtures(bg, name, correct):
    d = _dump(bg, os.path.join(GOLD, name))
    assert d["layout"] == 2 and d["n_sample"] == 4 and d["n_marker"] == 10
    assert _check_against_vcf(d, os.path.join(GOLD, correct)) > 100
    # polyploid / haploid samples and multi-allelic variants: -9 unless ploidy 1 or 2 (BGenGenotypeExtractor.cpp:466-468)
    m10 = d["variants"][-1]
    assert m10[2] == "M10" and all(s.endswith("|-9") for s in m10[6:])


@pytest.mark.parametrize("name", ["example.v11.bgen", "example.16bits.zstd.bgen"])
def test_example_fixtures_from_the_reference_tree(bg, name):
    path = os.path.join(REFT, name)
    if not os.path.exists(path):
        pytest.skip("no /root/reference here")
    d = _dump(bg, path)
    assert d["n_sample"] == 500 and d["n_marker"] == 199
    assert d["layout"] == (1 if "v11" in name else 2) and d["compression"] == (1 if "v11" in name else 2)
    assert _check_against_vcf(d, path + ".vcf.correct") == 500 * 199 * 3


def test_range_filter_equals_the_reference_range_output(bg):
    """testBGenFileByRange.output.correct = the records of example.16bits.bgen with 2000 <= pos < 5000 or 6000 <= pos < 8000
    (the awk line in libBgen/test/Makefile); the reference gets them through the .bgi index, this reader by filtering."""
    path = os.path.join(REFT, "example.16bits.bgen")
    if not os.path.exists(path):
        pytest.skip("no /root/reference here")
    full = _dump(bg, path)
    assert full["compression"] == 1 and full["layout"] == 2
    a = _dump(bg, path, ("01", 2000, 4999))
    b = _dump(bg, path, ("01", 6000, 7999))
    want = [v for v in full["variants"] if v[0] == "01" and (2000 <= int(v[1]) < 5000 or 6000 <= int(v[1]) < 8000)]
    key = lambda v: (int(v[1]), v[2])                       # (the file is not sorted by position)
    assert sorted(a["variants"] + b["variants"], key=key) == sorted(want, key=key) and len(want) == 10
    assert a["variants"] == [v for v in want if int(v[1]) < 5000]
    ref = [l.split("\t") for l in open(os.path.join(REFT, "testBGenFileByRange.output.correct")) if not l.startswith("#")]
    assert sorted((r[0], int(r[1]), r[2]) for r in ref) == sorted((v[0], int(v[1]), v[2] + ("," + v[3] if v[3] else "")) for v in want)
