"""CPU: the BGZF writer + tabix index builder of rvtests_b200/host/rvt_bgzf.h (the bgzipped, tabix-indexed
`.assoc.gz` of ModelManager, src/ModelManager.cpp:285-327, src/TabixUtil.cpp) against
  * zlib / gzip: the file is a valid multi-member gzip whose payload is the text, every member carries the BC subfield with
    its own size, the 28-byte EOF marker ends it;
  * tabix 0.2.6 as vendored by the reference (third/tabix-0.2.6.tar.bz2 -> oracle/_ref/libtabix_ref.so): ti_index_build run on
    the file THIS writer produced gives the same index (names, configuration, bins -> chunks, linear index);
  * a random-access read through the index (query regions == a linear scan)."""
import ctypes as C
import gzip
import os
import struct
import subprocess
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_bgzf_check():
    d = os.path.join(ROOT, "tests", "hostcheck")
    so = os.path.join(d, "libbgzfcheck.so")
    src = os.path.join(d, "bgzf_check.cpp")
    hdr = os.path.join(ROOT, "rvtests_b200", "host", "rvt_bgzf.h")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++11", "-Wall", "-Werror", "-fPIC", "-shared", "-I", os.path.dirname(hdr), "-o", so, src, "-lz"])
    L = C.CDLL(so)
    L.bz_write_indexed.argtypes = [C.c_char_p, C.c_char_p, C.c_long, C.c_int]
    L.bz_printf_check.argtypes = [C.c_char_p]
    L.bz_reg2bin.argtypes = [C.c_uint, C.c_uint]
    return L


@pytest.fixture(scope="module")
def bz():
    return load_bgzf_check()


def _tabix_ref():
    p = os.path.join(ROOT, "oracle", "_ref", "libtabix_ref.so")
    if not os.path.exists(p):
        from oracle import oracle as O
        O.build()
    if not os.path.exists(p):
        return None
    L = C.CDLL(p)
    L.ti_index_build.argtypes = [C.c_char_p, C.c_void_p]
    return L


def _members(raw):
    """-> list of (file offset, member size, payload bytes)"""
    out, o = [], 0
    while o < len(raw):
        assert raw[o:o + 4] == b"\x1f\x8b\x08\x04", o
        xlen = struct.unpack_from("<H", raw, o + 10)[0]
        assert xlen == 6 and raw[o + 12:o + 16] == b"BC\x02\x00"
        bsize = struct.unpack_from("<H", raw, o + 16)[0] + 1
        data = zlib.decompress(raw[o + 18:o + bsize - 8], -15)
        crc, isize = struct.unpack_from("<II", raw, o + bsize - 8)
        assert isize == len(data) and crc == zlib.crc32(data)
        out.append((o, bsize, data))
        o += bsize
    return out


def _parse_tbi(path):
    b = gzip.open(path, "rb").read()
    assert b[:4] == b"TBI\x01"
    n_ref, = struct.unpack_from("<i", b, 4)
    conf = struct.unpack_from("<6i", b, 8)
    l_nm, = struct.unpack_from("<i", b, 32)
    names = b[36:36 + l_nm].split(b"\0")[:-1]
    o = 36 + l_nm
    refs = []
    for _ in range(n_ref):
        n_bin, = struct.unpack_from("<i", b, o)
        o += 4
        bins = {}
        for _ in range(n_bin):
            bn, n_chunk = struct.unpack_from("<Ii", b, o)
            o += 8
            bins[bn] = [struct.unpack_from("<QQ", b, o + 16 * k) for k in range(n_chunk)]
            o += 16 * n_chunk
        n_intv, = struct.unpack_from("<i", b, o)
        o += 4
        lin = list(struct.unpack_from("<%dQ" % n_intv, b, o))
        o += 8 * n_intv
        refs.append((bins, lin))
    assert o == len(b)
    return dict(conf=conf, names=names, refs=refs)


def _assoc_text(seed, n_lines, wide=False):
    rng = np.random.default_rng(seed)
    lines = ["##ProgramName=Rvtests", "##NullModelEstimates", "CHROM\tPOS\tREF\tALT\tN_INFORMATIVE\tAF\tU_STAT"]
    lines[2] = "#" + lines[2] if seed % 2 else lines[2].replace("CHROM", "#CHROM")
    for chrom in ("1", "2", "X"):
        pos = np.sort(rng.integers(1, 3_000_000 if not wide else 200_000_000, n_lines))
        for p in pos:
            pad = "," .join("%g" % x for x in rng.normal(size=int(rng.integers(1, 40))))
            lines.append(f"{chrom}\t{int(p)}\tA\tC\t500\t{rng.random():g}\t{pad}")
    return ("\n".join(lines) + "\n").encode()


@pytest.mark.parametrize("case", [(1, 400, 7, False), (2, 4000, 4096, False), (3, 2500, 100000, True)])
def test_bgzf_file_and_tabix_index(bz, tmp_path, case):
    seed, n_lines, piece, wide = case
    text = _assoc_text(seed, n_lines, wide)
    path = str(tmp_path / "out.assoc.gz")
    assert bz.bz_write_indexed(path.encode(), text, len(text), piece) == 0
    raw = open(path, "rb").read()
    assert gzip.decompress(raw) == text
    mem = _members(raw)
    assert mem[-1][1] == 28 and mem[-1][2] == b"" and all(len(m[2]) > 0 for m in mem[:-1])
    assert b"".join(m[2] for m in mem) == text
    if len(text) > 70000:
        assert len(mem) > 2
    mine = _parse_tbi(path + ".tbi")
    assert mine["conf"] == (0, 1, 2, 0, ord("#"), 0) and mine["names"] == [b"1", b"2", b"X"]
    # random access through the index == linear scan
    ustart, start_of = 0, {}
    for off, _size, data in mem:
        start_of[off] = ustart
        ustart += len(data)
    rows = [ln.split("\t") for ln in text.decode().splitlines() if not ln.startswith("#")]
    rng = np.random.default_rng(seed)
    for _ in range(30):
        tid = int(rng.integers(0, 3))
        name = ["1", "2", "X"][tid]
        hi = 3_000_000 if not wide else 200_000_000
        beg = int(rng.integers(0, hi))
        end = beg + int(rng.integers(1, hi // 10))
        want = [r for r in rows if r[0] == name and beg < int(r[1]) <= end]            # 0-based half-open [beg, end)
        bins, lin = mine["refs"][tid]
        min_off = lin[min(beg >> 14, len(lin) - 1)] if lin else 0
        cand = set()
        for k, shift in ((0, 29), (1, 26), (9, 23), (73, 20), (585, 17), (4681, 14)):
            for bn in range(k + (beg >> shift), k + ((end - 1) >> shift) + 1):
                for (u, v) in bins.get(bn, []):
                    if v > min_off:
                        cand.add((u, v))
        got = []
        for (u, v) in sorted(cand):
            a = start_of[u >> 16] + (u & 0xffff)
            b = start_of[v >> 16] + (v & 0xffff) if (v >> 16) in start_of else len(text)
            for ln in text[a:b].decode().splitlines():
                r = ln.split("\t")
                if r[0] == name and beg < int(r[1]) <= end:
                    got.append(r)
        assert sorted(map(tuple, got)) == sorted(map(tuple, want))
    # the reference's tabix on the same data file
    T = _tabix_ref()
    if T is None:
        pytest.skip("oracle/_ref/libtabix_ref.so not built (no /root/reference here)")
    d2 = tmp_path / "ref"
    d2.mkdir()
    p2 = str(d2 / "out.assoc.gz")
    open(p2, "wb").write(raw)
    conf = (C.c_int32 * 6)(0, 1, 2, 0, ord("#"), 0)
    assert T.ti_index_build(p2.encode(), conf) == 0
    ref = _parse_tbi(p2 + ".tbi")
    assert ref["conf"] == mine["conf"] and ref["names"] == mine["names"]
    for (rb, rl), (mb, ml) in zip(ref["refs"], mine["refs"]):
        assert rl == ml
        assert rb == mb


def test_printf_pieces_and_bins(bz, tmp_path):
    path = str(tmp_path / "p.assoc.gz")
    assert bz.bz_printf_check(path.encode()) == 0
    txt = gzip.open(path, "rb").read().decode().splitlines()
    assert txt[0] == "#CHROM\tPOS\tX" and len(txt) == 52 and txt[-1].startswith("3\t77\txxxx") and len(txt[-1]) == 10005
    t = _parse_tbi(path + ".tbi")
    assert t["names"] == [b"1", b"2", b"3"]
    assert bz.bz_reg2bin(0, 1) == 4681 and bz.bz_reg2bin(16383, 16385) == 585 and bz.bz_reg2bin(0, 1 << 29) == 0


def test_unsorted_input_reports_an_index_error_but_keeps_the_data(bz, tmp_path):
    text = b"#CHROM\tPOS\n1\t500\n1\t100\n"
    path = str(tmp_path / "u.assoc.gz")
    assert bz.bz_write_indexed(path.encode(), text, len(text), 5) == -2
    assert gzip.open(path, "rb").read() == text and not os.path.exists(path + ".tbi")


@pytest.mark.parametrize("n_cov", [0, 2])
def test_summary_header_matches_the_reference_block(bz, tmp_path, n_cov):
    """SummaryHeaderB200 (rvtests_b200/host/rvt_summary.h) against the '##' block the reference's own SummaryHeader
    (src/Summary.h) writes at the top of its MetaScore file -- the reference model layer run on the CPU
    (oracle/_ref/libdropin_ref.so with the stock fitters)."""
    from oracle import oracle as O
    if O.ref_dropin() is None:
        pytest.skip("oracle/_ref/libdropin_ref.so not built")
    rng = np.random.default_rng(17 + n_cov)
    N, nv = 403, 3
    G = rng.binomial(2, 0.3, (N, nv)).astype(np.float64)
    y = rng.normal(size=N) * 3 + 1
    cov = rng.normal(size=(N, n_cov))
    pos = np.array([10, 20, 30], dtype=np.int32)
    ref = O.dropin_run_meta_models(G, pos, cov, y, 1000, str(tmp_path / "ref"), use_b200=False)
    block = [c for c in ref["MetaScore"][0] if not c.startswith("##NullModel") and not c.startswith("## - ")]
    bz.sh_render.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_char_p, C.c_char_p, C.c_int]
    out = C.create_string_buffer(1 << 16)
    covf = np.asfortranarray(cov)
    n = bz.sh_render(N, n_cov, y.ctypes.data, covf.ctypes.data if n_cov else None, b"reference-build-under-test", out, len(out))
    assert n > 0
    assert out.value.decode().splitlines() == block


def _tabix_query_ref(T, path, region):
    T.ti_open.restype = C.c_void_p
    T.ti_open.argtypes = [C.c_char_p, C.c_char_p]
    T.ti_querys.restype = C.c_void_p
    T.ti_querys.argtypes = [C.c_void_p, C.c_char_p]
    T.ti_read.restype = C.c_char_p
    T.ti_read.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
    T.ti_iter_destroy.argtypes = [C.c_void_p]
    T.ti_close.argtypes = [C.c_void_p]
    t = T.ti_open(path.encode(), None)
    it = T.ti_querys(t, region.encode())
    out = []
    if it:
        n = C.c_int(0)
        while True:
            s = T.ti_read(t, it, C.byref(n))
            if s is None:
                break
            out.append(s.decode())
        T.ti_iter_destroy(it)
    T.ti_close(t)
    return out


def _query(bz, path, chrom, beg, end):
    bz.bz_query.restype = C.c_long
    bz.bz_query.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_char_p, C.c_long]
    out = C.create_string_buffer(1 << 24)
    n = bz.bz_query(path.encode(), chrom.encode(), beg, end, out, len(out))
    assert n >= -1, n
    return None if n == -1 else out.value.decode().splitlines()


def test_tabix_reader_on_own_files(bz, tmp_path):
    text = _assoc_text(5, 3000)
    path = str(tmp_path / "q.assoc.gz")
    assert bz.bz_write_indexed(path.encode(), text, len(text), 1000) == 0
    nb = C.c_long(0)
    bz.bz_count_lines.restype = C.c_long
    bz.bz_count_lines.argtypes = [C.c_char_p, C.POINTER(C.c_long)]
    assert bz.bz_count_lines(path.encode(), C.byref(nb)) == text.count(b"\n") and nb.value == len(text)
    rows = [ln for ln in text.decode().splitlines() if not ln.startswith("#")]
    T = _tabix_ref()
    rng = np.random.default_rng(1)
    assert _query(bz, path, "nope", 1, 10) is None
    for _ in range(25):
        name = ["1", "2", "X"][int(rng.integers(0, 3))]
        beg = int(rng.integers(1, 3_000_000))
        end = beg + int(rng.integers(0, 200_000))
        want = [r for r in rows if r.split("\t")[0] == name and beg <= int(r.split("\t")[1]) <= end]
        got = _query(bz, path, name, beg, end)
        assert got == want
        if T is not None:
            assert _tabix_query_ref(T, path, f"{name}:{beg}-{end}") == want


def test_tabix_reader_on_the_reference_example_vcf(bz):
    """example/example.vcf.gz + .tbi of the reference (written by the stock bgzip / tabix -p vcf): header lines and region
    queries against a plain gzip scan -- and against tabix 0.2.6's own ti_querys."""
    path = "/root/reference/example/example.vcf.gz"
    if not os.path.exists(path):
        pytest.skip("no /root/reference here")
    text = gzip.open(path, "rt").read().splitlines()
    bz.bz_header.restype = C.c_long
    bz.bz_header.argtypes = [C.c_char_p, C.c_char_p, C.c_long]
    out = C.create_string_buffer(1 << 20)
    assert bz.bz_header(path.encode(), out, len(out)) > 0
    assert out.value.decode().splitlines() == [l for l in text if l.startswith("#")]
    recs = [l for l in text if not l.startswith("#")]
    assert len(recs) > 0
    T = _tabix_ref()
    pos = sorted(int(r.split("\t")[1]) for r in recs)
    chrom = recs[0].split("\t")[0]
    for beg, end in [(1, 10 ** 8), (pos[0], pos[0]), (pos[1], pos[-2]), (pos[-1] + 1, pos[-1] + 5), (pos[2] - 1, pos[2] - 1)]:
        want = [r for r in recs if r.split("\t")[0] == chrom and int(r.split("\t")[1]) <= end
                and int(r.split("\t")[1]) + len(r.split("\t")[3]) - 1 >= beg]
        assert _query(bz, path, chrom, beg, end) == want
        if T is not None:
            assert _tabix_query_ref(T, path, f"{chrom}:{beg}-{end}") == want
