"""CPU: PLINK fileset reader of the ingestion layer (rvtests_b200/host/rvt_bed_file.h; reference: libVcf/PlinkInputFile.h:14-139
for the three files, PlinkInputFile.cpp:23-47 for the 2-bit codes): rows, AF / counts, marker keys, range selection, the
pushes it makes, and the error cases the reference aborts on."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def bf():
    d = os.path.join(ROOT, "tests", "hostcheck")
    so = os.path.join(d, "libbedfile_check.so")
    deps = [os.path.join(d, "bedfile_check.cpp"), os.path.join(ROOT, "rvtests_b200", "host", "rvt_bed_file.h"),
            os.path.join(ROOT, "rvtests_b200", "host", "rvt_vcf_pack.h")]
    if not os.path.exists(so) or any(os.path.getmtime(p) > os.path.getmtime(so) for p in deps):
        subprocess.run(["g++", "-O2", "-std=c++11", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"), deps[0], "-o", so], check=True)
    L = C.CDLL(so)
    L.bf_open.argtypes = [C.c_char_p]
    L.bf_error.restype = C.c_char_p
    L.bf_num_sample.restype = C.c_longlong
    L.bf_stride.restype = C.c_longlong
    L.bf_sample.restype = C.c_char_p
    L.bf_pheno.restype = C.c_double
    L.bf_chrom.restype = C.c_char_p
    L.bf_marker_index.argtypes = [C.c_char_p]
    L.bf_af.restype = C.c_double
    L.bf_af.argtypes = [C.c_int, C.c_void_p]
    L.bf_row.argtypes = [C.c_int, C.c_void_p]
    L.bf_select.argtypes = [C.c_char_p, C.c_void_p, C.c_int]
    L.bf_push.argtypes = [C.c_void_p, C.c_int]
    L.bf_pushed.argtypes = [C.c_void_p, C.c_void_p]
    return L


def write_fileset(prefix, Gt, miss, chrom, pos, ids, mode=1, magic=(0x6C, 0x1B), sep="\t"):
    from rvtests_b200.synth import pack_bed
    M, N = Gt.shape
    bed = pack_bed(Gt, miss)
    with open(prefix + ".bed", "wb") as f:
        f.write(bytes([magic[0], magic[1], mode]))
        f.write(bed.tobytes())
    with open(prefix + ".bim", "w") as f:
        for j in range(M):
            f.write(sep.join([str(chrom[j]), ids[j], "0", str(pos[j]), "A", "G"]) + "\n")
    with open(prefix + ".fam", "w") as f:
        for i in range(N):
            f.write(" ".join([f"F{i}", f"S{i}", "0", "0", str(1 + i % 2), "%.3f" % (0.5 * i)]) + "\n")
    return bed


@pytest.mark.parametrize("N", [1, 4, 7, 130])
def test_fileset_rows_af_and_selection(bf, oracle, tmp_path, N):
    rng = np.random.default_rng(N)
    M = 12
    Gt = rng.integers(0, 3, size=(M, N)).astype(np.int8)
    miss = rng.random((M, N)) < 0.1
    chrom = ["1"] * 5 + ["2"] * 4 + ["X"] * 3
    pos = [100, 200, 300, 400, 500, 100, 150, 900, 901, 5, 6, 7]
    ids = [f"rs{j}" if j % 3 else "." for j in range(M)]
    prefix = str(tmp_path / "set")
    bed = write_fileset(prefix, Gt, miss, chrom, pos, ids, sep="\t" if N % 2 else " ")
    assert bf.bf_open(prefix.encode()) == 0, bf.bf_error()
    assert bf.bf_num_sample() == N and bf.bf_num_marker() == M and bf.bf_stride() == (N + 3) // 4
    assert [bf.bf_sample(i).decode() for i in range(N)] == [f"S{i}" for i in range(N)]
    assert [bf.bf_sex(i) for i in range(N)] == [1 + i % 2 for i in range(N)]
    assert [bf.bf_pheno(i) for i in range(N)] == [float("%.3f" % (0.5 * i)) for i in range(N)]
    want = np.where(miss, -9.0, Gt.astype(float))
    for j in range(M):
        row = np.zeros((N + 3) // 4, dtype=np.uint8)
        bf.bf_row(j, row.ctypes.data)
        assert np.array_equal(row, bed[j])
        assert np.array_equal(oracle.bed_decode(row[None, :], N)[0], want[j])        # the reference's decode table
        cnt = (C.c_int * 4)()
        af = bf.bf_af(j, cnt)
        r0, r1, r2, rm, a = oracle.genotype_counter(want[j])
        assert list(cnt) == [r0, r1, r2, rm] and af == a
        assert bf.bf_pos(j) == pos[j] and bf.bf_chrom(j).decode() == chrom[j]
        key = ids[j] if ids[j] != "." else f"{chrom[j]}:{pos[j]}"
        assert bf.bf_marker_index(key.encode()) == j
    assert bf.bf_marker_index(b"nope") == -1
    sel = np.zeros(M, dtype=np.int32)
    assert bf.bf_select(b"1:200-400", sel.ctypes.data, M) == 3 and list(sel[:3]) == [1, 2, 3]
    assert bf.bf_select(b"2:100-150,X:1", sel.ctypes.data, M) == 5 and list(sel[:5]) == [5, 6, 9, 10, 11]
    assert bf.bf_select(b"3:1", sel.ctypes.data, M) == 0 and bf.bf_select(b"X", sel.ctypes.data, M) == 0   # ("X" alone does not conform)
    # pushes: consecutive rows leave the mapping as they are, scattered rows are gathered; AF rides along
    for rows in ([1, 2, 3], [0, 5, 11], [7]):
        r = np.array(rows, dtype=np.int32)
        assert bf.bf_push(r.ctypes.data, len(r)) == 0
        got = np.zeros((len(rows), (N + 3) // 4), dtype=np.uint8)
        af = np.zeros(len(rows))
        assert bf.bf_pushed(got.ctypes.data, af.ctypes.data) == len(rows)
        assert np.array_equal(got, bed[rows])
        assert np.array_equal(af, np.array([oracle.genotype_counter(want[j])[4] for j in rows]))


def test_fileset_errors(bf, tmp_path):
    rng = np.random.default_rng(0)
    Gt = rng.integers(0, 3, size=(3, 5)).astype(np.int8)
    miss = np.zeros((3, 5), dtype=bool)
    args = (Gt, miss, ["1"] * 3, [1, 2, 3], ["a", "b", "c"])
    p = str(tmp_path / "x")
    assert bf.bf_open((p + "missing").encode()) != 0
    write_fileset(p, *args, magic=(0x6C, 0x1C))
    assert bf.bf_open(p.encode()) != 0 and b"Magic" in bf.bf_error()
    write_fileset(p, *args, mode=0)
    assert bf.bf_open(p.encode()) != 0 and b"individual-major" in bf.bf_error()
    write_fileset(p, *args, mode=7)
    assert bf.bf_open(p.encode()) != 0 and b"Unrecognized" in bf.bf_error()
    write_fileset(p, Gt, miss, ["1"] * 3, [1, 2, 3], ["a", "b", "a"])
    assert bf.bf_open(p.encode()) != 0 and b"duplicated marker" in bf.bf_error()
    write_fileset(p, Gt, miss, ["1"] * 3, [1, 2, 2], [".", ".", "."])
    assert bf.bf_open(p.encode()) != 0 and b"duplicated marker" in bf.bf_error()
    write_fileset(p, *args)
    with open(p + ".bim", "a") as f:
        f.write("1 z 0 9 A\n")
    assert bf.bf_open(p.encode()) != 0 and b"bim" in bf.bf_error()
    write_fileset(p, *args)
    with open(p + ".fam", "a") as f:
        f.write("F9 S1 0 0 1 2.0\n")
    assert bf.bf_open(p.encode()) != 0 and b"duplicated person" in bf.bf_error()
    write_fileset(p, *args)
    with open(p + ".bed", "r+b") as f:
        f.truncate(3 + 2 * 2 + 1)
    assert bf.bf_open(p.encode()) != 0 and b"shorter" in bf.bf_error()
    write_fileset(p, *args)
    assert bf.bf_open(p.encode()) == 0
