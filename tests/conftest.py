import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _cuda_device_present():
    try:
        import torch
        return bool(torch.cuda.is_available())
    except Exception:
        return os.path.exists("/dev/nvidia0")


# per-test wall-clock limit of the GPU tests.  The device code has its own watchdogs (an mbarrier wait or a SKAT-O
# quadrature that exceeds its cycle budget ends the kernel with an error, csrc/sweep_tc.cuh / skato_tail.cuh); this is the
# backstop: pytest-timeout's thread method ends the PROCESS, because a host thread blocked in cudaStreamSynchronize never
# sees a signal.  One hung test then costs two minutes instead of the whole GPU run.
GPU_TEST_TIMEOUT_S = 150


def pytest_collection_modifyitems(config, items):
    have = _cuda_device_present()
    skip = pytest.mark.skip(reason="no CUDA device (the engine has no CPU fallback)")
    for it in items:
        if "gpu" not in it.keywords:
            continue
        if not have:
            it.add_marker(skip)
        elif it.get_closest_marker("timeout") is None:
            it.add_marker(pytest.mark.timeout(GPU_TEST_TIMEOUT_S, method="thread"))


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def hostcheck():
    """g++ build of the product's __host__ __device__ math headers (logic check on the CPU)."""
    import ctypes as C
    import subprocess
    d = os.path.join(ROOT, "tests", "hostcheck")
    so = os.path.join(d, "libhostcheck.so")
    src = os.path.join(d, "hostcheck.cpp")
    deps = [src] + [os.path.join(ROOT, "rvtests_b200", "csrc", f)
                    for f in os.listdir(os.path.join(ROOT, "rvtests_b200", "csrc")) if f.endswith(".cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(p) > os.path.getmtime(so) for p in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", src, "-o", so], check=True)
    H = C.CDLL(so)
    dp = C.POINTER(C.c_double)
    H.hc_gamma_q.restype = C.c_double
    H.hc_gamma_q.argtypes = [C.c_double, C.c_double]
    H.hc_chisq_q.restype = C.c_double
    H.hc_chisq_q.argtypes = [C.c_double, C.c_double]
    H.hc_beta_weight.restype = C.c_double
    H.hc_beta_weight.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int]
    H.hc_liu.restype = C.c_double
    H.hc_liu.argtypes = [dp, C.c_int, C.c_double]
    H.hc_mixchisq.restype = C.c_double
    H.hc_mixchisq.argtypes = [dp, C.c_int, C.c_double, C.POINTER(C.c_int)]
    H.hc_qf.restype = C.c_double
    H.hc_qf.argtypes = [dp, C.c_int, C.c_double, C.c_int, C.c_double, C.POINTER(C.c_int)]
    H.hc_eigen.restype = C.c_int
    H.hc_eigen.argtypes = [dp, C.c_int, dp]
    H.hc_chisq_qinv.restype = C.c_double
    H.hc_chisq_qinv.argtypes = [C.c_double, C.c_double]
    H.hc_qags.restype = C.c_int
    H.hc_qags.argtypes = [C.CFUNCTYPE(C.c_double, C.c_double), C.c_double, C.c_double, C.c_double, C.c_double, dp, dp,
                          C.POINTER(C.c_int)]
    H.hc_skato_tail.restype = C.c_int
    H.hc_skato_tail.argtypes = [dp, C.c_int, dp, C.c_double, dp]
    H.hc_qf_fast.restype = C.c_double
    H.hc_qf_fast.argtypes = [dp, C.c_int, C.c_double, C.c_int, C.c_double, C.POINTER(C.c_int)]
    H.hc_qf_fast_many.restype = None
    H.hc_qf_fast_many.argtypes = [dp, C.c_int, dp, C.c_int, dp, C.POINTER(C.c_int)]
    H.hc_skato_fast.restype = C.c_int
    H.hc_skato_fast.argtypes = [dp, C.c_int, dp, C.c_double, C.c_double, dp]
    H.hc_lfg_draws.restype = None
    H.hc_lfg_draws.argtypes = [C.c_uint, C.c_ulonglong, C.c_int, C.c_void_p]
    H.hc_fy_roots.restype = None
    H.hc_fy_roots.argtypes = [C.c_void_p, C.c_uint, C.c_void_p]
    H.hc_eigen_tridiag.restype = C.c_int
    H.hc_eigen_tridiag.argtypes = [dp, C.c_int, dp]
    return H


@pytest.fixture(scope="session")
def engine_cls():
    import rvtests_b200
    return rvtests_b200.GeneEngine


class VcfPacker:
    """ctypes view of the product's host-side VCF packer (rvtests_b200/host/rvt_vcf_pack.h) through
    tests/hostcheck/vcfpack_check.cpp."""

    def __init__(self, lib):
        import ctypes as C
        self.L = lib
        lib.vp_gt.argtypes = [C.c_char_p, C.c_int]
        lib.vp_header.argtypes = [C.c_char_p, C.c_char_p]
        lib.vp_set_range.argtypes = [C.c_char_p]
        lib.vp_add.argtypes = [C.c_char_p, C.c_int]
        lib.vp_stride.restype = C.c_longlong
        lib.vp_num_sample.restype = C.c_longlong
        lib.vp_get.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lib.vp_variant_name.restype = C.c_char_p
        lib.vp_variant_name.argtypes = [C.c_int]
        lib.vp_sample_name.restype = C.c_char_p
        lib.vp_sample_name.argtypes = [C.c_int]
        lib.vp_set_dosage_tag.argtypes = [C.c_char_p]
        lib.vp_gt_male02.argtypes = [C.c_char_p, C.c_int]
        lib.vp_parse_range.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        lib.vp_load_gene_file.argtypes = [C.c_char_p, C.c_char_p]
        lib.vp_load_range_file.argtypes = [C.c_char_p, C.c_char_p]
        lib.vp_map_name.restype = C.c_char_p
        lib.vp_map_contains.argtypes = [C.c_int, C.c_char_p, C.c_int]
        lib.vp_set_freq.argtypes = [C.c_double, C.c_double]
        lib.vp_count_alt.argtypes = [C.c_char_p, C.c_int, C.c_int]
        lib.vp_count_male_alt2.argtypes = [C.c_char_p, C.c_int, C.c_int]
        lib.vp_par_is_hemi.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]
        lib.vp_set_sex.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_char_p]
        lib.vp_get_dosages.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]

    def gt(self, s):
        b = s.encode("latin-1")
        return self.L.vp_gt(b, len(b))

    def header(self, line, keep=None):
        return self.L.vp_header(line.encode(), None if keep is None else "\n".join(keep).encode())

    def set_range(self, spec):
        return self.L.vp_set_range((spec or "").encode())

    def clear(self):
        self.L.vp_clear()

    def add(self, line):
        b = line.encode("latin-1")
        return self.L.vp_add(b, len(b))

    def sample_names(self):
        return [self.L.vp_sample_name(i).decode() for i in range(self.L.vp_num_sample())]

    def gt_male02(self, s):
        b = s.encode("latin-1")
        return self.L.vp_gt_male02(b, len(b))

    def count_alt(self, s, alt):
        b = s.encode("latin-1")
        return self.L.vp_count_alt(b, len(b), alt)

    def count_male_alt2(self, s, alt):
        b = s.encode("latin-1")
        return self.L.vp_count_male_alt2(b, len(b), alt)

    def parse_range(self, s):
        import ctypes as C
        chrom = C.create_string_buffer(64)
        b, e = C.c_int(0), C.c_int(0)
        rc = self.L.vp_parse_range(s.encode("latin-1"), chrom, C.byref(b), C.byref(e))
        return None if rc else (chrom.value.decode("latin-1"), b.value, e.value)

    def set_freq(self, lo=0.0, hi=0.0):
        self.L.vp_set_freq(float(lo), float(hi))

    def set_multi(self, on):
        self.L.vp_set_multi(1 if on else 0)

    def par_is_hemi(self, x_label, par_region, chrom, pos):
        return self.L.vp_par_is_hemi(x_label.encode(), par_region.encode(), chrom.encode(), int(pos))

    def set_sex(self, sex=None, x_label="", par_region=""):
        import numpy as np
        if sex is None:
            return self.L.vp_set_sex(None, 0, b"", b"")
        sx = np.ascontiguousarray(sex, dtype=np.int32)
        return self.L.vp_set_sex(sx.ctypes.data, len(sx), x_label.encode(), par_region.encode())

    def set_filters(self, gd=(-1, -1), gq=(-1, -1)):
        self.L.vp_set_filters(int(gd[0]), int(gd[1]), int(gq[0]), int(gq[1]))

    def set_dosage_tag(self, tag):
        self.L.vp_set_dosage_tag((tag or "").encode())

    def dosage_gene(self, raw=True):
        """dosage mode: (G (N, M) doubles, af (M,), counts (M, 4)); raw=False: after imputeDosagesToMean()"""
        import numpy as np
        m, n = self.L.vp_num_variant(), self.L.vp_num_sample()
        G = np.zeros((m, n))
        af = np.zeros(m)
        counts = np.zeros((m, 4), dtype=np.int32)
        if m:
            self.L.vp_get_dosages(G.ctypes.data, af.ctypes.data, counts.ctypes.data, 1 if raw else 0)
        return G.T.copy(), af, counts

    def gene(self):
        """(rows uint8 (M, stride), af (M,), counts (M, 4) = hom-ref / het / hom-alt / missing, names)"""
        import numpy as np
        m, st = self.L.vp_num_variant(), self.L.vp_stride()
        rows = np.zeros((m, st), dtype=np.uint8)
        af = np.zeros(m)
        counts = np.zeros((m, 4), dtype=np.int32)
        if m:
            self.L.vp_get(rows.ctypes.data, af.ctypes.data, counts.ctypes.data)
        return rows, af, counts, [self.L.vp_variant_name(j).decode() for j in range(m)]


@pytest.fixture(scope="session")
def vcfpack():
    import ctypes as C
    import subprocess
    d = os.path.join(ROOT, "tests", "hostcheck")
    so = os.path.join(d, "libvcfpack_check.so")
    deps = [os.path.join(d, "vcfpack_check.cpp"), os.path.join(ROOT, "rvtests_b200", "host", "rvt_vcf_pack.h"),
            os.path.join(ROOT, "include", "rvtests_b200.h")]
    if not os.path.exists(so) or any(os.path.getmtime(p) > os.path.getmtime(so) for p in deps):
        subprocess.run(["g++", "-O2", "-std=c++11", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"), deps[0], "-o", so],
                       check=True)
    return VcfPacker(C.CDLL(so))
