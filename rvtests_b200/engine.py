"""ctypes mirror of include/rvtests_b200.h (names, argument meaning and error behaviour follow
the C ABI one to one).  No arithmetic happens here."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

RVT_OK = 0
ENGINE_AUTO, ENGINE_SIMT, ENGINE_TC = 0, 1, 2
GENE_OK, GENE_NA, GENE_BADFLAGS, GENE_BADVALUE = 0, 2, 4, 5


class RvtError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"rvtests_b200 error {code}: {msg}")
        self.code = code


class GeneResult(C.Structure):
    """struct rvt_gene_result"""
    _fields_ = [
        ("Q", C.c_double), ("p_skat", C.c_double), ("p_davies", C.c_double), ("p_liu", C.c_double),
        ("davies_fault", C.c_int32), ("n_lambda", C.c_int32), ("m_poly", C.c_int32),
        ("status", C.c_int32), ("cmc_nonref", C.c_int32), ("cmc_ok", C.c_int32),
        ("cmc_U", C.c_double), ("cmc_V", C.c_double), ("cmc_stat", C.c_double), ("cmc_p", C.c_double),
        ("zeg_ok", C.c_int32), ("skato_ok", C.c_int32),
        ("zeg_U", C.c_double), ("zeg_V", C.c_double), ("zeg_stat", C.c_double), ("zeg_p", C.c_double),
        ("skato_Q", C.c_double), ("skato_rho", C.c_double), ("skato_p", C.c_double),
        ("lambda_max", C.c_double),
    ]


RESULT_DTYPE = np.dtype([
    ("Q", "f8"), ("p_skat", "f8"), ("p_davies", "f8"), ("p_liu", "f8"),
    ("davies_fault", "i4"), ("n_lambda", "i4"), ("m_poly", "i4"), ("status", "i4"),
    ("cmc_nonref", "i4"), ("cmc_ok", "i4"),
    ("cmc_U", "f8"), ("cmc_V", "f8"), ("cmc_stat", "f8"), ("cmc_p", "f8"),
    ("zeg_ok", "i4"), ("skato_ok", "i4"),
    ("zeg_U", "f8"), ("zeg_V", "f8"), ("zeg_stat", "f8"), ("zeg_p", "f8"),
    ("skato_Q", "f8"), ("skato_rho", "f8"), ("skato_p", "f8"), ("lambda_max", "f8"),
])
assert RESULT_DTYPE.itemsize == C.sizeof(GeneResult)

EXPORTS = [
    "rvt_ctx_create", "rvt_ctx_destroy", "rvt_last_error", "rvt_set_option", "rvt_get_info",
    "rvt_set_stream",
    "rvt_set_null_model", "rvt_set_null_model_dev", "rvt_get_null_model", "rvt_set_null_residual",
    "rvt_gene_push_f64", "rvt_gene_push_i8", "rvt_gene_push_dev_i8", "rvt_gene_push_bed", "rvt_pending",
    "rvt_flush", "rvt_flush_dev", "rvt_synth_load", "rvt_loaded_genes", "rvt_run_loaded", "rvt_push_loaded",
    "rvt_loaded_read", "rvt_last_timing", "rvt_debug_partials",
    "rvt_debug_phases",
    "rvt_meta_plan", "rvt_meta_flush", "rvt_meta_binary_extras", "rvt_perm_results", "rvt_perm_debug_q", "rvt_debug_rand", "rvt_lmm_set_null", "rvt_lmm_flush", "rvt_lmm_meta_flush", "rvt_get_null_beta", "rvt_bolt_fit_null", "rvt_bolt_fit_null_sharded",
]

# rvt_allreduce_fn (include/rvtests_b200.h): int (*)(void* user, double* buf, int64_t count, void* cuda_stream)
ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p)
BOLT_DTYPE = np.dtype([("delta", "f8"), ("sigma2_g", "f8"), ("sigma2_e", "f8"), ("h2", "f8"), ("h_inv_y_norm2", "f8"),
                       ("inf_stat_calibration", "f8"), ("xvx_xx_ratio", "f8"), ("log_delta", "f8", (7,)), ("f", "f8", (7,)),
                       ("mc_trials", "i4"), ("reml_evals", "i4"), ("cg_iterations", "i4"), ("n_covariates_kept", "i4"),
                       ("ms_xtv", "f8"), ("ms_xw", "f8"), ("h_products", "i4"), ("allreduce_calls", "i4"),
                       ("h_products_calibration", "i4"), ("pad", "i4")])

LMM_DTYPE = np.dtype([("af", "f8"), ("U", "f8"), ("V", "f8"), ("stat", "f8"), ("pvalue", "f8"), ("ok", "i4"), ("pad", "i4")])

PERM_DTYPE = np.dtype([
    ("num_perm", "i4"), ("actual_perm", "i4"), ("num_greater", "i4"), ("num_equal", "i4"),
    ("stat", "f8"), ("p_perm", "f8"), ("stream_pos", "i8"), ("done", "i4"), ("stream_ok", "i4"),
])

CC_DTYPE = np.dtype([("n", "i4", 2), ("n_ref", "i4", 2), ("n_het", "i4", 2), ("n_alt", "i4", 2), ("hwe_p", "f8", 2)])
VARIANT_DTYPE = np.dtype([
    ("af", "f8"), ("ac", "f8"), ("call_rate", "f8"), ("hwe_p", "f8"),
    ("n_ref", "i4"), ("n_het", "i4"), ("n_alt", "i4"), ("ok", "i4"), ("polymorphic", "i4"), ("pad", "i4"),
    ("U", "f8"), ("sqrtV", "f8"), ("effect", "f8"), ("effect_se", "f8"), ("pvalue", "f8"),
])

_lib = None
_dp = C.POINTER(C.c_double)


def load_library(rebuild: bool = False):
    """dlopen the in-tree librvtests_b200.so (building it first when sources are newer)."""
    global _lib
    if _lib is not None and not rebuild:
        return _lib
    path = _build.build_lib(force=rebuild)
    # diagnostics only (A/B timing of build variants, tools/): an alternate build of the same sources
    path = os.environ.get("RVT_B200_LIB_VARIANT", path)
    L = C.CDLL(path)
    vp = C.c_void_p
    L.rvt_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.rvt_ctx_destroy.argtypes = [vp]
    L.rvt_ctx_destroy.restype = None
    L.rvt_last_error.argtypes = [vp]
    L.rvt_last_error.restype = C.c_char_p
    L.rvt_set_option.argtypes = [vp, C.c_char_p, C.c_double]
    L.rvt_get_info.argtypes = [vp, C.c_char_p]
    L.rvt_get_info.restype = C.c_double
    L.rvt_set_stream.argtypes = [vp, vp]
    L.rvt_set_null_model.argtypes = [vp, C.c_int64, C.c_int, _dp, _dp, C.c_int]
    L.rvt_set_null_model_dev.argtypes = [vp, C.c_int64, C.c_int, vp, vp]
    L.rvt_set_null_residual.argtypes = [vp, C.c_int64, C.c_int, _dp, _dp, C.c_double]
    L.rvt_get_null_model.argtypes = [vp, _dp, _dp, _dp]
    L.rvt_gene_push_f64.argtypes = [vp, _dp, C.c_int, _dp]
    L.rvt_gene_push_i8.argtypes = [vp, vp, C.c_int, C.c_int64, _dp]
    L.rvt_gene_push_dev_i8.argtypes = [vp, vp, C.c_int, C.c_int64, _dp, vp]
    L.rvt_gene_push_bed.argtypes = [vp, vp, C.c_int, C.c_int64, _dp]
    L.rvt_pending.argtypes = [vp]
    L.rvt_flush.argtypes = [vp, vp, C.c_int, C.POINTER(C.c_int)]
    L.rvt_flush_dev.argtypes = [vp, vp, C.c_int, C.POINTER(C.c_int)]
    L.rvt_synth_load.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp]
    L.rvt_loaded_genes.argtypes = [vp]
    L.rvt_run_loaded.argtypes = [vp, vp, C.c_int, C.POINTER(C.c_int), C.c_int]
    L.rvt_push_loaded.argtypes = [vp]
    L.rvt_meta_binary_extras.argtypes = [vp, vp, vp, vp, C.c_int64]
    L.rvt_lmm_meta_flush.argtypes = [vp, vp, vp, C.c_int64, vp, C.c_int64, vp, C.c_int64, C.POINTER(C.c_int)]
    L.rvt_loaded_read.argtypes = [vp, C.c_int64, C.c_int, vp]
    L.rvt_last_timing.argtypes = [vp, _dp]
    L.rvt_debug_partials.argtypes = [vp, vp, C.c_int64, C.POINTER(C.c_int64)]
    L.rvt_debug_phases.argtypes = [vp, vp, C.c_int]
    L.rvt_meta_plan.argtypes = [vp, vp, vp, C.c_int64, C.c_int64, C.POINTER(C.c_int)]
    L.rvt_meta_flush.argtypes = [vp, vp, vp, C.c_int64, vp, C.c_int64, vp, C.c_int64, C.POINTER(C.c_int)]
    L.rvt_perm_results.argtypes = [vp, vp, C.c_int, C.POINTER(C.c_int)]
    L.rvt_perm_debug_q.argtypes = [vp, vp, C.c_int, C.POINTER(C.c_int)]
    L.rvt_debug_rand.argtypes = [vp, C.c_uint32, C.c_uint64, C.c_int64, vp]
    L.rvt_lmm_set_null.argtypes = [vp, C.c_int64, C.c_int, vp, vp, C.c_double, C.c_double, vp, vp]
    L.rvt_lmm_flush.argtypes = [vp, vp, C.c_int64]
    L.rvt_bolt_fit_null.argtypes = [vp, vp, C.c_int64, C.c_int64, C.c_int64, _dp, _dp, C.c_int, C.c_int, vp, vp, vp]
    L.rvt_bolt_fit_null_sharded.argtypes = [vp, vp, C.c_int64, C.c_int64, C.c_int64, _dp, _dp, C.c_int, C.c_int, C.c_int64, C.c_int64,
                                            ALLREDUCE_FN, vp, vp, vp, vp]
    _lib = L
    return L


def _pd(a):
    return a.ctypes.data_as(_dp) if a is not None else None


class GeneEngine:
    """One context on one GPU.  Mirrors the call sequence a ModelFitter adapter makes:
    set_null_model (once) -> push genes -> flush."""

    def __init__(self, device: int = 0):
        self.L = load_library()
        self.h = C.c_void_p()
        rc = self.L.rvt_ctx_create(device, C.byref(self.h))
        if rc != RVT_OK:
            msg = self.L.rvt_last_error(self.h).decode() if self.h else "context allocation failed"
            if self.h:
                self.L.rvt_ctx_destroy(self.h)
                self.h = C.c_void_p()
            raise RvtError(rc, msg)
        self.N = 0
        self.C = 0

    def close(self):
        if getattr(self, "h", None):
            self.L.rvt_ctx_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc != RVT_OK:
            raise RvtError(rc, self.L.rvt_last_error(self.h).decode())

    def set_option(self, key: str, value: float):
        self._chk(self.L.rvt_set_option(self.h, key.encode(), float(value)))

    def set_stream(self, cuda_stream_ptr):
        self._chk(self.L.rvt_set_stream(self.h, cuda_stream_ptr))

    def info(self, key: str) -> float:
        return self.L.rvt_get_info(self.h, key.encode())

    def set_null_model(self, X, y, binary: bool = False):
        """X: (N, C) incl. intercept column 0; y: (N,)"""
        Xc = np.asfortranarray(X, dtype=np.float64)
        yc = np.ascontiguousarray(y, dtype=np.float64)
        self.N, self.C = Xc.shape
        self._chk(self.L.rvt_set_null_model(self.h, self.N, self.C, _pd(Xc), _pd(yc), int(binary)))

    def set_null_residual(self, X, resid, sigma2):
        """caller-supplied score vector and variance scale (mixed-model score step)"""
        Xc = np.asfortranarray(X, dtype=np.float64)
        rc_ = np.ascontiguousarray(resid, dtype=np.float64)
        self.N, self.C = Xc.shape
        self._chk(self.L.rvt_set_null_residual(self.h, self.N, self.C, _pd(Xc), _pd(rc_), float(sigma2)))

    def set_null_model_dev(self, N, Cc, dX_ptr, dy_ptr):
        self.N, self.C = int(N), int(Cc)
        self._chk(self.L.rvt_set_null_model_dev(self.h, self.N, self.C, dX_ptr, dy_ptr))

    def get_null_model(self, want_resid=True):
        resid = np.zeros(self.N) if want_resid else None
        s2 = C.c_double(0)
        xi = np.zeros((self.C, self.C))
        self._chk(self.L.rvt_get_null_model(self.h, _pd(resid), C.byref(s2), _pd(xi)))
        return dict(resid=resid, sigma2=s2.value, xtx_inv=xi)

    def push_f64(self, G, af=None):
        """G: (N, M) doubles as dc->getGenotype() (any layout; sent column-major)."""
        Gc = np.asfortranarray(G, dtype=np.float64)
        afc = None if af is None else np.ascontiguousarray(af, dtype=np.float64)
        self._chk(self.L.rvt_gene_push_f64(self.h, _pd(Gc), Gc.shape[1], _pd(afc)))

    def push_i8(self, Gt, af=None):
        """Gt: (M, N) int8 variant-major hard calls."""
        Gc = np.ascontiguousarray(Gt, dtype=np.int8)
        afc = None if af is None else np.ascontiguousarray(af, dtype=np.float64)
        self._chk(self.L.rvt_gene_push_i8(self.h, Gc.ctypes.data, Gc.shape[0], Gc.shape[1], _pd(afc)))

    def push_bed(self, bed, af=None):
        """bed: (M, >= ceil(N/4)) uint8 PLINK SNP-major rows (00 -> 0, 10 -> 1, 11 -> 2, 01 -> missing)."""
        assert bed.dtype == np.uint8 and bed.ndim == 2 and bed.strides[1] == 1, (bed.dtype, bed.shape, bed.strides)
        afc = None if af is None else np.ascontiguousarray(af, dtype=np.float64)
        self._chk(self.L.rvt_gene_push_bed(self.h, bed.ctypes.data, bed.shape[0], bed.strides[0], _pd(afc)))

    def push_dev_i8(self, dptr, M, ld, af=None, flags=None):
        afc = None if af is None else np.ascontiguousarray(af, dtype=np.float64)
        fl = None if flags is None else np.ascontiguousarray(flags, dtype=np.uint8)
        self._chk(self.L.rvt_gene_push_dev_i8(self.h, dptr, int(M), int(ld), _pd(afc),
                                              None if fl is None else fl.ctypes.data))

    def pending(self) -> int:
        return self.L.rvt_pending(self.h)

    def flush(self):
        n = self.pending()
        out = np.zeros(max(n, 1), dtype=RESULT_DTYPE)
        got = C.c_int(0)
        self._chk(self.L.rvt_flush(self.h, out.ctypes.data, len(out), C.byref(got)))
        return out[: got.value]

    def bolt_fit_null(self, bed, N, y, covar, mc_trials=0, M_total=None, m_offset=0, allreduce=None, bed_dev=None):
        """BoltLMM null fit on a PLINK panel: bed (M, >= ceil(N/4)) uint8 SNP-major rows, covar (N, C) with the intercept
        first.  Returns (record, h [N + C'], Z [N, C']).
        Sharded over ranks (rvt_bolt_fit_null_sharded): bed holds rows [m_offset, m_offset + M) of M_total and `allreduce`
        is a Python callable (dev_ptr, count, cuda_stream) -> 0 that sums `count` doubles in place over the ranks
        (rvtests_b200.sharding.torch_allreduce).  bed_dev = (device_ptr, M, stride): a panel already in device memory."""
        if bed_dev is not None:
            bed_ptr, M, stride = int(bed_dev[0]), int(bed_dev[1]), int(bed_dev[2])
        else:
            assert bed.dtype == np.uint8 and bed.ndim == 2 and bed.strides[1] == 1
            bed_ptr, M, stride = bed.ctypes.data, bed.shape[0], bed.strides[0]
        y = np.ascontiguousarray(y, dtype=np.float64)
        cv = np.asfortranarray(covar, dtype=np.float64)
        Cc = cv.shape[1]
        rec = np.zeros(1, dtype=BOLT_DTYPE)
        h = np.zeros(N + Cc)
        Z = np.zeros((N, Cc), order="F")
        if M_total is None and allreduce is None:
            self._chk(self.L.rvt_bolt_fit_null(self.h, bed_ptr, M, stride, int(N), _pd(y), _pd(cv), Cc,
                                               int(mc_trials), rec.ctypes.data, h.ctypes.data, Z.ctypes.data))
        else:
            def _cb(_user, buf, count, stream):
                try:
                    return int(allreduce(int(buf), int(count), int(stream or 0)) or 0)
                except Exception as e:            # an exception must not unwind through the C frames
                    self._cb_error = e
                    return 1
            cb = ALLREDUCE_FN(_cb) if allreduce is not None else C.cast(None, ALLREDUCE_FN)
            self._cb_error = None
            rc = self.L.rvt_bolt_fit_null_sharded(self.h, bed_ptr, M, stride, int(N), _pd(y), _pd(cv), Cc, int(mc_trials),
                                                  int(M if M_total is None else M_total), int(m_offset), cb, None,
                                                  rec.ctypes.data, h.ctypes.data, Z.ctypes.data)
            if self._cb_error is not None:
                raise self._cb_error
            self._chk(rc)
        k = int(rec[0]["n_covariates_kept"])
        return rec[0], h[: N + k], np.ascontiguousarray(Z[:, :k])

    def lmm_set_null(self, U, lam, delta, sigma2, u_resid, ux):
        """FastLMM score step: U (N, N) with eigenvectors in COLUMNS, lam (N,), uResid (N,), ux (N, C)"""
        U = np.asfortranarray(U, dtype=np.float32)
        lam = np.ascontiguousarray(lam, dtype=np.float32)
        u_resid = np.ascontiguousarray(u_resid, dtype=np.float32)
        ux = np.asfortranarray(ux, dtype=np.float32)
        N = U.shape[0]
        assert U.shape == (N, N) and lam.shape == (N,) and u_resid.shape == (N,) and ux.shape[0] == N
        self._chk(self.L.rvt_lmm_set_null(self.h, N, ux.shape[1], U.ctypes.data, lam.ctypes.data, float(delta), float(sigma2),
                                          u_resid.ctypes.data, ux.ctypes.data))

    def lmm_flush(self, n_variants):
        """one record per pushed variant (blocks of <= 64 variants pushed with push_i8 / push_bed)"""
        out = np.zeros(max(int(n_variants), 1), dtype=LMM_DTYPE)
        self._chk(self.L.rvt_lmm_flush(self.h, out.ctypes.data, len(out)))
        return out[: int(n_variants)]

    def lmm_meta_flush(self, n_variants, pos, chrom, window_bp):
        """score records + the MetaCovFamQtl covariance band -> (records, band (nv, wmax+1), wmax)"""
        p = np.ascontiguousarray(pos, dtype=np.int32)
        c = np.ascontiguousarray(chrom, dtype=np.int32)
        wmax = C.c_int(0)
        self._chk(self.L.rvt_meta_plan(self.h, p.ctypes.data, c.ctypes.data, len(p), int(window_bp), C.byref(wmax)))
        out = np.zeros(max(int(n_variants), 1), dtype=LMM_DTYPE)
        band = np.zeros((int(n_variants), wmax.value + 1))
        self._chk(self.L.rvt_lmm_meta_flush(self.h, p.ctypes.data, c.ctypes.data, int(window_bp), out.ctypes.data, len(out),
                                            band.ctypes.data, band.size, C.byref(wmax)))
        return out[: int(n_variants)], band, wmax.value

    def perm_results(self):
        """permutation records (rvt_perm_result) of the genes of the last flush / run_loaded"""
        got = C.c_int(0)
        self._chk(self.L.rvt_perm_results(self.h, None, 0, C.byref(got)))
        out = np.zeros(max(got.value, 1), dtype=PERM_DTYPE)
        self._chk(self.L.rvt_perm_results(self.h, out.ctypes.data, len(out), C.byref(got)))
        return out[: got.value]

    def debug_rand(self, n, seed=1, pos=0):
        out = np.zeros(max(int(n), 1), dtype=np.int32)
        self._chk(self.L.rvt_debug_rand(self.h, int(seed), int(pos), int(n), out.ctypes.data))
        return out[: int(n)]

    def perm_debug_q(self):
        got = C.c_int(0)
        self._chk(self.L.rvt_perm_debug_q(self.h, None, 0, C.byref(got)))
        out = np.zeros(max(got.value, 1))
        self._chk(self.L.rvt_perm_debug_q(self.h, out.ctypes.data, len(out), C.byref(got)))
        return out[: got.value]

    def flush_dev(self, d_out_ptr, cap):
        got = C.c_int(0)
        self._chk(self.L.rvt_flush_dev(self.h, d_out_ptr, int(cap), C.byref(got)))
        return got.value

    def synth_load(self, keys, t0, t1, n_genes, M):
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        t0 = np.ascontiguousarray(t0, dtype=np.uint32)
        t1 = np.ascontiguousarray(t1, dtype=np.uint32)
        assert len(keys) == n_genes * M
        self._chk(self.L.rvt_synth_load(self.h, int(n_genes), int(M), keys.ctypes.data, t0.ctypes.data,
                                        t1.ctypes.data))

    def run_loaded(self, d_out_ptr=None):
        n = self.L.rvt_loaded_genes(self.h)
        got = C.c_int(0)
        if d_out_ptr is None:
            out = np.zeros(max(n, 1), dtype=RESULT_DTYPE)
            self._chk(self.L.rvt_run_loaded(self.h, out.ctypes.data, len(out), C.byref(got), 0))
            return out[: got.value]
        self._chk(self.L.rvt_run_loaded(self.h, d_out_ptr, n, C.byref(got), 1))
        return got.value

    def loaded_read(self, row0, rows):
        out = np.zeros((rows, self.N), dtype=np.int8)
        self._chk(self.L.rvt_loaded_read(self.h, int(row0), int(rows), out.ctypes.data))
        return out

    def debug_partials(self):
        """raw SweepPartial records of the last batch: dict(d=(n,64,96) int32, coll=(n,66) int64)"""
        nb = C.c_int64(0)
        self._chk(self.L.rvt_debug_partials(self.h, None, 0, C.byref(nb)))
        rec = np.dtype([("d", "i4", (64, 96)), ("coll", "i8", (66,)), ("pad", "i8", (2,))])
        buf = np.zeros(nb.value // rec.itemsize, dtype=rec)
        self._chk(self.L.rvt_debug_partials(self.h, buf.ctypes.data, buf.nbytes, C.byref(nb)))
        return buf

    def meta_flush(self, n_variants, pos=None, chrom=None, window_bp=1_000_000, want_cov=True):
        """--meta score[,cov] over the pending variant blocks.  Returns (variant records, band, wmax)."""
        vout = np.zeros(max(n_variants, 1), dtype=VARIANT_DTYPE)
        wmax = C.c_int(0)
        band = None
        p = c = None
        if want_cov:
            p = np.ascontiguousarray(pos, dtype=np.int32)
            c = np.ascontiguousarray(chrom, dtype=np.int32)
            self._chk(self.L.rvt_meta_plan(self.h, p.ctypes.data, c.ctypes.data, len(p), int(window_bp), C.byref(wmax)))
            band = np.zeros((n_variants, wmax.value + 1))
        self._chk(self.L.rvt_meta_flush(self.h, None if p is None else p.ctypes.data, None if c is None else c.ctypes.data,
                                        int(window_bp), vout.ctypes.data, len(vout),
                                        None if band is None else band.ctypes.data, 0 if band is None else band.size,
                                        C.byref(wmax)))
        return vout[:n_variants], band, wmax.value

    def push_loaded(self):
        self._chk(self.L.rvt_push_loaded(self.h))

    def meta_binary_extras(self, n_variants, n_cov):
        """case / control counts, covXZ (nv, C) and covZZ (C, C) of the last binary-trait meta_flush"""
        cc = np.zeros(max(int(n_variants), 1), dtype=CC_DTYPE)
        xz = np.zeros((max(int(n_variants), 1), int(n_cov)))
        zz = np.zeros((int(n_cov), int(n_cov)))
        self._chk(self.L.rvt_meta_binary_extras(self.h, cc.ctypes.data, xz.ctypes.data, zz.ctypes.data, len(cc)))
        return cc[: int(n_variants)], xz[: int(n_variants)], zz

    def meta_flush_dev(self, n_variants, pos, chrom, window_bp, d_vout, d_band, band_elems):
        """as meta_flush with the outputs left in device memory (d_vout: n_variants records, d_band: band_elems doubles)"""
        p = np.ascontiguousarray(pos, dtype=np.int32)
        c = np.ascontiguousarray(chrom, dtype=np.int32)
        wmax = C.c_int(0)
        self._chk(self.L.rvt_meta_flush(self.h, p.ctypes.data, c.ctypes.data, int(window_bp), C.c_void_p(d_vout), int(n_variants),
                                        C.c_void_p(d_band), int(band_elems), C.byref(wmax)))
        return wmax.value

    def debug_phases(self, n):
        out = np.zeros((n, 6), dtype=np.int64)
        self._chk(self.L.rvt_debug_phases(self.h, out.ctypes.data, n))
        return out

    def last_timing(self):
        t = np.zeros(4)
        self._chk(self.L.rvt_last_timing(self.h, _pd(t)))
        return dict(sweep_ms=t[0], finalize_ms=t[1], total_ms=t[2], launches=int(t[3]))
