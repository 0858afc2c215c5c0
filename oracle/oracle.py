"""oracle/oracle.py -- TEST INFRASTRUCTURE ONLY (never imported by the product package).

ctypes doors onto
  * oracle/liboracle.so                 our C restatement (oracle/skat_oracle.c)
  * oracle/_ref/libmixchisq_ref.so      the REFERENCE's own Davies/Liu code, compiled in place
  * oracle/_ref/libgsl_ref.so           GSL 1.16 from the tarball the reference vendors
plus the numpy/scipy restatement of SKAT-O (regression/SkatO.cpp) and of the synthetic-data
stream of SURVEY.md section 8(d) (so that the oracle can regenerate on the host exactly the
genotypes the device generator wrote).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_dbl_p = C.POINTER(C.c_double)
_int_p = C.POINTER(C.c_int)


def _p(a, t=C.c_double):
    return a.ctypes.data_as(C.POINTER(t))


class SkatOut(C.Structure):
    _fields_ = [("Q", C.c_double), ("pvalue", C.c_double), ("p_davies", C.c_double),
                ("p_liu", C.c_double), ("fault", C.c_int), ("n_lambda", C.c_int)]


class GeneOut(C.Structure):
    _fields_ = [("m_poly", C.c_int), ("status", C.c_int), ("skat", SkatOut),
                ("cmc_nonref", C.c_int),
                ("cmc_U", C.c_double), ("cmc_V", C.c_double), ("cmc_stat", C.c_double),
                ("cmc_p", C.c_double), ("cmc_ok", C.c_int),
                ("zeg_U", C.c_double), ("zeg_V", C.c_double), ("zeg_stat", C.c_double),
                ("zeg_p", C.c_double), ("zeg_ok", C.c_int)]


def build(native: bool = False) -> str:
    """Compile the checker (make -C oracle).  `native` additionally builds liboracle_native.so
    with -march=native on THIS machine (used by the CPU-baseline timing legs)."""
    targets = ["all"] + (["liboracle_native.so"] if native else [])
    # make's chatter goes to stderr: bench.py's stdout is ONE JSON line
    subprocess.run(["make", "-s", "-C", _HERE] + targets, check=True, stdout=sys.stderr)
    return os.path.join(_HERE, "liboracle_native.so" if native else "liboracle.so")


_lib_cache = {}


def lib(native: bool = False):
    key = "native" if native else "portable"
    if key in _lib_cache:
        return _lib_cache[key]
    path = os.path.join(_HERE, "liboracle_native.so" if native else "liboracle.so")
    if not os.path.exists(path):
        build(native)
    L = C.CDLL(path)
    L.orc_qf.restype = C.c_double
    L.orc_qf.argtypes = [_dbl_p, _dbl_p, _int_p, C.c_int, C.c_double, C.c_double, C.c_int,
                         C.c_double, _dbl_p, _int_p]
    L.orc_mixchisq_pvalue.restype = C.c_double
    L.orc_mixchisq_pvalue.argtypes = [_dbl_p, C.c_int, C.c_double, _int_p]
    L.orc_liu_pvalue.restype = C.c_double
    L.orc_liu_pvalue.argtypes = [_dbl_p, C.c_int, C.c_double]
    L.orc_gamma_q.restype = C.c_double
    L.orc_gamma_q.argtypes = [C.c_double, C.c_double]
    L.orc_chisq_q.restype = C.c_double
    L.orc_chisq_q.argtypes = [C.c_double, C.c_double]
    L.orc_beta_pdf.restype = C.c_double
    L.orc_beta_pdf.argtypes = [C.c_double] * 3
    L.orc_skat_weight.restype = C.c_double
    L.orc_skat_weight.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int]
    L.orc_sym_eigenvalues.restype = None
    L.orc_sym_eigenvalues.argtypes = [C.c_int, _dbl_p, _dbl_p]
    L.orc_fit_null_linear.restype = C.c_int
    L.orc_fit_null_linear.argtypes = [C.c_int64, C.c_int, _dbl_p, _dbl_p, _dbl_p, _dbl_p, _dbl_p,
                                      _dbl_p]
    L.orc_flip_minor_polymorphic.restype = C.c_int
    L.orc_flip_minor_polymorphic.argtypes = [C.c_int64, C.c_int, _dbl_p, _dbl_p, _int_p, _int_p]
    for name in ("orc_skat_reduced64", "orc_skat_faithful32"):
        f = getattr(L, name)
        f.restype = C.c_int
    L.orc_skat_reduced64.argtypes = [C.c_int64, C.c_int, C.c_int, _dbl_p, _dbl_p, _dbl_p, _dbl_p,
                                     _dbl_p, C.POINTER(SkatOut), _dbl_p]
    L.orc_skat_faithful32.argtypes = [C.c_int, C.c_int, C.c_int, _dbl_p, _dbl_p, _dbl_p, _dbl_p,
                                      _dbl_p, C.POINTER(SkatOut), _dbl_p]
    L.orc_cmc_collapse.restype = None
    L.orc_cmc_collapse.argtypes = [C.c_int64, C.c_int, _dbl_p, _dbl_p]
    L.orc_zeggini_collapse.restype = None
    L.orc_zeggini_collapse.argtypes = [C.c_int64, C.c_int, _dbl_p, _dbl_p]
    L.orc_nonref_sites.restype = C.c_int
    L.orc_nonref_sites.argtypes = [C.c_int64, _dbl_p]
    L.orc_score_test_1.restype = C.c_int
    L.orc_score_test_1.argtypes = [C.c_int64, C.c_int, _dbl_p, _dbl_p, C.c_double, _dbl_p, _dbl_p,
                                   _dbl_p, _dbl_p, _dbl_p]
    L.orc_gene.restype = C.c_int
    L.orc_gene.argtypes = [C.c_int64, C.c_int, C.c_int, _dbl_p, _dbl_p, _dbl_p, _dbl_p, C.c_double,
                           C.c_double, C.c_double, C.POINTER(GeneOut), _dbl_p]
    L.orc_gene_batch.restype = C.c_int
    L.orc_gene_batch.argtypes = [C.c_int64, C.c_int, C.c_int, C.c_int, _dbl_p, _dbl_p, _dbl_p,
                                 _dbl_p, C.c_double, C.c_double, C.c_double, C.POINTER(GeneOut),
                                 C.c_int]
    L.orc_gene_batch_idx.restype = C.c_int
    L.orc_gene_batch_idx.argtypes = [C.c_int64, C.c_int, C.c_int, C.c_int, _int_p, _dbl_p, _dbl_p, _dbl_p,
                                     _dbl_p, C.c_double, C.c_double, C.c_double, C.POINTER(GeneOut),
                                     C.c_int]
    L.orc_synth_rows_f64.restype = None
    L.orc_synth_rows_f64.argtypes = [C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, _dbl_p, C.c_int]
    L.orc_max_threads.restype = C.c_int
    _lib_cache[key] = L
    return L


def ref_mix():
    """The reference's own MixtureChiSquare (None when oracle/_ref was never built)."""
    if "mix" not in _lib_cache:
        path = os.path.join(_HERE, "_ref", "libmixchisq_ref.so")
        if not os.path.exists(path):
            _lib_cache["mix"] = None
        else:
            L = C.CDLL(path)
            L.ref_mixchisq_pvalue.restype = C.c_double
            L.ref_mixchisq_pvalue.argtypes = [_dbl_p, C.c_int, C.c_double]
            L.ref_liu_pvalue.restype = C.c_double
            L.ref_liu_pvalue.argtypes = [_dbl_p, C.c_int, C.c_double]
            L.ref_qf.restype = C.c_double
            L.ref_qf.argtypes = [_dbl_p, _dbl_p, _int_p, C.c_int, C.c_double, C.c_double, C.c_int,
                                 C.c_double, _dbl_p, _int_p]
            _lib_cache["mix"] = L
    return _lib_cache["mix"]


_QAGS_CB = C.CFUNCTYPE(C.c_double, C.c_double, C.c_void_p)


def ref_gsl():
    """GSL 1.16 as vendored by the reference (None when oracle/_ref was never built)."""
    if "gsl" not in _lib_cache:
        path = os.path.join(_HERE, "_ref", "libgsl_ref.so")
        if not os.path.exists(path):
            _lib_cache["gsl"] = None
        else:
            L = C.CDLL(path)
            for n, na in (("ref_gsl_ran_beta_pdf", 3), ("ref_gsl_cdf_chisq_Q", 2),
                          ("ref_gsl_cdf_chisq_P", 2), ("ref_gsl_cdf_chisq_Qinv", 2),
                          ("ref_gsl_ran_chisq_pdf", 2)):
                f = getattr(L, n)
                f.restype = C.c_double
                f.argtypes = [C.c_double] * na
            L.ref_gsl_qags.restype = C.c_int
            L.ref_gsl_qags.argtypes = [_QAGS_CB, C.c_void_p, C.c_double, C.c_double, C.c_double,
                                       C.c_double, C.c_int, _dbl_p, _dbl_p, _int_p]
            _lib_cache["gsl"] = L
    return _lib_cache["gsl"]


def ref_skat():
    """The reference's own Skat.cpp / SkatO.cpp / LinearRegression.cpp / LinearRegressionScoreTest.cpp,
    compiled unmodified against oracle/eigen_standin + the vendored GSL (oracle/Makefile,
    oracle/ref_skat_shim.cpp).  None when oracle/_ref was never built."""
    if "skat" not in _lib_cache:
        path = os.path.join(_HERE, "_ref", "libskat_ref.so")
        if not os.path.exists(path):
            _lib_cache["skat"] = None
        else:
            L = C.CDLL(path)
            L.ref_skat_fit.restype = C.c_int
            L.ref_skat_fit.argtypes = [C.c_int, C.c_int, C.c_int, _dbl_p, _dbl_p, _dbl_p, _dbl_p, _dbl_p,
                                       _dbl_p, _dbl_p, C.c_int, _dbl_p, _dbl_p]
            L.ref_skato_fit.restype = C.c_int
            L.ref_skato_fit.argtypes = [C.c_int, C.c_int, C.c_int, _dbl_p, _dbl_p, _dbl_p, _dbl_p, _dbl_p,
                                        C.c_char_p, _dbl_p, _dbl_p, _dbl_p]
            L.ref_linear_fit.restype = C.c_int
            L.ref_linear_fit.argtypes = [C.c_int, C.c_int, _dbl_p, _dbl_p, _dbl_p, _dbl_p, _dbl_p, _dbl_p, _dbl_p]
            L.ref_score_test.restype = C.c_int
            L.ref_score_test.argtypes = [C.c_int, C.c_int, C.c_int, _dbl_p, _dbl_p, _dbl_p, C.c_int,
                                         _dbl_p, _dbl_p, _dbl_p, _dbl_p, _dbl_p, _dbl_p, _dbl_p]
            L.ref_fastlmm_score.restype = C.c_int
            L.ref_fastlmm_score.argtypes = [C.c_int, C.c_int, C.c_int, _dbl_p, _dbl_p, C.c_void_p, C.c_void_p, _dbl_p,
                                            _dbl_p, _dbl_p, _dbl_p, _dbl_p, _dbl_p, _dbl_p]
            L.ref_genotype_counter.restype = None
            L.ref_genotype_counter.argtypes = [C.c_int, _dbl_p, _dbl_p]
            L.ref_logistic_fit.restype = C.c_int
            L.ref_logistic_fit.argtypes = [C.c_int, C.c_int, _dbl_p, _dbl_p, C.c_int, _dbl_p, _dbl_p, _dbl_p, _dbl_p]
            L.ref_logistic_score_test.restype = C.c_int
            L.ref_logistic_score_test.argtypes = [C.c_int, C.c_int, _dbl_p, _dbl_p, _dbl_p, _dbl_p, _dbl_p, _dbl_p, _dbl_p]
            L.ref_skat_perm.restype = C.c_int
            L.ref_skat_perm.argtypes = [C.c_int, C.c_int, C.c_int, _dbl_p, _dbl_p, _dbl_p, _dbl_p, _dbl_p, C.c_int,
                                        C.c_double, C.c_uint, _dbl_p, _int_p, _int_p, _int_p, _dbl_p, _dbl_p]
            _lib_cache["skat"] = L
    return _lib_cache["skat"]


def ref_skat_perm(res, v, X, G, w, n_perm=10000, alpha=0.05, reseed=1):
    """The permutation loop of SkatTest::fit (src/Model.h:2707-2717) run by the reference build: its own permute(),
    Permutation and Skat::GetQFromNewResidual on the process-wide glibc rand() stream (reseed=1: fresh process)."""
    Gc = np.asfortranarray(G, dtype=np.float64)
    Xc = np.asfortranarray(X, dtype=np.float64)
    N, M = Gc.shape
    res = np.ascontiguousarray(res, dtype=np.float64)
    v = np.ascontiguousarray(v, dtype=np.float64)
    w = np.ascontiguousarray(w, dtype=np.float64)
    stat, p = C.c_double(0), C.c_double(0)
    a, g, e = C.c_int(0), C.c_int(0), C.c_int(0)
    q = np.zeros(max(n_perm, 1))
    rc = ref_skat().ref_skat_perm(N, M, Xc.shape[1], _p(res), _p(v), _p(Xc), _p(Gc), _p(w), int(n_perm), float(alpha),
                                  int(reseed), C.byref(stat), C.byref(a), C.byref(g), C.byref(e), C.byref(p), _p(q))
    return dict(rc=rc, stat=stat.value, actual=a.value, greater=g.value, equal=e.value, p=p.value, q=q[: a.value].copy())


def ref_skat_fit(res, v, X, G, w, res_perm=None):
    """Skat::Fit of the reference build on (N,) res, (N,) v, (N,C) X, (N,M) G (already flipped to the minor
    allele, polymorphic columns only), (M,) w (squared Beta densities, src/Model.h:2644-2661).
    Returns dict(rc, Q, pvalue[, q_perm])."""
    Gc = np.asfortranarray(G, dtype=np.float64)
    Xc = np.asfortranarray(X, dtype=np.float64)
    N, M = Gc.shape
    res = np.ascontiguousarray(res, dtype=np.float64)
    v = np.ascontiguousarray(v, dtype=np.float64)
    w = np.ascontiguousarray(w, dtype=np.float64)
    Q, p = C.c_double(0), C.c_double(0)
    n_perm, rp, qp = 0, None, None
    if res_perm is not None:
        rp = np.ascontiguousarray(res_perm, dtype=np.float64)
        n_perm = rp.shape[0]
        qp = np.zeros(n_perm)
    rc = ref_skat().ref_skat_fit(N, M, Xc.shape[1], _p(res), _p(v), _p(Xc), _p(Gc), _p(w), C.byref(Q), C.byref(p),
                                 n_perm, _p(rp) if n_perm else None, _p(qp) if n_perm else None)
    out = dict(rc=rc, Q=Q.value, pvalue=p.value)
    if n_perm:
        out["q_perm"] = qp
    return out


def ref_skato_fit(res, v, X, G, w, binary=False):
    """SkatO::Fit of the reference build; w are the UN-squared Beta densities (src/Model.h:2799-2813)."""
    Gc = np.asfortranarray(G, dtype=np.float64)
    Xc = np.asfortranarray(X, dtype=np.float64)
    N, M = Gc.shape
    res = np.ascontiguousarray(res, dtype=np.float64)
    v = np.ascontiguousarray(v, dtype=np.float64)
    w = np.ascontiguousarray(w, dtype=np.float64)
    Q, rho, p = C.c_double(0), C.c_double(0), C.c_double(0)
    rc = ref_skat().ref_skato_fit(N, M, Xc.shape[1], _p(res), _p(v), _p(Xc), _p(Gc), _p(w), b"D" if binary else b"C",
                                  C.byref(Q), C.byref(rho), C.byref(p))
    return dict(rc=rc, Q=Q.value, rho=rho.value, pvalue=p.value)


def ref_linear_fit(X, y):
    """LinearRegression::FitLinearModel of the reference build."""
    Xc = np.asfortranarray(X, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    N, Cc = Xc.shape
    beta, resid, pred, covB = np.zeros(Cc), np.zeros(N), np.zeros(N), np.zeros((Cc, Cc), order="F")
    s2 = C.c_double(0)
    rc = ref_skat().ref_linear_fit(N, Cc, _p(Xc), _p(y), _p(beta), _p(resid), _p(pred), C.byref(s2), _p(covB))
    return dict(rc=rc, beta=beta, resid=resid, predicted=pred, sigma2=s2.value, covB=covB)


def ref_score_test(Xnull, y, Xcol, force_matrix=False):
    """LinearRegressionScoreTest::FitNullModel + TestCovariate of the reference build; Xcol (N,) or (N, M)."""
    Xc = np.asfortranarray(Xnull, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    g = np.asfortranarray(np.asarray(Xcol, dtype=np.float64).reshape(len(y), -1))
    N, Cc = Xc.shape
    M = g.shape[1]
    U, V, beta = np.zeros(M), np.zeros((M, M), order="F"), np.zeros(M)
    stat, p, s2, se = C.c_double(0), C.c_double(0), C.c_double(0), C.c_double(0)
    rc = ref_skat().ref_score_test(N, Cc, M, _p(Xc), _p(y), _p(g), int(force_matrix), _p(U), _p(V), _p(beta),
                                   C.byref(stat), C.byref(p), C.byref(s2), C.byref(se))
    return dict(rc=rc, U=U, V=V, beta=beta, stat=stat.value, pvalue=p.value, sigma2=s2.value, se_beta=se.value)


def ref_fastlmm_score(X, y, U, S, G):
    """FastLMM(SCORE, MLE)::FitNullModel + TestCovariate per column of G (N, M) of the reference build;
    U (N, N) float32 kinship eigenvectors, S (N,) float32 eigenvalues."""
    Xc = np.asfortranarray(X, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    Uf = np.asfortranarray(U, dtype=np.float32)
    Sf = np.ascontiguousarray(S, dtype=np.float32)
    Gc = np.asfortranarray(G, dtype=np.float64)
    N, Cc = Xc.shape
    M = Gc.shape[1]
    delta, s2 = C.c_double(0), C.c_double(0)
    beta, Us, Vs, ps = np.zeros(Cc), np.zeros(M), np.zeros(M), np.zeros(M)
    rc = ref_skat().ref_fastlmm_score(N, Cc, M, _p(Xc), _p(y), Uf.ctypes.data, Sf.ctypes.data, _p(Gc), C.byref(delta),
                                      C.byref(s2), _p(beta), _p(Us), _p(Vs), _p(ps))
    return dict(rc=rc, delta=delta.value, sigma2=s2.value, beta=beta, U=Us, V=Vs, pvalue=ps)


def ref_logistic_fit(X, y, rounds=100):
    """LogisticRegression::FitLogisticModel of the reference build -> dict(rc, beta, p, v, covB)."""
    Xc = np.asfortranarray(X, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    N, Cc = Xc.shape
    beta, p, v, covB = np.zeros(Cc), np.zeros(N), np.zeros(N), np.zeros((Cc, Cc), order="F")
    rc = ref_skat().ref_logistic_fit(N, Cc, _p(Xc), _p(y), int(rounds), _p(beta), _p(p), _p(v), _p(covB))
    return dict(rc=rc, beta=beta, p=p, v=v, covB=covB)


def ref_logistic_score_test(Xnull, y, xcol):
    """LogisticRegressionScoreTest::FitNullModel + TestCovariate (Matrix overload) of the reference build;
    intercept-only null models only (rc = -3 otherwise: the reference indexes out of bounds there)."""
    Xc = np.asfortranarray(Xnull, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    g = np.ascontiguousarray(xcol, dtype=np.float64)
    U, V, stat, p = C.c_double(0), C.c_double(0), C.c_double(0), C.c_double(0)
    rc = ref_skat().ref_logistic_score_test(len(y), Xc.shape[1], _p(Xc), _p(y), _p(g), C.byref(U), C.byref(V),
                                            C.byref(stat), C.byref(p))
    return dict(rc=rc, U=U.value, V=V.value, stat=stat.value, pvalue=p.value)


def ref_genotype_counter(g):
    """GenotypeCounter of the reference build on one variant -> dict of the MetaScore site columns."""
    g = np.ascontiguousarray(g, dtype=np.float64)
    out = np.zeros(8)
    ref_skat().ref_genotype_counter(len(g), _p(g), _p(out))
    return dict(n_ref=int(out[0]), n_het=int(out[1]), n_alt=int(out[2]), n_missing=int(out[3]), call_rate=out[4],
                af=out[5], ac=out[6], hwe_p=out[7])


def ref_vcf():
    """The reference's own VCF record parser (libVcf/VCFRecord, VCFIndividual, VCFValue ... compiled unmodified into
    oracle/_ref/libvcf_ref.so behind oracle/ref_vcf_shim.cpp).  None when oracle/_ref was never built."""
    if "vcf" not in _lib_cache:
        path = os.path.join(_HERE, "_ref", "libvcf_ref.so")
        if not os.path.exists(path):
            _lib_cache["vcf"] = None
        else:
            L = C.CDLL(path)
            L.ref_vcf_genotypes.restype = C.c_int
            L.ref_vcf_genotypes.argtypes = [C.c_char_p, C.c_char_p, _int_p, C.c_int, C.c_char_p, _int_p]
            L.ref_vcf_gt.restype = C.c_int
            L.ref_vcf_gt.argtypes = [C.c_char_p, C.c_int]
            L.ref_vcf_genotypes_filtered.restype = C.c_int
            L.ref_vcf_genotypes_filtered.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, _int_p, C.c_int]
            L.ref_vcf_gt_male02.restype = C.c_int
            L.ref_vcf_gt_male02.argtypes = [C.c_char_p, C.c_int]
            L.ref_par_is_hemi.restype = C.c_int
            L.ref_par_is_hemi.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]
            L.ref_vcf_genotypes_sex.restype = C.c_int
            L.ref_vcf_genotypes_sex.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, _int_p, C.c_char_p, _dbl_p, C.c_int]
            L.ref_vcf_count_alt.restype = C.c_int
            L.ref_vcf_count_alt.argtypes = [C.c_char_p, C.c_int, C.c_int]
            L.ref_vcf_count_male_alt2.restype = C.c_int
            L.ref_vcf_count_male_alt2.argtypes = [C.c_char_p, C.c_int, C.c_int]
            L.ref_vcf_genotypes_alt.restype = C.c_int
            L.ref_vcf_genotypes_alt.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, _int_p, C.c_int, _int_p, C.c_int, _int_p]
            L.ref_parse_range.restype = C.c_int
            L.ref_parse_range.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(C.c_uint), C.POINTER(C.c_uint)]
            L.ref_vcf_dosages.restype = C.c_int
            L.ref_vcf_dosages.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, _dbl_p, C.c_int]
            _lib_cache["vcf"] = L
    return _lib_cache["vcf"]


def ref_vcf_genotypes(header: str, record: str):
    """One VCF data line through the reference's parser: (chrom, pos, genotypes per sample with -9 = missing)."""
    L = ref_vcf()
    cap = header.count("\t") + 1
    out = np.zeros(cap, dtype=np.int32)
    chrom = C.create_string_buffer(64)
    pos = C.c_int(0)
    n = L.ref_vcf_genotypes(header.encode(), record.encode(), out.ctypes.data_as(_int_p), cap, chrom, C.byref(pos))
    if n < 0:
        return None
    return chrom.value.decode(), pos.value, out[:n].copy()


def ref_vcf_genotypes_filtered(header: str, record: str, gd=(-1, -1), gq=(-1, -1)):
    """hard calls through the reference's parser with --indvDepthMin/Max, --indvQualMin/Max (oracle/ref_vcf_shim.cpp)"""
    L = ref_vcf()
    cap = header.count("\t") + 1
    out = np.zeros(cap, dtype=np.int32)
    n = L.ref_vcf_genotypes_filtered(header.encode(), record.encode(), gd[0], gd[1], gq[0], gq[1], out.ctypes.data_as(_int_p), cap)
    return None if n < 0 else out[:n].copy()


def ref_vcf_genotypes_sex(header: str, record: str, sex, x_label="", par_region="", dosage_tag=""):
    """one VCF line with the X handling of VCFGenotypeExtractor::getGenotype over the reference's VCFValue + ParRegion"""
    L = ref_vcf()
    cap = header.count("\t") + 1
    out = np.zeros(cap, dtype=np.float64)
    sx = np.ascontiguousarray(sex, dtype=np.int32)
    n = L.ref_vcf_genotypes_sex(header.encode(), record.encode(), x_label.encode(), par_region.encode(),
                                sx.ctypes.data_as(_int_p), dosage_tag.encode(), _p(out), cap)
    return None if n < 0 else out[:n].copy()


def ref_vcf_genotypes_alt(header: str, record: str, alt: int, sex=None, x_label="", par_region=""):
    """--multipleAllele: copies of alt allele `alt` per sample through the reference's VCFValue (+ ParRegion / sex);
    returns (genotypes, number of ALT alleles of the record)"""
    L = ref_vcf()
    cap = header.count("\t") + 1
    out = np.zeros(cap, dtype=np.int32)
    n_alt = C.c_int(0)
    sx = None if sex is None else np.ascontiguousarray(sex, dtype=np.int32)
    n = L.ref_vcf_genotypes_alt(header.encode(), record.encode(), x_label.encode(), par_region.encode(),
                                None if sx is None else sx.ctypes.data_as(_int_p), int(alt), out.ctypes.data_as(_int_p), cap,
                                C.byref(n_alt))
    return (None, 0) if n < 0 else (out[:n].copy(), n_alt.value)


def ref_vcf_dosages(header: str, record: str, tag: str):
    """One VCF data line through the reference's parser in dosage mode: toDouble() of the `tag` subfield per sample."""
    L = ref_vcf()
    cap = header.count("\t") + 1
    out = np.zeros(cap, dtype=np.float64)
    n = L.ref_vcf_dosages(header.encode(), record.encode(), tag.encode(), _p(out), cap)
    return None if n < 0 else out[:n].copy()


def vcf_record_dosages(header: str, record: str, tag: str):
    """dosage mode restated (src/VCFGenotypeExtractor.cpp:70-76, 404-406, 434-438): first FORMAT key that starts with `tag`;
    value = C atof of the subfield ("." / garbage / an absent subfield -> 0.0); no such key -> -9 for everyone."""
    import ctypes.util
    libc = C.CDLL(ctypes.util.find_library("c"))
    libc.atof.restype = C.c_double
    libc.atof.argtypes = [C.c_char_p]
    names = header.rstrip("\r\n").split("\t")[9:]
    f = record.rstrip("\r\n").split("\t")
    if len(f) < 10 or len(f) - 9 != len(names):
        return None
    idx = -1
    for k, key in enumerate(f[8].split(":")):
        if key.startswith(tag):
            idx = k
            break
    out = np.full(len(names), -9.0)
    if idx >= 0:
        for i, col in enumerate(f[9:]):
            sub = col.split(":")
            out[i] = libc.atof((sub[idx] if idx < len(sub) else "").encode())
    return out


def genotype_counter(g):
    """GenotypeCounter::add / getAF (src/GenotypeCounter.h:14-52) over one variant: (hom-ref, het, hom-alt, missing, AF)."""
    g = np.asarray(g, dtype=np.float64)
    miss = (g < 0) | (g > 2.0)
    ref_ = ~miss & (g < 2.0 / 3)
    het = ~miss & ~ref_ & (g < 4.0 / 3)
    alt = ~miss & ~ref_ & ~het
    s = 0.0
    for x in g[~miss]:          # same summation order as the reference's running sumAC
        s += x
    return int(ref_.sum()), int(het.sum()), int(alt.sum()), int(miss.sum()), (0.5 * s / len(g) if len(g) else -1.0)


def impute_mean_literal(G):
    """DataConsolidator::imputeGenotypeToMean with its INTEGER accumulator (`int ac; ac += m(j, i)` truncates after every
    addition, src/DataConsolidator.cpp:223-229): differs from impute_mean only for non-integer dosages."""
    G = np.array(G, dtype=np.float64)
    for i in range(G.shape[1]):
        col = G[:, i]
        if not (col < 0).any():
            continue
        ac, an = 0, 0
        for x in col:
            if x >= 0:
                ac = int(ac + x)
                an += 2
        col[col < 0] = 2.0 * (0.0 if an == 0 else 1.0 * ac / an)
    return G


def vcf_gt(s: str) -> int:
    """VCFValue::getGenotype (libVcf/VCFValue.h:74-116) restated: the GT grammar of the default --inVcf path.
    '0' / '1' haploid; a|b or a/b with single-digit alleles 0/1; '.' anywhere, any allele > 1 (multi-allelic), a wrong
    separator or trailing characters -> missing (-9); a second allele below '0' (e.g. '-') is reported and IGNORED."""
    c = s[0] if len(s) > 0 else "\0"
    if c == "." or c < "0":
        return -9
    g = ord(c) - ord("0")
    if g > 1:
        return -9
    if len(s) == 1:
        return g
    if s[1] not in "|/":
        return -9
    if len(s) == 2:
        return -9
    c = s[2]
    if c == ".":
        return -9
    if not c < "0":
        a2 = ord(c) - ord("0")
        if a2 > 1:
            return -9
        g += a2
    if len(s) != 3:
        return -9
    return g


def vcf_record_genotypes(header: str, record: str):
    """VCFRecord::parse + getFormatIndex("GT") (prefix match) + VCFIndividual::parse + justGet + getGenotype restated:
    (chrom, pos, genotypes) or None for a record whose sample count differs from the header's."""
    names = header.rstrip("\r\n").split("\t")[9:]
    f = record.rstrip("\r\n").split("\t")
    if len(f) < 10 or len(f) - 9 != len(names):
        return None
    idx = -1
    for k, key in enumerate(f[8].split(":")):
        if key.startswith("GT"):
            idx = k
            break
    out = np.full(len(names), -9, dtype=np.int32)
    if idx >= 0:
        for i, col in enumerate(f[9:]):
            sub = col.split(":")
            out[i] = vcf_gt(sub[idx] if idx < len(sub) else "")
    return f[0], int(f[1]) if f[1].lstrip("+-").isdigit() else 0, out


def ref_model():
    """The reference's own model layer (src/Model.cpp + Model.h fitters, src/DataConsolidator.cpp ...) built into
    oracle/_ref/libmodel_ref.so behind oracle/ref_model_shim.cpp.  None when oracle/_ref was never built."""
    if "model" not in _lib_cache:
        path = os.path.join(_HERE, "_ref", "libmodel_ref.so")
        if not os.path.exists(path):
            _lib_cache["model"] = None
        else:
            L = C.CDLL(path)
            L.ref_run_gene_models.restype = C.c_int
            L.ref_run_gene_models.argtypes = [C.c_int, C.c_int, _int_p, _dbl_p, C.c_int, _dbl_p, _dbl_p, C.c_int,
                                              C.c_double, C.c_int, C.c_char_p]
            L.ref_run_meta_models.restype = C.c_int
            L.ref_run_meta_models.argtypes = [C.c_int, C.c_int, _dbl_p, _int_p, C.c_int, _dbl_p, _dbl_p, C.c_int,
                                              C.c_char_p]
            _lib_cache["model"] = L
    return _lib_cache["model"]


def _read_assoc(path):
    """-> (comment lines, header fields, rows as lists of strings)"""
    comments, rows, header = [], [], None
    with open(path) as f:
        for line in f:
            line = line.rstrip("\n")
            if line.startswith("#"):
                comments.append(line)
            elif header is None:
                header = line.split("\t")
            else:
                rows.append(line.split("\t"))
    return comments, header, rows


def ref_run_gene_models(genes, cov, pheno, prefix, n_perm=0, alpha=0.05, binary=False):
    """Gene loop of src/Main.cpp:1221-1254 run by the reference build on `genes` (list of (N, M_k) raw genotype
    matrices, missing < 0), cov (N, C-1) WITHOUT the intercept, pheno (N,).  Returns {model name: (comments, header, rows)}
    parsed from the `.assoc` files the reference wrote under `prefix`."""
    N = len(pheno)
    flat = np.concatenate([np.asfortranarray(g, dtype=np.float64).ravel(order="F") for g in genes])
    M = np.array([g.shape[1] for g in genes], dtype=np.int32)
    covf = np.asfortranarray(np.asarray(cov, dtype=np.float64).reshape(N, -1))
    y = np.ascontiguousarray(pheno, dtype=np.float64)
    rc = ref_model().ref_run_gene_models(N, len(genes), _p(M, C.c_int), _p(flat), covf.shape[1], _p(covf), _p(y),
                                         int(n_perm), float(alpha), int(binary), prefix.encode())
    if rc:
        raise RuntimeError(f"ref_run_gene_models rc={rc}")
    return {m: _read_assoc(f"{prefix}.{m}.assoc") for m in ("Skat", "SkatO", "CMC", "Zeggini")}


def ref_dropin():
    """oracle/_ref/libdropin_ref.so: the reference's own ModelManager (registration patch of rvtests_b200/host/ModelB200.h
    applied) + Main.cpp's gene loop (oracle/ref_dropin_shim.cpp), linked with the reference model layer AND the product's
    C ABI.  None when oracle/_ref was never built."""
    if "dropin" not in _lib_cache:
        path = os.path.join(_HERE, "_ref", "libdropin_ref.so")
        if not os.path.exists(path):
            _lib_cache["dropin"] = None
        else:
            # its DT_NEEDED entries (libmodel_ref.so, librvtests_b200.so) resolve through $ORIGIN rpaths
            L = C.CDLL(path, mode=C.RTLD_GLOBAL)
            L.dropin_run_gene_models.restype = C.c_int
            L.dropin_run_gene_models.argtypes = [C.c_int, C.c_int, _int_p, _dbl_p, C.c_int, _dbl_p, _dbl_p, C.c_char_p,
                                                 C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_char_p]
            L.dropin_run_meta_models.restype = C.c_int
            L.dropin_run_meta_models.argtypes = [C.c_int, C.c_int, _dbl_p, _int_p, C.c_int, _dbl_p, _dbl_p, C.c_char_p, C.c_int,
                                                 C.c_int, C.c_char_p, C.c_int]
            _lib_cache["dropin"] = L
    return _lib_cache["dropin"]


def dropin_run_meta_models(G, pos, cov, pheno, window, prefix, use_b200, se=False, segment=0, binary=False):
    """`--meta score[se],cov[windowSize=window]` through the reference's ModelManager and single-variant loop (see
    dropin_run_gene_models); the files carry ModelManager's `.assoc.gz` names but are plain text in this build."""
    Gc = np.asfortranarray(G, dtype=np.float64)
    N, nv = Gc.shape
    pos = np.ascontiguousarray(pos, dtype=np.int32)
    covf = np.asfortranarray(np.asarray(cov, dtype=np.float64).reshape(N, -1))
    y = np.ascontiguousarray(pheno, dtype=np.float64)
    spec = "score%s,cov[windowSize=%d]" % ("[se]" if se else "", int(window))
    rc = ref_dropin().dropin_run_meta_models(N, nv, _p(Gc), _p(pos, C.c_int), covf.shape[1], _p(covf), _p(y), spec.encode(),
                                             int(use_b200), int(segment), prefix.encode(), int(binary))
    if rc:
        raise RuntimeError(f"dropin_run_meta_models rc={rc}")
    return {m: _read_assoc(f"{prefix}.{m}.assoc.gz") for m in ("MetaScore", "MetaCov")}


def dropin_run_gene_models(genes, cov, pheno, prefix, use_b200, kernel="skat[nPerm=0],skato", burden="cmc,zeggini",
                           binary=False, batch=0):
    """`--kernel <kernel> --burden <burden>` through the reference's ModelManager and gene loop; use_b200 selects the
    reference's own fitters (False) or the B200 adapters registered behind the same names (True)."""
    N = len(pheno)
    flat = np.concatenate([np.asfortranarray(g, dtype=np.float64).ravel(order="F") for g in genes])
    M = np.array([g.shape[1] for g in genes], dtype=np.int32)
    covf = np.asfortranarray(np.asarray(cov, dtype=np.float64).reshape(N, -1))
    y = np.ascontiguousarray(pheno, dtype=np.float64)
    rc = ref_dropin().dropin_run_gene_models(N, len(genes), _p(M, C.c_int), _p(flat), covf.shape[1], _p(covf), _p(y),
                                             kernel.encode(), burden.encode(), int(binary), int(use_b200), int(batch),
                                             prefix.encode())
    if rc:
        raise RuntimeError(f"dropin_run_gene_models rc={rc}")
    return {m: _read_assoc(f"{prefix}.{m}.assoc") for m in ("Skat", "SkatO", "CMC", "Zeggini")}


def ref_run_meta_models(G, pos, cov, pheno, window, prefix):
    """Single-variant loop of src/Main.cpp:1092-1147 with MetaScoreTest + MetaCovTest(window) run by the reference build;
    variant j = column j of G (N, n_var) at 1:pos[j]."""
    Gc = np.asfortranarray(G, dtype=np.float64)
    N, nv = Gc.shape
    pos = np.ascontiguousarray(pos, dtype=np.int32)
    covf = np.asfortranarray(np.asarray(cov, dtype=np.float64).reshape(N, -1))
    y = np.ascontiguousarray(pheno, dtype=np.float64)
    rc = ref_model().ref_run_meta_models(N, nv, _p(Gc), _p(pos, C.c_int), covf.shape[1], _p(covf), _p(y), int(window),
                                         prefix.encode())
    if rc:
        raise RuntimeError(f"ref_run_meta_models rc={rc}")
    return {m: _read_assoc(f"{prefix}.{m}.assoc") for m in ("MetaScore", "MetaCov")}


# ------------------------------------------------------------------------------------------------
# thin numpy-level helpers
# ------------------------------------------------------------------------------------------------
def qf(lam, Q, lim=10000, acc=1e-6, which="oracle"):
    """Davies qf() -> (value, ifault, trace[7]).  which = 'oracle' | 'reference'."""
    lam = np.ascontiguousarray(lam, dtype=np.float64)
    n = len(lam)
    nc = np.zeros(n)
    df = np.ones(n, dtype=np.int32)
    trace = np.zeros(7)
    fault = C.c_int(0)
    if which == "oracle":
        v = lib().orc_qf(_p(lam), _p(nc), _p(df, C.c_int), n, 0.0, float(Q), int(lim), float(acc),
                         _p(trace), C.byref(fault))
    else:
        v = ref_mix().ref_qf(_p(lam), _p(nc), _p(df, C.c_int), n, 0.0, float(Q), int(lim),
                             float(acc), _p(trace), C.byref(fault))
    return v, fault.value, trace


def mix_pvalue(lam, Q, which="oracle"):
    lam = np.ascontiguousarray(lam, dtype=np.float64)
    if which == "oracle":
        fault = C.c_int(0)
        p = lib().orc_mixchisq_pvalue(_p(lam), len(lam), float(Q), C.byref(fault))
        return p, fault.value
    return ref_mix().ref_mixchisq_pvalue(_p(lam), len(lam), float(Q)), None


def liu_pvalue(lam, Q, which="oracle"):
    lam = np.ascontiguousarray(lam, dtype=np.float64)
    if which == "oracle":
        return lib().orc_liu_pvalue(_p(lam), len(lam), float(Q))
    return ref_mix().ref_liu_pvalue(_p(lam), len(lam), float(Q))


def skat_final_pvalue(lam, Q, which="oracle"):
    """Skat.cpp:100-103: Davies, then Liu when p<=0 or p==1."""
    p, fault = mix_pvalue(lam, Q, which)
    if p <= 0.0 or p == 1.0:
        p = liu_pvalue(lam, Q, which)
    return p, fault


def sym_eigenvalues(a):
    a = np.array(a, dtype=np.float64, order="C")
    n = a.shape[0]
    ev = np.zeros(n)
    lib().orc_sym_eigenvalues(n, _p(a), _p(ev))
    return ev


def fit_null_linear(X, y):
    """X: (N, C) any layout; returns dict(resid, sigma2, xtx_inv, beta)."""
    Xc = np.asfortranarray(X, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    N, Cc = Xc.shape
    resid = np.zeros(N)
    s2 = C.c_double(0)
    xi = np.zeros((Cc, Cc))
    beta = np.zeros(Cc)
    rc = lib().orc_fit_null_linear(N, Cc, _p(Xc), _p(y), _p(resid), C.byref(s2), _p(xi), _p(beta))
    if rc:
        raise RuntimeError("null model: X'X not positive definite")
    return dict(resid=resid, sigma2=s2.value, xtx_inv=xi, beta=beta)


def gene(G_raw, af, X, resid, sigma2, beta1=1.0, beta2=25.0, native=False):
    """Whole-gene oracle.  G_raw (N, M) doubles (unflipped, imputed), af per ORIGINAL column."""
    Gc = np.asfortranarray(G_raw, dtype=np.float64)
    Xc = np.asfortranarray(X, dtype=np.float64)
    N, M = Gc.shape
    out = GeneOut()
    lam = np.zeros(max(M, 1))
    af = np.ascontiguousarray(af, dtype=np.float64)
    resid = np.ascontiguousarray(resid, dtype=np.float64)
    lib(native).orc_gene(N, M, Xc.shape[1], _p(Gc), _p(af), _p(Xc), _p(resid), float(sigma2),
                         float(beta1), float(beta2), C.byref(out), _p(lam))
    return out, lam[: out.skat.n_lambda].copy()



def gene_perm(G_raw, af, resid, obs, n_perm=10000, alpha=0.05, reseed=0, beta1=1.0, beta2=25.0, native=False):
    """A6: the permutation loop of SkatTest::fit on one gene; consumes the process-wide glibc rand() stream
    (reseed=1 restarts it as in a fresh process).  Returns dict(actual, greater, equal, p, q)."""
    Gc = np.asfortranarray(G_raw, dtype=np.float64)
    N, M = Gc.shape
    af = np.ascontiguousarray(af, dtype=np.float64)
    resid = np.ascontiguousarray(resid, dtype=np.float64)
    out = (C.c_int * 3)()
    p = C.c_double(1.0)
    q = np.zeros(max(n_perm, 1))
    L = lib(native)
    L.orc_gene_perm.argtypes = [C.c_int64, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                C.c_double, C.c_double, C.c_double, C.c_int, C.c_double, C.c_uint,
                                C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.orc_gene_perm.restype = C.c_int
    rc = L.orc_gene_perm(N, M, _p(Gc), _p(af), _p(resid), float(obs), float(beta1), float(beta2), int(n_perm),
                         float(alpha), int(reseed), out, C.byref(p), _p(q))
    return dict(rc=rc, actual=out[0], greater=out[1], equal=out[2], p=p.value, q=q[: out[0]].copy())


def glibc_rand(n, reseed=0, skip=0, native=False):
    """the next n values of glibc rand() (reseed=1: as in a fresh process)"""
    out = np.zeros(n, dtype=np.int32)
    L = lib(native)
    L.orc_glibc_rand.argtypes = [C.c_uint, C.c_int64, C.c_int64, C.c_void_p]
    L.orc_glibc_rand.restype = None
    L.orc_glibc_rand(int(reseed), int(skip), int(n), out.ctypes.data)
    return out

def gene_batch(G_all, af_all, X, resid, sigma2, threads=1, beta1=1.0, beta2=25.0, native=False):
    """G_all: (n_genes, M, N) C-contiguous == each gene N x M column-major."""
    G_all = np.ascontiguousarray(G_all, dtype=np.float64)
    ng, M, N = G_all.shape
    Xc = np.asfortranarray(X, dtype=np.float64)
    out = (GeneOut * ng)()
    af_all = np.ascontiguousarray(af_all, dtype=np.float64)
    resid = np.ascontiguousarray(resid, dtype=np.float64)
    lib(native).orc_gene_batch(N, M, Xc.shape[1], ng, _p(G_all), _p(af_all), _p(Xc), _p(resid),
                               float(sigma2), float(beta1), float(beta2), out, int(threads))
    return out


def gene_batch_idx(G_all, af_all, index, X, resid, sigma2, threads=1, beta1=1.0, beta2=25.0, native=False):
    """tasks t -> gene index[t] of G_all (n_distinct, M, N); returns GeneOut array of len(index)."""
    G_all = np.ascontiguousarray(G_all, dtype=np.float64)
    nd, M, N = G_all.shape
    index = np.ascontiguousarray(index, dtype=np.int32)
    Xc = np.asfortranarray(X, dtype=np.float64)
    out = (GeneOut * len(index))()
    af_all = np.ascontiguousarray(af_all, dtype=np.float64)
    resid = np.ascontiguousarray(resid, dtype=np.float64)
    lib(native).orc_gene_batch_idx(N, M, Xc.shape[1], len(index), _p(index, C.c_int), _p(G_all), _p(af_all),
                                   _p(Xc), _p(resid), float(sigma2), float(beta1), float(beta2), out,
                                   int(threads))
    return out


def synth_rows_f64(keys, t0, t1, N, threads=0, native=False):
    """C twin of the device generator: (len(keys), N) doubles."""
    keys = np.ascontiguousarray(keys, dtype=np.uint64)
    t0 = np.ascontiguousarray(t0, dtype=np.uint32)
    t1 = np.ascontiguousarray(t1, dtype=np.uint32)
    out = np.empty((len(keys), N), dtype=np.float64)
    lib(native).orc_synth_rows_f64(len(keys), N, keys.ctypes.data, t0.ctypes.data, t1.ctypes.data, _p(out),
                                   int(threads))
    return out


# ------------------------------------------------------------------------------------------------
# Synthetic data stream -- the HOST twin of rvtests_b200/csrc/synth.cuh.
# Counter-based: genotype(variant v, sample i) depends only on (seed, v, i), so any subset can be
# regenerated anywhere.  mix64 is the splitmix64 finaliser.
# ------------------------------------------------------------------------------------------------
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _mix64(z):
    z = np.asarray(z, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z


def synth_variant_key(seed, vid):
    with np.errstate(over="ignore"):
        return _mix64(np.uint64(seed) + np.asarray(vid, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15))


def synth_maf(seed, vid, lo=1e-4, hi=0.05):
    """MAF_j ~ log-uniform[lo, hi] from the variant key (fp64 on the host; the device only ever
    sees the integer thresholds derived from it)."""
    k = _mix64(synth_variant_key(seed, vid) ^ np.uint64(0xA5A5A5A5A5A5A5A5))
    u = (k >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    return lo * (hi / lo) ** u


def synth_thresholds(maf):
    """uint32 thresholds t0,t1: g = (h>=t0) + (h>=t1) with h the 32-bit hash, HWE proportions."""
    maf = np.asarray(maf, dtype=np.float64)
    q0 = (1.0 - maf) ** 2
    q01 = q0 + 2.0 * maf * (1.0 - maf)
    t0 = np.minimum(np.floor(q0 * 4294967296.0), 4294967295.0).astype(np.uint64).astype(np.uint32)
    t1 = np.minimum(np.floor(q01 * 4294967296.0), 4294967295.0).astype(np.uint64).astype(np.uint32)
    return t0, t1


def synth_genotypes(seed, vid, N, maf=None):
    """(len(vid), N) int8 genotypes in {0,1,2}."""
    vid = np.atleast_1d(np.asarray(vid, dtype=np.uint64))
    if maf is None:
        maf = synth_maf(seed, vid)
    t0, t1 = synth_thresholds(maf)
    key = synth_variant_key(seed, vid)[:, None]
    i = np.arange(N, dtype=np.uint64)[None, :]
    with np.errstate(over="ignore"):
        h = (_mix64(key + i * np.uint64(0xD1B54A32D192ED03)) >> np.uint64(32)).astype(np.uint32)
    return ((h >= t0[:, None]).astype(np.int8) + (h >= t1[:, None]).astype(np.int8))


def synth_covariates(seed, N, C=3):
    """Intercept + (C-1) N(0,1) covariates and the null quantitative trait
    y = 0.5 x1 - 0.3 x2 + N(0,1)   (SURVEY.md 8(d))."""
    rng = np.random.Generator(np.random.Philox(key=int(seed)))
    X = np.ones((N, C))
    if C > 1:
        X[:, 1:] = rng.standard_normal((N, C - 1))
    y = rng.standard_normal(N)
    if C > 1:
        y = y + 0.5 * X[:, 1]
    if C > 2:
        y = y - 0.3 * X[:, 2]
    return X, y


# ---- PLINK 2-bit rows and mean imputation (checker for rvt_gene_push_bed) --------------------------
def bed_decode(bed, N):
    """PlinkInputFile::readIntoMatrix, SNP-major branch (libVcf/PlinkInputFile.cpp:23-47 with the
    codes of PlinkInputFile.h:206-209): geno = (byte >> 2*(p&3)) & 3; 0 -> 0, 2 -> 1, 3 -> 2,
    1 -> -9 (missing).  bed: (M, >= ceil(N/4)) uint8 -> (M, N) float64."""
    bed = np.asarray(bed, dtype=np.uint8)
    M = bed.shape[0]
    out = np.empty((M, N), dtype=np.float64)
    table = {0: 0.0, 2: 1.0, 3: 2.0, 1: -9.0}
    for m in range(M):
        for p in range(N):
            out[m, p] = table[(int(bed[m, p >> 2]) >> ((p & 3) << 1)) & 3]
    return out


def bed_decode_fast(bed, N):
    """vectorised twin of bed_decode (same table), for the larger test cases"""
    bed = np.asarray(bed, dtype=np.uint8)
    sh = (np.arange(N) & 3) << 1
    code = (bed[:, np.arange(N) >> 2] >> sh.astype(np.uint8)) & 3
    return np.array([0.0, -9.0, 1.0, 2.0])[code]


def impute_mean(G):
    """DataConsolidator::imputeGenotypeToMean (src/DataConsolidator.cpp:217-245): per column with a
    missing (<0) entry, p = ac / an over the called entries (0 if none), fill 2p.  G: (N, M)."""
    G = np.array(G, dtype=np.float64)
    for i in range(G.shape[1]):
        col = G[:, i]
        called = col >= 0
        if called.all():
            continue
        ac = int(col[called].sum())
        an = 2 * int(called.sum())
        p = 0.0 if an == 0 else 1.0 * ac / an
        col[~called] = 2.0 * p
    return G


# ------------------------------------------------------------------ the reference's own BoltLMM (oracle/_ref/libbolt_ref.so)
_ref_bolt = [False]


def ref_bolt():
    """The reference's own regression/BoltLMM.cpp + BoltPlinkLoader.cpp (+ PlinkInputFile, base/IO.cpp, cnpy), compiled
    unmodified against oracle/eigen_standin behind oracle/ref_bolt_shim.cpp.  None when oracle/_ref was never built."""
    if _ref_bolt[0] is False:
        path = os.path.join(_HERE, "_ref", "libbolt_ref.so")
        if not os.path.exists(path):
            _ref_bolt[0] = None
        else:
            L = C.CDLL(path)
            L.bolt_ref_fit.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_char_p, C.c_char_p, C.c_int]
            L.bolt_ref_test.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
            L.bolt_ref_covxx.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
            L.bolt_ref_free.restype = None
            _ref_bolt[0] = L
    return _ref_bolt[0]


def pack_plink(G):
    """(M, N) int8 hard calls with -1 = missing -> PLINK 2-bit SNP-major rows (00 hom-ref, 10 het, 11 hom-alt, 01 missing:
    libVcf/PlinkInputFile.h:14-20)."""
    G = np.asarray(G)
    code = np.where(G == 0, 0, np.where(G == 1, 2, np.where(G == 2, 3, 1))).astype(np.uint8)
    M, N = G.shape
    code = np.pad(code, ((0, 0), (0, (-N) % 4)))
    c4 = code.reshape(M, -1, 4)
    return (c4[:, :, 0] | (c4[:, :, 1] << 2) | (c4[:, :, 2] << 4) | (c4[:, :, 3] << 6)).astype(np.uint8)


def write_bolt_fileset(prefix, G, y, covar):
    """prefix.bed/.bim/.fam (+ .covar when there is more than the intercept) as `--boltPlink prefix` reads them
    (BoltPlinkLoader::open / loadCovariate, regression/BoltPlinkLoader.cpp:37-117)."""
    M, N = np.asarray(G).shape
    with open(prefix + ".bed", "wb") as f:
        f.write(bytes([0x6C, 0x1B, 1]) + pack_plink(G).tobytes())
    with open(prefix + ".bim", "w") as f:
        for j in range(M):
            f.write(f"1\trs{j}\t0\t{100 + j}\tA\tG\n")
    with open(prefix + ".fam", "w") as f:
        for i in range(N):
            f.write(f"F{i} S{i} 0 0 1 {y[i]:.9g}\n")
    covar = np.asarray(covar)
    if covar.shape[1] > 1:
        with open(prefix + ".covar", "w") as f:
            f.write("FID IID " + " ".join(f"c{k}" for k in range(1, covar.shape[1])) + "\n")
            for i in range(N):
                f.write(f"F{i} S{i} " + " ".join(f"{covar[i, k]:.9g}" for k in range(1, covar.shape[1])) + "\n")


def ref_bolt_fit(prefix, npz, log, pheno=None, binary=False):
    """BoltLMM::FitNullModel(prefix, phenotype) of the reference build; returns what the reference exports through
    BOLTLMM_SAVE_NULL_MODEL plus the secant path parsed from its BOLTLMM_DEBUG log."""
    import re
    L = ref_bolt()
    ph = None if pheno is None else np.ascontiguousarray(pheno, dtype=np.float64)
    rc = L.bolt_ref_fit(prefix.encode(), None if ph is None else ph.ctypes.data, 0 if ph is None else ph.size,
                        npz.encode(), log.encode(), int(binary))
    if rc != 0:
        raise RuntimeError(f"reference BoltLMM::FitNullModel returned {rc}")
    z = np.load(npz)
    text = open(log).read()
    ld, f = [], []
    for m in re.finditer(r"^i = (\d+)\tlogDelta = (\S+)\tf = (\S+)\tdelta = (\S+)\th2 = (\S+)$", text, re.M):
        ld.append(float(m.group(2)))
        f.append(float(m.group(3)))
    fin = re.search(r"^i = (\d+), delta = (\S+), sigma2_g = (\S+), sigma2_e = (\S+), h2 = (\S+)$", text, re.M)
    return dict(H_inv_y=np.array(z["H_inv_y"]).reshape(-1), H_inv_y_norm2=float(z["H_inv_y_norm2"][0]),
                infStatCalibration=float(z["infStatCalibration"][0]), xVx_xx_ratio=float(z["xVx_xx_ratio"][0]),
                log_delta=np.array(ld), f=np.array(f), final_i=int(fin.group(1)), delta=float(fin.group(2)),
                sigma2_g=float(fin.group(3)), sigma2_e=float(fin.group(4)), h2=float(fin.group(5)),
                n_solves=text.count("=> Enter solve()"))


def ref_bolt_test(g):
    """BoltLMM::TestCovariate on one variant: (af, U, V, effect, pvalue)"""
    g = np.ascontiguousarray(g, dtype=np.float64)
    out = np.zeros(5)
    ref_bolt().bolt_ref_test(g.ctypes.data, g.size, out.ctypes.data)
    return out


def ref_bolt_covxx(g1, g2):
    """BoltLMM::GetCovXX, both overloads: (vector<double> form, FloatMatrixRef form)"""
    a = np.ascontiguousarray(g1, dtype=np.float64)
    b = np.ascontiguousarray(g2, dtype=np.float64)
    out = np.zeros(2)
    ref_bolt().bolt_ref_covxx(a.ctypes.data, b.ctypes.data, a.size, out.ctypes.data)
    return out
