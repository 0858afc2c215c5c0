"""Regenerates tests/golden/ref_skat_golden.npz.  Run in the BUILD container (needs /root/reference):
    make -C oracle ref && python tests/golden/make_golden_ref_skat.py
The OUTPUTS stored here come from the REFERENCE's own sources -- regression/Skat.cpp, SkatO.cpp,
LinearRegression.cpp, LinearRegressionScoreTest.cpp compiled unmodified into oracle/_ref/libskat_ref.so
(oracle/Makefile; Eigen is replaced by oracle/eigen_standin, GSL is the vendored 1.16) -- called the way
SkatTest::fit / SkatOTest::fit / CMCTest::fit / ZegginiTest::fit call them (src/Model.h:2630-2720,
2780-2860, 820-870).  The INPUT preparation that lives in the reference's DataConsolidator / Model.cpp
(flip to the minor allele, drop monomorphic columns, the CMC / Zeggini collapse, the Beta weights) is
done here with the oracle's restatement of those steps: they are integer / elementwise operations that
the vectors store explicitly (G_flipped, weights, collapsed columns), so a reader can check them by eye."""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import oracle as O  # noqa: E402
from util import af_of, make_problem  # noqa: E402

O.build()
assert O.ref_skat() is not None and O.ref_gsl() is not None, "oracle/_ref not built"

# seed, N, M, C, n_mono, n_flip, maf_hi (None: the synthetic stream's own MAF spectrum)
CASES = [
    (110, 50, 1, 1, 0, 0, 0.2),
    (111, 97, 5, 1, 1, 1, 0.3),
    (116, 513, 7, 2, 7, 0, None),   # every variant monomorphic: the fitters return before Fit (Model.h:2636-2639)
    (112, 1000, 30, 3, 2, 3, None),
    (113, 1500, 64, 3, 0, 5, None),
    (114, 3001, 50, 3, 3, 4, None),
    (117, 2048, 33, 4, 1, 1, None),
    (141, 1500, 2, 3, 0, 1, 0.35),
    (142, 2000, 12, 3, 1, 2, 0.35),
    (143, 3000, 50, 3, 1, 2, 0.02),
    (145, 900, 7, 1, 1, 2, 0.35),
]


def prepare(G, af):
    """flip / drop (DataConsolidator.cpp:46-142), weights (Model.h:2644-2661, 2799-2813), collapse (Model.cpp:73-130)."""
    Gc = np.asfortranarray(G, dtype=np.float64)
    N, M = Gc.shape
    out = np.zeros((N, M), order="F")
    keep = np.zeros(M, dtype=np.int32)
    ip = C.POINTER(C.c_int)
    mp = O.lib().orc_flip_minor_polymorphic(N, M, O._p(Gc), O._p(out), keep.ctypes.data_as(ip), None)
    Gf = np.ascontiguousarray(out[:, :mp])
    gsl = O.ref_gsl()
    # the reference looks the frequency up by the index of the KEPT column in the caller-order table (F9)
    w1 = np.array([gsl.ref_gsl_ran_beta_pdf(min(a, 1 - a), 1.0, 25.0) if min(a, 1 - a) > 1e-30 else 0.0
                   for a in af[:mp]])
    cmc, zeg = np.zeros(N), np.zeros(N)
    Gff = np.asfortranarray(Gf)
    O.lib().orc_cmc_collapse(N, mp, O._p(Gff), O._p(cmc))
    O.lib().orc_zeggini_collapse(N, mp, O._p(Gff), O._p(zeg))
    return Gf, w1, cmc, zeg


store = dict(cases=np.array([c[:6] for c in CASES], dtype=np.int64), maf_hi=np.array([np.nan if c[6] is None else c[6] for c in CASES]))
for k, (seed, N, M, Cc, n_mono, n_flip, hi) in enumerate(CASES):
    maf = None if hi is None else (np.linspace(0.002, hi, M) if M > 1 else np.array([hi]))
    G, X, y = make_problem(O, seed, N, M, Cc, maf=maf, n_mono=n_mono, n_flip=n_flip)
    af = af_of(G)
    lin = O.ref_linear_fit(X, y)
    assert lin["rc"] == 0
    res, s2 = lin["resid"], lin["sigma2"]
    Gf, w1, cmc, zeg = prepare(G, af)
    v = np.full(N, s2)
    rng = np.random.default_rng(seed)
    perms = np.stack([rng.permutation(res) for _ in range(3)])
    if Gf.shape[1] == 0:  # genotype.cols == 0: no Fit call in the reference
        nan = float("nan")
        sk = dict(rc=-1, Q=nan, pvalue=nan, q_perm=np.full(3, nan))
        so = dict(rc=-1, Q=nan, rho=nan, pvalue=nan)
        sc = sz = dict(rc=-1, U=np.array([nan]), V=np.array([[nan]]), stat=nan, pvalue=nan)
    else:
        sk = O.ref_skat_fit(res, v, X, Gf, w1 * w1, res_perm=perms)
        so = O.ref_skato_fit(res, v, X, Gf, w1, binary=False)
        # CMCTest / ZegginiTest hold the collapsed genotype in a Matrix => the MATRIX overload of TestCovariate
        # (LinearRegressionScoreTest.cpp:173-263) is the one the reference runs (src/Model.h:855, :902)
        sc = O.ref_score_test(X, y, cmc, force_matrix=True)
        sz = O.ref_score_test(X, y, zeg, force_matrix=True)
    print(f"case {k} N={N} M={M} C={Cc}: m_poly={Gf.shape[1]} skat Q={sk['Q']:.8g} p={sk['pvalue']:.6g} | "
          f"skato rc={so['rc']} Q={so['Q']:.8g} rho={so['rho']} p={so['pvalue']:.6g} | "
          f"cmc rc={sc['rc']} p={sc['pvalue']:.6g} zeg rc={sz['rc']} p={sz['pvalue']:.6g}")
    store.update({
        f"G{k}": G.astype(np.int8), f"X{k}": X, f"y{k}": y, f"Gf{k}": Gf.astype(np.int8), f"w1_{k}": w1,
        f"cmc{k}": cmc.astype(np.int8), f"zeg{k}": zeg.astype(np.int16), f"perm{k}": perms,
        f"lin{k}": np.concatenate([[s2], lin["beta"], [np.abs(res).sum()], res[:8]]),
        f"skat{k}": np.array([sk["rc"], sk["Q"], sk["pvalue"], *sk["q_perm"]]),
        f"skato{k}": np.array([so["rc"], so["Q"], so["rho"], so["pvalue"]]),
        f"cmcst{k}": np.array([sc["rc"], sc["U"][0], sc["V"][0, 0], sc["stat"], sc["pvalue"]]),
        f"zegst{k}": np.array([sz["rc"], sz["U"][0], sz["V"][0, 0], sz["stat"], sz["pvalue"]]),
    })
np.savez_compressed(os.path.join(HERE, "ref_skat_golden.npz"), **store)
print("written", os.path.getsize(os.path.join(HERE, "ref_skat_golden.npz")), "bytes")
