"""Generator of tests/golden/ref_bolt_golden.npz -- outputs of the REFERENCE's own BoltLMM (regression/BoltLMM.cpp +
BoltPlinkLoader.cpp compiled unmodified into oracle/_ref/libbolt_ref.so, oracle/ref_bolt_shim.cpp) on seeded PLINK
filesets: the null model as the reference itself exports it (BOLTLMM_SAVE_NULL_MODEL: H_inv_y, H_inv_y_norm2,
infStatCalibration, xVx_xx_ratio), the secant path of EstimateHeritabilityBolt from its BOLTLMM_DEBUG log, then
TestCovariate on 24 test variants and both GetCovXX overloads on 12 pairs.  Run in the build container (needs
/root/reference for the build):   python tests/golden/make_golden_ref_bolt.py
The inputs (2-bit panel, phenotype, covariates as the text files carried them) are stored next to the outputs, so the
CPU suite holds oracle/bolt_oracle.py and the GPU suite holds the device against the reference with nothing in between."""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import oracle as orc  # noqa: E402

CASES = [(101, 800, 300, 3, 0.5), (102, 1200, 500, 2, 0.2), (103, 400, 128, 1, 0.4)]


def panel(seed, N, M, miss=0.01):
    rng = np.random.default_rng(seed)
    maf = rng.uniform(0.05, 0.5, M)
    G = rng.binomial(2, maf[:, None], size=(M, N)).astype(np.int8)
    G[rng.random((M, N)) < miss] = -1
    G[3] = 0
    return G


def make_case(seed, N, M, C, h2):
    from oracle import bolt_oracle as BO
    G = panel(seed, N, M)
    rng = np.random.default_rng(seed + 1)
    covar = np.column_stack([np.ones(N)] + [rng.normal(size=N) for _ in range(C - 1)])
    X, _, _ = BO.prepare(G, covar, np.zeros(N))
    y = X @ rng.normal(size=M) * np.sqrt(h2 / M) + rng.normal(size=N) * np.sqrt(1 - h2) + covar @ rng.normal(size=C)
    # the values the text files carry
    y = np.array([float("%.9g" % v) for v in y])
    covar = np.array([[float("%.9g" % v) for v in row] for row in covar])
    Gt = rng.binomial(2, rng.uniform(0.05, 0.5, 24)[:, None], size=(24, N)).astype(np.int8)
    return G, y, covar, Gt


def run_reference(G, y, covar, Gt):
    N = G.shape[1]
    with tempfile.TemporaryDirectory() as d:
        prefix = os.path.join(d, "panel")
        orc.write_bolt_fileset(prefix, G, y, covar)
        out = orc.ref_bolt_fit(prefix, npz=os.path.join(d, "null.npz"), log=os.path.join(d, "log.txt"))
    tests = np.array([orc.ref_bolt_test(Gt[j].astype(np.float64)) for j in range(Gt.shape[0])])
    pairs = [(j, (j * 7 + 3) % Gt.shape[0]) for j in range(12)]
    cov = np.array([orc.ref_bolt_covxx(Gt[a].astype(np.float64), Gt[b].astype(np.float64)) for a, b in pairs])
    orc.ref_bolt().bolt_ref_free()
    out.update(tests=tests, pairs=np.array(pairs), covxx=cov)
    assert out["H_inv_y"].shape[0] >= N
    return out


def main():
    assert orc.ref_bolt() is not None, "build oracle/_ref/libbolt_ref.so first (make -C oracle ref)"
    blob = {}
    for k, case in enumerate(CASES):
        G, y, covar, Gt = make_case(*case)
        out = run_reference(G, y, covar, Gt)
        blob[f"c{k}_case"] = np.array(case, dtype=np.float64)
        blob[f"c{k}_bed"] = orc.pack_plink(G)
        blob[f"c{k}_y"] = y
        blob[f"c{k}_covar"] = covar
        blob[f"c{k}_gtest"] = Gt
        for key, v in out.items():
            blob[f"c{k}_{key}"] = np.asarray(v)
        print(case, "log_delta", out["log_delta"], "f", out["f"], "norm2", out["H_inv_y_norm2"], "calib",
              out["infStatCalibration"], "ratio", out["xVx_xx_ratio"])
    np.savez_compressed(os.path.join(HERE, "ref_bolt_golden.npz"), **blob)
    print("wrote", os.path.join(HERE, "ref_bolt_golden.npz"), os.path.getsize(os.path.join(HERE, "ref_bolt_golden.npz")), "bytes")


if __name__ == "__main__":
    main()
