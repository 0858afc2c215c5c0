"""GPU: wall time of the BoltLMM null-model fit (csrc/bolt.cuh) on a synthetic panel, beside a per-product estimate of what
the reference's float32 Eigen loop costs on the host (numpy sgemm on the same decoded batch sizes)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rvtests_b200  # noqa: E402

N, M, C = int(os.environ.get("BOLT_N", 100_000)), int(os.environ.get("BOLT_M", 10_000)), 3
rng = np.random.default_rng(7)
t = time.perf_counter()
maf = rng.uniform(0.05, 0.5, M)
bed = np.zeros((M, (N + 3) // 4), dtype=np.uint8)
beta = rng.normal(size=M) * np.sqrt(0.4 / M)
gv = np.zeros(N)
for m0 in range(0, M, 500):                                   # panel in slabs: 2-bit rows + the polygenic score
    G = rng.binomial(2, maf[m0:m0 + 500, None], size=(min(500, M - m0), N)).astype(np.uint8)
    code = np.where(G == 0, 0, np.where(G == 1, 2, 3)).astype(np.uint8)
    c4 = np.pad(code, ((0, 0), (0, (-N) % 4))).reshape(G.shape[0], -1, 4)
    bed[m0:m0 + 500] = c4[:, :, 0] | (c4[:, :, 1] << 2) | (c4[:, :, 2] << 4) | (c4[:, :, 3] << 6)
    p = maf[m0:m0 + 500, None]
    gv += ((G - 2 * p) / np.sqrt(2 * p * (1 - p))).T @ beta[m0:m0 + 500]
covar = np.column_stack([np.ones(N), rng.normal(size=N), rng.normal(size=N)])
y = gv + rng.normal(size=N) * np.sqrt(0.6) + covar @ np.array([1.0, 0.3, -0.2])
print(f"synthetic panel N={N} M={M} ({bed.nbytes / 1e6:.0f} MB as 2-bit rows) built in {time.perf_counter() - t:.1f} s", flush=True)
eng = rvtests_b200.GeneEngine(0)
t = time.perf_counter()
rec, h, Z = eng.bolt_fit_null(bed, N, y, covar)
dt = time.perf_counter() - t
hx = int(rec["cg_iterations"]) + 2 * int(rec["reml_evals"]) + 2
R = int(rec["mc_trials"]) + 1
print(f"bolt null fit: {dt:.2f} s  (h2 = {rec['h2']:.3f}, delta = {rec['delta']:.3f}, {rec['reml_evals']} REML evaluations, "
      f"{rec['cg_iterations']} CG iterations, MCtrial = {rec['mc_trials']}, calibration = {rec['inf_stat_calibration']:.4f})")
print(f"  ~{hx} H-products of 2 x N x M x R = {2 * N * M * R / 1e9:.1f} G multiply-adds each; panel bytes per product {2 * bed.nbytes / 1e6:.0f} MB")
# the reference's per-product cost on this host: decode 64-SNP batches to float and two sgemms per batch (computeHx)
nb = 40
Xb = rng.normal(size=(N, 64)).astype(np.float32)
V = rng.normal(size=(N, R)).astype(np.float32)
t = time.perf_counter()
for _ in range(nb):
    xy = Xb.T @ V
    out = Xb @ xy
dt_b = (time.perf_counter() - t) / nb
print(f"  host float32 sgemm pair per 64-SNP batch: {dt_b * 1e3:.2f} ms -> {dt_b * (M / 64):.2f} s per H-product (excluding the decode), "
      f"~{dt_b * (M / 64) * hx:.0f} s for the same fit on {os.cpu_count()} host threads")
