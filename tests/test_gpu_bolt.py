"""GPU (-m gpu): A15, the BoltLMM null-model fit (regression/BoltLMM.cpp:169-299, 463-859, 1141-1214) against the numpy
restatement, which draws the reference's own random numbers (MT19937 seed 12345 + polar Box-Muller, libsrc/Random.cpp);
then the fitted null drives the A14 score step (BoltLMM::TestCovariate) end to end."""
import numpy as np
import pytest

from util import rel

pytestmark = pytest.mark.gpu


def _panel(seed, N, M, miss=0.01):
    rng = np.random.default_rng(seed)
    maf = rng.uniform(0.05, 0.5, M)
    G = rng.binomial(2, maf[:, None], size=(M, N)).astype(np.int8)
    G[rng.random((M, N)) < miss] = -1                       # missing calls: code 01 in the .bed
    G[3] = 0                                                # a monomorphic panel SNP (sd = 0 -> all-zero column)
    return G


def _pack(G):
    """(M, N) int8 with -1 = missing -> PLINK 2-bit rows (00 -> 0, 10 -> 1, 11 -> 2, 01 -> missing)"""
    code = np.where(G == 0, 0, np.where(G == 1, 2, np.where(G == 2, 3, 1))).astype(np.uint8)
    M, N = G.shape
    pad = (-N) % 4
    code = np.pad(code, ((0, 0), (0, pad)))
    c4 = code.reshape(M, -1, 4)
    return (c4[:, :, 0] | (c4[:, :, 1] << 2) | (c4[:, :, 2] << 4) | (c4[:, :, 3] << 6)).astype(np.uint8)


@pytest.mark.parametrize("case", [(101, 800, 300, 3, 0.5), (102, 1200, 500, 2, 0.2)])
def test_bolt_null_fit_vs_oracle(engine_cls, oracle, case):
    from oracle import bolt_oracle as BO
    seed, N, M, C, h2 = case
    G = _panel(seed, N, M)
    rng = np.random.default_rng(seed + 1)
    covar = np.column_stack([np.ones(N)] + [rng.normal(size=N) for _ in range(C - 1)])
    X, Z, _ = BO.prepare(G, covar, np.zeros(N))
    y = X @ rng.normal(size=M) * np.sqrt(h2 / M) + rng.normal(size=N) * np.sqrt(1 - h2) + covar @ rng.normal(size=C)
    X, Z, yc = BO.prepare(G, covar, y)
    ref = BO.Fit(X, Z, yc).fit().calibrate()
    eng = engine_cls(0)
    rec, h, Zd = eng.bolt_fit_null(_pack(G), N, y, covar)
    assert int(rec["mc_trials"]) == ref.mc == 15 and int(rec["n_covariates_kept"]) == C
    # same covariate space
    assert np.max(np.abs(Zd @ Zd.T - Z @ Z.T)) <= 1e-10
    n_ld = len(ref.log_delta)
    assert int(rec["reml_evals"]) == len(ref.f)
    assert np.max(np.abs(rec["log_delta"][:n_ld] - np.array(ref.log_delta))) <= 1e-6, (rec["log_delta"], ref.log_delta)
    assert np.max(np.abs(rec["f"][:len(ref.f)] - np.array(ref.f))) <= 1e-6
    assert int(rec["cg_iterations"]) == sum(ref.cg_iters)
    for k, v in (("delta", ref.delta), ("sigma2_g", ref.sigma2_g), ("sigma2_e", ref.sigma2_e), ("h2", ref.h2),
                 ("h_inv_y_norm2", ref.h_norm2), ("inf_stat_calibration", ref.calibration), ("xvx_xx_ratio", ref.xvx_xx_ratio)):
        assert rel(rec[k], v) <= 1e-6, (k, rec[k], v)
    assert np.max(np.abs(h[:N] - ref.h)) <= 1e-7 * np.max(np.abs(ref.h))
    assert 0.05 < rec["h2"] < 0.95
    # the fitted null drives the score step (A14): u = g'Ph, v = g'Pg |h|^2_proj calib / N
    hz = Zd.T @ h[:N]
    r_b = h[:N] - Zd @ hz
    kappa = float(rec["h_inv_y_norm2"] * rec["inf_stat_calibration"] / N)
    eng.set_null_residual(covar, r_b, kappa)
    Gt = rng.binomial(2, 0.2, size=(40, N)).astype(np.int8)
    eng.push_i8(Gt, None)
    vout, _b, _w = eng.meta_flush(40, want_cov=False)
    for j in range(40):
        g = Gt[j].astype(np.float64)
        zg = Z.T @ g
        u = float(g @ ref.h - zg @ (Z.T @ ref.h))
        v = float(g @ g - zg @ zg) * ref.h_norm2 * ref.calibration / N
        assert abs(vout[j]["U"] * kappa - u) <= 1e-6 * max(abs(u), np.sqrt(v))
        assert rel(vout[j]["pvalue"], oracle.lib().orc_chisq_q(u * u / v, 1.0)) <= 1e-5
    eng.close()


@pytest.mark.parametrize("k", [0, 1, 2])
def test_bolt_device_vs_reference_golden(engine_cls, oracle, k):
    """The device against outputs of the REFERENCE's own BoltLMM (regression/BoltLMM.cpp + BoltPlinkLoader.cpp compiled
    unmodified: tests/golden/ref_bolt_golden.npz, generator tests/golden/make_golden_ref_bolt.py), nothing in between:
    A15 the null fit (secant path, variance components, H^-1 y, calibration, xVx/xx), then A14 on top of the DEVICE's own
    fit -- TestCovariate (:315-338) per variant through rvt_set_null_residual + rvt_meta_flush, GetCovXX (:414-460) per pair
    through the band with option meta_cov_scale.  The reference is float32 and its secant steps amplify that noise
    (tests/test_oracle_pin_reference_bolt.py holds the path step by step): 5e-3 on what follows from log-delta, float32
    tolerances on the score step."""
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_bolt_golden.npz"))
    _seed, N, M, C, _h2 = (int(v) if i < 4 else v for i, v in enumerate(z[f"c{k}_case"]))
    bed, y, covar, Gt = z[f"c{k}_bed"], z[f"c{k}_y"], z[f"c{k}_covar"], z[f"c{k}_gtest"]
    ref = {key[len(f"c{k}_"):]: z[key] for key in z.files if key.startswith(f"c{k}_")}
    eng = engine_cls(0)
    try:
        rec, h, Zd = eng.bolt_fit_null(np.ascontiguousarray(bed), N, y, covar)
        n = len(ref["f"])
        assert int(rec["reml_evals"]) == n and int(rec["n_covariates_kept"]) == C and int(rec["mc_trials"]) == 15
        assert np.max(np.abs(rec["log_delta"][:n] - ref["log_delta"])) <= 5e-3
        assert np.max(np.abs(rec["f"][:n] - ref["f"])) <= 5e-3
        for mine, key in (("delta", "delta"), ("sigma2_g", "sigma2_g"), ("sigma2_e", "sigma2_e"), ("h2", "h2"),
                          ("h_inv_y_norm2", "H_inv_y_norm2"), ("inf_stat_calibration", "infStatCalibration"),
                          ("xvx_xx_ratio", "xVx_xx_ratio")):
            assert rel(rec[mine], float(ref[key])) <= 5e-3, (mine, rec[mine], float(ref[key]))
        assert np.max(np.abs(h[:N] - ref["H_inv_y"][:N])) <= 5e-3 * np.max(np.abs(ref["H_inv_y"][:N]))
        # A14 on the device's fit
        r_b = h[:N] - Zd @ (Zd.T @ h[:N])
        kappa = float(rec["h_inv_y_norm2"] * rec["inf_stat_calibration"] / N)
        eng.set_null_residual(covar, r_b, kappa)
        eng.set_option("meta_cov_scale", float(rec["xvx_xx_ratio"]))
        nv = Gt.shape[0]
        eng.push_i8(np.ascontiguousarray(Gt), None)
        pos = (100 * np.arange(nv)).astype(np.int32)
        vout, band, wmax = eng.meta_flush(nv, pos, np.ones(nv, dtype=np.int32), 100 * nv)
        for j in range(nv):
            af, ru, rv, reff, rp = ref["tests"][j]
            u, v = vout[j]["U"] * kappa, (vout[j]["sqrtV"] * kappa) ** 2     # U = g'r / kappa, V = g'Pg / kappa as MetaScore prints
            assert abs(u - ru) <= 5e-3 * np.sqrt(rv) and rel(v, rv) <= 5e-3, (j, u, ru, v, rv)
            assert rel(vout[j]["pvalue"], rp) <= 2e-2 or abs(vout[j]["pvalue"] - rp) <= 1e-6
        assert wmax >= nv - 1
        for (a, b), (c64, c32) in zip(ref["pairs"], ref["covxx"]):
            i, j = (int(a), int(b)) if a <= b else (int(b), int(a))
            got = band[i, j - i] * N                                   # printed divided by N (src/Model.cpp:990-996)
            scale = np.sqrt(band[i, 0] * band[j, 0]) * N
            assert abs(got - c64) <= 5e-3 * scale and abs(got - c32) <= 5e-3 * scale, (a, b, got, c64, c32)
    finally:
        eng.close()


def _bolt_rank(rank, world, port, case, q):
    """one rank of the SNP-sharded fit; both ranks share cuda:0, so the sum over ranks is staged through gloo"""
    import os
    import sys
    import torch.distributed as dist
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    from rvtests_b200 import engine, sharding
    from oracle import bolt_oracle as BO
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    seed, N, M, C, h2 = case
    G = _panel(seed, N, M)
    rng = np.random.default_rng(seed + 1)
    covar = np.column_stack([np.ones(N)] + [rng.normal(size=N) for _ in range(C - 1)])
    X, _, _ = BO.prepare(G, covar, np.zeros(N))
    y = X @ rng.normal(size=M) * np.sqrt(h2 / M) + rng.normal(size=N) * np.sqrt(1 - h2) + covar @ rng.normal(size=C)
    bed = _pack(G)
    lo, hi = sharding.snp_shard(M, rank, world)
    eng = engine.GeneEngine(0)
    calls = [0]
    ar = sharding.torch_allreduce(dist, device="cuda:0", staged=True)

    def counted(ptr, count, stream):
        calls[0] += 1
        return ar(ptr, count, stream)

    rec, h, _Z = eng.bolt_fit_null(np.ascontiguousarray(bed[lo:hi]), N, y, covar, M_total=M, m_offset=lo, allreduce=counted)
    out = {k: (rec[k].tolist() if hasattr(rec[k], "tolist") else rec[k]) for k in rec.dtype.names}
    if rank == 0:
        full, hf, _ = eng.bolt_fit_null(bed, N, y, covar)
        out["full"] = {k: (full[k].tolist() if hasattr(full[k], "tolist") else full[k]) for k in full.dtype.names}
        out["h_err"] = float(np.max(np.abs(h - hf)) / np.max(np.abs(hf)))
    eng.close()
    q.put((rank, calls[0], out, h[:8].tolist()))
    dist.destroy_process_group()


def test_bolt_snp_sharded_two_ranks_equals_unsharded():
    """rvt_bolt_fit_null_sharded with the panel's SNPs split over two ranks (two processes on this GPU, the sum over ranks
    through the callback) against the unsharded fit: same secant path, same CG iteration counts, same H^-1 y; one
    collective per H-product (+ one per |beta_hat|^2, + the calibration columns)."""
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    case = (105, 640, 333, 2, 0.4)
    ps = [ctx.Process(target=_bolt_rank, args=(r, 2, port, case, q)) for r in range(2)]
    for p in ps:
        p.start()
    outs = sorted((q.get(timeout=140) for _ in ps), key=lambda o: o[0])
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, calls0, o0, h0), (r1, calls1, o1, h1) = outs
    full = o0["full"]
    assert calls0 == calls1 > 0
    assert o0["cg_iterations"] == o1["cg_iterations"] == full["cg_iterations"]
    assert o0["reml_evals"] == o1["reml_evals"] == full["reml_evals"]
    # H-products: one per CG iteration + one per solve start; |beta_hat|^2 once per REML evaluation; x_beta_rand; the columns
    n_solves = full["reml_evals"] + 1
    assert calls0 == full["cg_iterations"] + n_solves + full["reml_evals"] + 2
    for k in ("delta", "sigma2_g", "sigma2_e", "h_inv_y_norm2", "inf_stat_calibration", "xvx_xx_ratio"):
        assert rel(o0[k], full[k]) <= 1e-9 and o0[k] == o1[k], (k, o0[k], o1[k], full[k])
    assert np.max(np.abs(np.array(o0["log_delta"]) - np.array(full["log_delta"]))) <= 1e-9
    assert o0["h_err"] <= 1e-9 and h0 == h1


def test_bolt_odd_stride_tail_bits_and_device_resident_panel(engine_cls, oracle):
    """N = 810: two samples in the last byte of a row (the padding bits stay out of every count), a row stride of 203 bytes.
    The engine's own copy of a host panel gets a 16-byte pitch (cp.async kernels); a panel that already lives in device memory
    is used in place, and with this stride that means the fall-back to the second-generation kernels -- same fit either way,
    and the fit of the restatement."""
    import torch
    from oracle import bolt_oracle as BO
    seed, N, M, C, h2 = 131, 810, 257, 2, 0.35
    G = _panel(seed, N, M, miss=0.03)
    rng = np.random.default_rng(seed + 1)
    covar = np.column_stack([np.ones(N), rng.normal(size=N)])
    X, Z, _ = BO.prepare(G, covar, np.zeros(N))
    y = X @ rng.normal(size=M) * np.sqrt(h2 / M) + rng.normal(size=N) * np.sqrt(1 - h2)
    X, Z, yc = BO.prepare(G, covar, y)
    ref = BO.Fit(X, Z, yc).fit().calibrate()
    bed = _pack(G)
    assert bed.shape[1] == 203
    eng = engine_cls(0)
    try:
        rec, h, _ = eng.bolt_fit_null(bed, N, y, covar)
        dev = torch.from_numpy(bed).cuda()
        rec_d, h_d, _ = eng.bolt_fit_null(None, N, y, covar, bed_dev=(dev.data_ptr(), M, dev.stride(0)))
        eng.set_option("bolt_kernels", 1)
        rec_1, h_1, _ = eng.bolt_fit_null(bed, N, y, covar)
    finally:
        eng.close()
    for r_, h_ in ((rec, h), (rec_d, h_d), (rec_1, h_1)):
        assert int(r_["reml_evals"]) == len(ref.f) and int(r_["cg_iterations"]) == sum(ref.cg_iters)
        assert np.max(np.abs(r_["log_delta"][:len(ref.log_delta)] - np.array(ref.log_delta))) <= 1e-6
        for k, v in (("delta", ref.delta), ("sigma2_g", ref.sigma2_g), ("h_inv_y_norm2", ref.h_norm2),
                     ("inf_stat_calibration", ref.calibration), ("xvx_xx_ratio", ref.xvx_xx_ratio)):
            assert rel(r_[k], v) <= 1e-6, (k, r_[k], v)
        assert np.max(np.abs(h_[:N] - ref.h)) <= 1e-7 * np.max(np.abs(ref.h))


def test_bolt_binary_mode_vs_oracle(engine_cls, oracle):
    """option bolt_binary = BoltLMM::enableBinaryMode (BASELINE configs[4]): the 0/1 phenotype is not centred
    (BoltPlinkLoader.cpp:155-158); the restatement in that mode is pinned on the reference's own build
    (tests/test_oracle_pin_reference_bolt.py::test_bolt_binary_mode_vs_live_reference_build)."""
    from oracle import bolt_oracle as BO
    seed, N, M, C = 151, 1000, 320, 3
    G = _panel(seed, N, M)
    rng = np.random.default_rng(seed + 1)
    covar = np.column_stack([np.ones(N)] + [rng.normal(size=N) for _ in range(C - 1)])
    X, Z, _ = BO.prepare(G, covar, np.zeros(N))
    liab = X @ rng.normal(size=M) * np.sqrt(0.5 / M) + rng.normal(size=N) * np.sqrt(0.5) + 0.3 * covar[:, 1]
    y = (liab > 0.5).astype(np.float64)
    X, Z, yc = BO.prepare(G, covar, y, binary=True)
    ref = BO.Fit(X, Z, yc).fit().calibrate()
    eng = engine_cls(0)
    try:
        eng.set_option("bolt_binary", 1)
        rec, h, Zd = eng.bolt_fit_null(_pack(G), N, y, covar)
        eng.set_option("bolt_binary", 0)
        rec_c, h_c, _ = eng.bolt_fit_null(_pack(G), N, y, covar)
    finally:
        eng.close()
    assert int(rec["reml_evals"]) == len(ref.f) and int(rec["cg_iterations"]) == sum(ref.cg_iters)
    assert np.max(np.abs(rec["log_delta"][:len(ref.log_delta)] - np.array(ref.log_delta))) <= 1e-6
    for k, v in (("delta", ref.delta), ("sigma2_g", ref.sigma2_g), ("h_inv_y_norm2", ref.h_norm2),
                 ("inf_stat_calibration", ref.calibration), ("xvx_xx_ratio", ref.xvx_xx_ratio)):
        assert rel(rec[k], v) <= 1e-6, (k, rec[k], v)
    assert np.max(np.abs(h[:N] - ref.h)) <= 1e-7 * np.max(np.abs(ref.h))
    # the two modes differ only inside the covariate space: the projected H^-1 y -- what the score step uses -- is the same
    pb, pc = h[:N] - Zd @ (Zd.T @ h[:N]), h_c[:N] - Zd @ (Zd.T @ h_c[:N])
    assert np.max(np.abs(pb - pc)) <= 1e-6 * np.max(np.abs(pb))
    assert np.max(np.abs(h[:N] - h_c[:N])) > 1e-3 * np.max(np.abs(h[:N]))
