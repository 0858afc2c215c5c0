"""bench.py --workload bolt: the BoltLMM null-model fit (BASELINE configs[4]) on a synthetic PLINK panel.

A "step" is one whole null fit (rvt_bolt_fit_null_sharded): MC-REML secant iteration on log(delta) with multi-RHS conjugate
gradients, then the calibration solve.  Strong scaling: the panel's SNP rows are sharded over the ranks, everything of
length N is replicated, ONE ncclAllReduce per H-product (rvtests_b200.sharding.torch_allreduce on the engine's stream).
Metric (both arms): panel genotype x right-hand-side multiply-adds per second,
    work = sum over H-products of 2 N M R   (two passes over the panel per H-product: X'v and X w, BoltLMM.cpp:942-966)
which is what the reference spends its time on too (its two GEMM loops over 64-SNP batches).

The panel is synthesised ON THE DEVICE in fixed 256-row chunks keyed by the global row index, so a rank's shard is the same
bytes whatever the world size; the phenotype has real heritability (256 causal panel SNPs), so the secant iteration behaves
as on data.  e2e: the same fit from a panel in PINNED HOST memory (H2D of N M / 4 bytes inside the timed region)."""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

BOLT_METRIC = "BoltLMM null fit: panel genotype x RHS multiply-adds per second (N x M_panel x R x 2 passes per H-product)"
SEED = 20260925
CHUNK = 256


def bolt_config(args):
    return {"workload": f"BoltLMM null fit (MC-REML secant + multi-RHS CG + calibration): N={args.bolt_samples} samples x "
                        f"M_panel={args.bolt_snps} SNPs (2-bit PLINK rows), C={args.covariates}, {max(min(int(4e9 / args.bolt_samples / args.bolt_samples), 15), 3)} MC trials "
                        "(BASELINE configs[4]: binary trait = 30 % cases of a liability with 256 causal panel SNPs, BoltLMM::enableBinaryMode; null fit + score test)",
            "samples": args.bolt_samples, "panel_snps": args.bolt_snps, "covariates_incl_intercept": args.covariates,
            "l2_policy": "every pass streams the 2-bit panel (N M / 4 bytes) and the N x R vectors; inputs >> 126 MB L2 at the "
                         "default size (4.1 GB panel + 32 MB per vector), no flush"}


def synth_rows(torch, dev, N, lo, hi, miss=0.01):
    """2-bit PLINK rows [lo, hi) of the synthetic panel as a (hi - lo, ceil(N/4)) uint8 device tensor (+ their MAFs)"""
    stride = (N + 3) // 4
    out = torch.empty((hi - lo, stride), dtype=torch.uint8, device=dev)
    c0 = (lo // CHUNK) * CHUNK
    while c0 < hi:
        g = torch.Generator(device=dev)
        g.manual_seed(SEED * 1_000_003 + c0)
        maf = 0.05 + 0.45 * torch.rand((CHUNK, 1), generator=g, device=dev)
        Npad = stride * 4
        a = (torch.rand((CHUNK, Npad), generator=g, device=dev) < maf).to(torch.uint8)
        a += (torch.rand((CHUNK, Npad), generator=g, device=dev) < maf).to(torch.uint8)
        # PLINK codes: 0 -> 00, 1 -> 10, 2 -> 11, missing -> 01
        code = torch.where(a == 0, 0, torch.where(a == 1, 2, 3)).to(torch.uint8)
        code[torch.rand((CHUNK, Npad), generator=g, device=dev) < miss] = 1
        if Npad > N:
            code[:, N:] = 0
        c4 = code.view(CHUNK, stride, 4)
        rows = c4[:, :, 0] | (c4[:, :, 1] << 2) | (c4[:, :, 2] << 4) | (c4[:, :, 3] << 6)
        a0, a1 = max(lo, c0), min(hi, c0 + CHUNK)
        out[a0 - lo:a1 - lo] = rows[a0 - c0:a1 - c0]
        c0 += CHUNK
        del a, code, c4, rows
    return out


def phenotype(torch, dev, N, M, C, h2=0.4, binary=True):
    """y = sum of 256 causal panel SNPs (normalised) * effect + covariates + noise, replicated on every rank"""
    rng = np.random.default_rng(SEED)
    causal = np.unique(np.linspace(0, M - 1, 256).astype(np.int64))
    beta = rng.normal(size=len(causal)) * np.sqrt(h2 / len(causal))
    g = np.zeros(N)
    dec = torch.tensor([0.0, float("nan"), 1.0, 2.0], dtype=torch.float64, device=dev)
    for b, m in zip(beta, causal):
        row = synth_rows(torch, dev, N, int(m), int(m) + 1)[0]
        codes = torch.stack([(row >> s) & 3 for s in (0, 2, 4, 6)], dim=1).reshape(-1)[:N].long()
        x = dec[codes]
        mu = torch.nanmean(x)
        sd = torch.sqrt(mu * (1 - mu / 2))
        x = torch.nan_to_num((x - mu) / sd, nan=0.0)
        g += b * x.cpu().numpy()
    covar = np.column_stack([np.ones(N)] + [rng.normal(size=N) for _ in range(C - 1)])
    y = g + rng.normal(size=N) * np.sqrt(1 - h2) + covar @ rng.normal(size=C)
    if binary:   # BASELINE configs[4] is a binary trait: cases = the upper 30 % of the liability
        y = (y > np.quantile(y, 0.7)).astype(np.float64)
    return y, covar


def run_bolt(args, ClockSampler, bind_numa):
    import torch
    import torch.distributed as dist
    import rvtests_b200
    from rvtests_b200 import sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    numa = bind_numa(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    stream = torch.cuda.current_stream()
    N, M, C = args.bolt_samples, args.bolt_snps, args.covariates
    lo, hi = sharding.snp_shard(M, rank, world)
    panel = synth_rows(torch, dev, N, lo, hi)
    y, covar = phenotype(torch, dev, N, M, C)
    eng = rvtests_b200.GeneEngine(local)
    eng.set_stream(stream.cuda_stream)
    eng.set_option("bolt_binary", 1)     # BoltLMM::enableBinaryMode (the phenotype is 0/1 and stays uncentred)
    if os.environ.get("RVT_BOLT_KERNELS"):
        eng.set_option("bolt_kernels", float(os.environ["RVT_BOLT_KERNELS"]))      # A/B of the product kernels (profiles/)
    ar = sharding.torch_allreduce(dist, device=dev) if world > 1 else None

    def fit(bed_host=None):
        kw = dict(M_total=M, m_offset=lo, allreduce=ar) if world > 1 else {}
        if bed_host is not None:
            return eng.bolt_fit_null(bed_host, N, y, covar, **kw)
        return eng.bolt_fit_null(None, N, y, covar, bed_dev=(panel.data_ptr(), hi - lo, panel.stride(0)), **kw)

    steps, warm = max(1, args.steps), max(1, args.warmup)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(warm):
        rec, h, Z = fit()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    recs = []
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(steps):
        rec, h, Z = fit()
        recs.append(rec.copy())
    e1.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    wall = (time.perf_counter() - t0) / steps
    clocks = sampler.stop() if rank == 0 else None
    tmax = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms = float(tmax.item()) / steps
    rec = recs[-1]
    R1 = int(rec["mc_trials"]) + 1
    nS = min(30, M)
    hx = int(rec["h_products"])
    # R-weighted H-products: the REML solves carry R1 = MCtrial + 1 right-hand sides, the calibration solve (the last one) nS
    hx_cal = int(rec["h_products_calibration"])
    work = 2.0 * N * M * 2.0 * (R1 * (hx - hx_cal) + nS * hx_cal)        # multiply-adds x 2 passes, whole job (all ranks)
    value = work / 2.0 / (ms * 1e-3)                                    # multiply-adds per second
    # ---- e2e: the panel in pinned host memory, H2D inside the timed region
    e2e = None
    if not args.no_e2e:
        host = torch.empty((hi - lo, panel.shape[1]), dtype=torch.uint8, pin_memory=True)
        host.copy_(panel)
        hb = host.numpy()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t = time.perf_counter()
        e0.record(stream)
        rec2, _, _ = fit(hb)
        e1.record(stream)
        torch.cuda.synchronize()
        tm = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        e2e = {"value": work / 2.0 / (float(tm.item()) * 1e-3), "unit": "multiply-adds/s", "h2d_bytes_per_step": int(hb.nbytes + 8 * N * (C + 1)),
               "d2h_bytes_per_step": int(8 * (N + C) + rec.nbytes), "wall_ms": (time.perf_counter() - t) * 1e3,
               "what": "rvt_bolt_fit_null(_sharded) with the 2-bit panel in pinned host memory; H^-1 y and the record back to the host"}
    # ---- the other half of BASELINE configs[4]: the score test of the variants of this rank's shard on the fitted null
    # (BoltLMM::TestCovariate, regression/BoltLMM.cpp:315-338 = rvt_set_null_residual + rvt_meta_flush, score columns only)
    score = None
    if not getattr(args, "no_score", False):
        from rvtests_b200.synth import variant_params
        nv = args.meta_variants
        hN = h[:N]
        r_b = hN - Z @ (Z.T @ hN)
        kappa = float(rec["h_inv_y_norm2"] * rec["inf_stat_calibration"] / N)
        eng.set_null_residual(covar, r_b, kappa)
        keys, t0v, t1v = variant_params(SEED, rank * nv, nv)
        eng.synth_load(keys, t0v, t1v, nv // 64, 64)
        for _ in range(2):
            eng.push_loaded()
            vout, _b, _w = eng.meta_flush(nv, want_cov=False)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        reps = 5
        e0.record(stream)
        for _ in range(reps):
            eng.push_loaded()
            vout, _b, _w = eng.meta_flush(nv, want_cov=False)
        e1.record(stream)
        torch.cuda.synchronize()
        ts = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        ms_s = float(ts.item()) / reps
        score = {"value": world * nv / (ms_s * 1e-3), "unit": "variants/s", "variants_per_gpu_per_step": nv, "ms_per_step": ms_s,
                 "what": "BoltLMM::TestCovariate on resident 64-variant tiles of N samples: rvt_set_null_residual(H^-1 y projected, kappa) "
                         "+ rvt_meta_flush (score columns, exact HWE, records to the host); 500 000 variants = "
                         f"{500000 / max(world * nv / (ms_s * 1e-3), 1e-9):.2f} s on {world} GPU(s)",
                 "ok": int((vout["ok"] == 1).sum()), "median_p": float(np.median(vout["pvalue"])),
                 "note": "the test variants are independent of phenotype and panel; infStatCalibration is estimated on IN-PANEL SNPs "
                         "(the reference's rule, BoltLMM.cpp:1141-1186) and falls far below 1 when the panel is small against N "
                         f"(here M/N = {M / N:.3f}, calibration {float(rec['inf_stat_calibration']):.3f}), which inflates the statistics of "
                         "out-of-panel variants: a property of this synthetic shape under the reference's algorithm, not of the kernels"}
    if rank != 0:
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6451.8))
    ms_x = float(rec["ms_xtv"]) + float(rec["ms_xw"])
    # per rank: both products stream the rank's shard of the panel once per H-product
    bytes_alg = 2.0 * (hx * (N / 4.0) * (hi - lo))
    flops = work / world                                                # work = 2 flop x (2 N M R multiply-adds) per H-product; this rank's shard
    fp64_peak = 148 * 64 * 2 * 1.965e9 / 1e12                            # 148 SMs x 64 fp64 FMA lanes x 2 x 1.965 GHz = 37.2 TFLOP/s
    line = {
        "metric": BOLT_METRIC, "value": value, "unit": "multiply-adds/s", "n_gpus": world, "steps": steps, "warmup": warm,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": bolt_config(args),
        "engine": {"parallelism": f"panel SNP rows sharded over {world} rank(s); one ncclAllReduce of (N + C) x R doubles per H-product",
                   "numa": numa, "allreduce_calls_per_fit": int(rec["allreduce_calls"])},
        "fit": {"wall_ms": wall * 1e3, "h_products": hx, "cg_iterations": int(rec["cg_iterations"]), "reml_evals": int(rec["reml_evals"]),
                "delta": float(rec["delta"]), "h2": float(rec["h2"]), "sigma2_g": float(rec["sigma2_g"]),
                "inf_stat_calibration": float(rec["inf_stat_calibration"]), "xvx_xx_ratio": float(rec["xvx_xx_ratio"]),
                "log_delta": [float(v) for v in rec["log_delta"][: int(rec["reml_evals"]) + 1]]},
        "kernel_ms_per_step": {"xtv": float(rec["ms_xtv"]), "xw": float(rec["ms_xw"])},
        "gpu_launches": int(2 * hx * (2 if R1 > 16 else 1) + 12 * hx),
        "clocks": clocks,
        "roofline": {"bound": "fp64 pipe (CUDA cores; no GEMM shape: a 4-entry table decode per genotype feeds R multiply-adds)",
                     "kernel": "k_bolt_xtv3 + k_bolt_xw3", "achieved": flops / (ms_x * 1e-3) / 1e12, "peak": fp64_peak,
                     "unit": "TFLOP/s", "frac": flops / (ms_x * 1e-3) / 1e12 / fp64_peak,
                     "peak_source": "148 SMs x 64 fp64 lanes x 2 x 1.965 GHz (nominal; MEASURED_PEAKS.json holds no fp64 figure)",
                     "hbm": {"achieved": bytes_alg / (ms_x * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                             "frac": bytes_alg / (ms_x * 1e-3) / 1e9 / hbm, "algorithmic_bytes_per_step": bytes_alg},
                     "traffic": None, "share_of_step": ms_x / ms},
        "e2e": e2e,
        "score_test": score,
    }
    print(json.dumps(line))


def run_reference_bolt(args):
    """--impl reference --workload bolt: the REFERENCE's own BoltLMM::FitNullModel (oracle/_ref/libbolt_ref.so: BoltLMM.cpp +
    BoltPlinkLoader.cpp compiled unmodified, float32, OpenMP over the 64 SNPs of a batch as upstream) on a bounded sample of
    the workload; its H-products are counted by the numpy restatement on the same data (same path, pinned in
    tests/test_oracle_pin_reference_bolt.py)."""
    import tempfile
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from oracle import oracle as orc
    from oracle import bolt_oracle as BO
    L = orc.ref_bolt()
    N, M, C = args.bolt_ref_samples, args.bolt_ref_snps, args.covariates
    rng = np.random.default_rng(SEED)
    G = rng.binomial(2, rng.uniform(0.05, 0.5, M)[:, None], size=(M, N)).astype(np.int8)
    G[rng.random((M, N)) < 0.01] = -1
    covar = np.column_stack([np.ones(N)] + [rng.normal(size=N) for _ in range(C - 1)])
    X, Z, _ = BO.prepare(G, covar, np.zeros(N))
    causal = np.unique(np.linspace(0, M - 1, 256).astype(np.int64))
    y = X[:, causal] @ (rng.normal(size=len(causal)) * np.sqrt(0.4 / len(causal))) + rng.normal(size=N) * np.sqrt(0.6) + covar @ rng.normal(size=C)
    y = np.array([float("%.9g" % v) for v in y])
    covar = np.array([[float("%.9g" % v) for v in row] for row in covar])
    X, Z, yc = BO.prepare(G, covar, y)
    fit = BO.Fit(X, Z, yc).fit().calibrate()
    R1, nS = fit.mc + 1, min(30, M)
    hx_w = sum((it + 1) * (nS if k == len(fit.cg_iters) - 1 else R1) for k, it in enumerate(fit.cg_iters))
    work = 2.0 * N * M * 2.0 * hx_w
    kind, vals = "reference", []
    with tempfile.TemporaryDirectory() as d:
        prefix = os.path.join(d, "panel")
        orc.write_bolt_fileset(prefix, G, y, covar)
        for k in range(args.warmup + args.steps):
            t = time.perf_counter()
            if L is not None:
                rc = L.bolt_ref_fit(prefix.encode(), None, 0, None, None, 0)
                assert rc == 0
            else:
                kind = "port"
                BO.Fit(X, Z, yc).fit().calibrate()
            dt = time.perf_counter() - t
            if k >= args.warmup:
                vals.append(dt)
    if L is not None:
        L.bolt_ref_free()
    dt = float(np.mean(vals))
    value = work / 2.0 / dt
    cores = int(os.environ.get("OMP_NUM_THREADS", "0")) or (os.cpu_count() or 1)
    sample = (f"each step = one whole null fit at N={N} x M_panel={M} (the device arm runs N={args.bolt_samples} x {args.bolt_snps}); "
              f"{'the reference build itself (BoltLMM.cpp + BoltPlinkLoader.cpp unmodified, float32, fileset read + fit, one thread; Eigen is absent here, its GEMMs are the plain loops of oracle/eigen_standin, so a real Eigen build is faster than this figure)' if kind == 'reference' else 'numpy restatement (no reference build on this box)'}; "
              f"{sum(fit.cg_iters) + len(fit.cg_iters)} H-products")
    print(json.dumps({
        "impl": "reference", "metric": BOLT_METRIC, "value": value, "unit": "multiply-adds/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32" if kind == "reference" else "f64", "data": "synthetic", "config": bolt_config(args),
        "cpu_baseline": {"value": value, "unit": "multiply-adds/s", "cores": 1, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "multiply-adds/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
