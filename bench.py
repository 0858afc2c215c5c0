#!/usr/bin/env python
"""bench.py -- gene-sets/sec of the per-gene SKAT (+CMC, Zeggini) hot path on B200.

  python bench.py --gpus N --steps K --warmup W            # our arm (one rank per GPU; torchrun for N>1)
  python bench.py --impl reference --gpus N --steps K --warmup W   # the reference algorithm on the host cores

A "step" is one pass of the hot path over the genes resident on each GPU: for every gene one
sweep over its packed N x M genotype block (K1), then eigenvalues + Davies/Liu + burden score tests
(K2/K3), results left in HBM and (N>1) gathered once over NCCL.  Workload (weak scaling): each
rank holds --genes genes of --samples x --variants synthetic HWE genotypes (SURVEY.md 8(d) stream,
seed 20260925), i.e. BASELINE.json configs[2] (500k samples, 20 000 genes x 50 variants over 8 GPUs
= 2 500 genes per GPU) at 8 ranks; the metric's "500k samples x 50 variants" at every N.
The data set per rank (62.5 GB) is ~500x L2, so every step streams from HBM (no L2 flush needed).

The one JSON line on stdout follows the driver contract; see DESIGN.md section 6 for how each
field is measured.  oracle/ is used ONLY by the cpu_baseline leg and by --impl reference.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SEED = 20260925
METRIC = "gene-sets/sec (SKAT, 500k samples x 50 variants)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--samples", type=int, default=500_000)
    ap.add_argument("--variants", type=int, default=50)
    ap.add_argument("--genes", type=int, default=2500, help="genes resident per GPU")
    ap.add_argument("--covariates", type=int, default=3, help="columns of X incl. intercept")
    ap.add_argument("--engine", type=int, default=0, help="0 auto, 1 dp4a, 2 tcgen05")
    ap.add_argument("--e2e-genes", type=int, default=256, help="genes per step of the host-buffer (e2e) leg (a quarter of it for the int8 form)")
    ap.add_argument("--cpu-genes", type=int, default=16, help="distinct genes of the CPU sample")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of the baseline sample")
    ap.add_argument("--workload", default="skat", choices=["skat", "meta", "bolt"],
                    help="skat: the headline metric (default).  meta: --meta score,cov at the BASELINE configs[3] shape.  "
                         "bolt: the BoltLMM null fit (BASELINE configs[4]), panel SNPs sharded over the ranks (tools/bolt_bench.py)")
    ap.add_argument("--bolt-samples", type=int, default=1_000_000, help="N of --workload bolt (BASELINE configs[4]: 1M samples)")
    ap.add_argument("--bolt-snps", type=int, default=16_384, help="panel SNPs of --workload bolt (whole job; 4.1 GB of 2-bit rows at 1M samples)")
    ap.add_argument("--bolt-ref-samples", type=int, default=4000, help="N of the bounded sample of --impl reference --workload bolt")
    ap.add_argument("--bolt-ref-snps", type=int, default=4000, help="panel SNPs of that sample")
    ap.add_argument("--meta-variants", type=int, default=8192, help="variants per GPU and step of --workload meta")
    ap.add_argument("--meta-spacing", type=int, default=1000, help="bp between consecutive variants (window 1 Mb)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-skato", action="store_true", help="skip the SKAT-O arm (profiler captures of the headline step)")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "50", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        sm = [float(r[1]) for r in self.rows if len(r) > 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) > 8 and r[2].replace(".", "").isdigit()]
        pw = [float(r[3]) for r in self.rows if len(r) > 8 and r[3].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) > 8 for i in range(4) if r[5 + i].lower() == "active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": reasons}


def workload_config(args):
    """the `config` object of BOTH arms (the driver compares them: same workload, same string)"""
    return {"workload": f"SKAT+CMC+Zeggini, N={args.samples} samples x M={args.variants} variants, "
                        f"{args.genes} genes per GPU (BASELINE configs[2] sharded: 2500 genes/GPU), C={args.covariates}",
            "genes_per_gpu": args.genes, "samples": args.samples, "variants": args.variants,
            "covariates_incl_intercept": args.covariates, "kernel_flags": "skat[nPerm=0] + cmc + zeggini",
            "l2_policy": "inputs (62.5 GB/rank at defaults) >> 126 MB L2; no flush needed"}


def bind_numa(local):
    """Bind this rank (threads + future allocations, i.e. its pinned host buffers) to the NUMA node of its GPU, so that the
    H2D copies of 8 ranks do not all pull from one socket's memory.  Best effort; reports what it did."""
    info = {"bound": False}
    try:
        import ctypes
        import torch
        pr = torch.cuda.get_device_properties(local)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        info.update({"gpu_pci": bdf, "gpu_numa_node": node})
        if node < 0:
            node = _numa_from_topo(local, info)
        if node < 0:
            # no placement information at all (virtualised PCI tree): at least spread this job's pinned pages over every
            # memory node the cpuset allows, so that N ranks do not all pull their H2D copies from one socket's DRAM
            nodes = _allowed_mem_nodes()
            info["mem_nodes_allowed"] = nodes
            if len(nodes) > 1:
                libc = ctypes.CDLL(None, use_errno=True)
                mask = (ctypes.c_ulong * 16)()
                for nd in nodes:
                    mask[nd // 64] |= 1 << (nd % 64)
                rc = libc.syscall(238, 3, mask, 16 * 64 + 1)   # set_mempolicy(MPOL_INTERLEAVE)
                info["mempolicy"] = "interleave over %s" % nodes if rc == 0 else "set_mempolicy errno %d" % ctypes.get_errno()
            return info
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        info["allowed_cpus"] = len(allowed)
        if use:
            os.sched_setaffinity(0, use)
            info["cpus"] = len(use)
            info["bound"] = True
        # memory policy MPOL_PREFERRED (1) on the node: first-touch / pinned allocations come from it when the cpuset allows
        libc = ctypes.CDLL(None, use_errno=True)
        mask = (ctypes.c_ulong * 16)()
        mask[node // 64] = 1 << (node % 64)
        rc = libc.syscall(238, 1, mask, 16 * 64 + 1)   # set_mempolicy, x86-64
        info["mempolicy"] = "preferred node %d" % node if rc == 0 else "set_mempolicy errno %d" % ctypes.get_errno()
    except Exception as e:   # never let a placement hint cost the bench line
        info["error"] = repr(e)
    return info


def _allowed_mem_nodes():
    try:
        for line in open("/proc/self/status"):
            if line.startswith("Mems_allowed_list:"):
                out = []
                for part in line.split(":", 1)[1].strip().split(","):
                    a, _, b = part.partition("-")
                    out.extend(range(int(a), int(b or a) + 1))
                return out
    except Exception:
        pass
    return []


def _numa_from_topo(local, info):
    """`nvidia-smi topo -m` prints a NUMA Affinity column even where sysfs says -1"""
    try:
        txt = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
        head = None
        for line in txt.splitlines():
            cells = [c.strip() for c in line.split("\t") if c.strip()]
            if not cells:
                continue
            if head is None and any("NUMA Affinity" in c for c in cells):
                head = cells
                continue
            if head and cells[0] == f"GPU{local}":
                # the data row has one more leading cell (the row label) than the header
                idx = [i for i, c in enumerate(head) if "NUMA Affinity" in c][0] + 1
                if idx < len(cells) and cells[idx].split(",")[0].split("-")[0].isdigit():
                    node = int(cells[idx].split(",")[0].split("-")[0])
                    if node in _allowed_mem_nodes():
                        info["gpu_numa_node_from_topo"] = node
                        return node
    except Exception as e:
        info["topo_error"] = repr(e)
    return -1


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return None


def cohort(args, rank):
    """keys / thresholds of this rank's variants + the shared covariates and trait."""
    from rvtests_b200.synth import variant_params, covariates
    vid0 = rank * args.genes * args.variants
    keys, t0, t1 = variant_params(SEED, vid0, args.genes * args.variants)
    X, y = covariates(SEED, args.samples, args.covariates)
    return keys, t0, t1, X, y


# ---------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import rvtests_b200

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    numa = bind_numa(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    stream = torch.cuda.current_stream()

    keys, t0, t1, X, y = cohort(args, rank)
    eng = rvtests_b200.GeneEngine(local)
    eng.set_stream(stream.cuda_stream)
    eng.set_option("engine", args.engine)
    eng.set_null_model(X, y)
    eng.synth_load(keys, t0, t1, args.genes, args.variants)

    rec = rvtests_b200.engine.RESULT_DTYPE.itemsize
    d_res = torch.empty(args.genes * rec, dtype=torch.uint8, device=dev)
    d_all = torch.empty(world * args.genes * rec, dtype=torch.uint8, device=dev) if world > 1 else None

    def step():
        n = eng.run_loaded(d_res.data_ptr())
        if world > 1:
            # the single collective of the path: one gather of the per-gene summary records
            dist.all_gather_into_tensor(d_all, d_res)
        return n

    sweep_ms, fin_ms, launches = [], [], 0
    # clocks / throttle reasons are sampled under load: from the warm-up steps through the timed region
    # (the timed region alone is a few tens of ms, shorter than nvidia-smi's sampling period)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    t_w = time.perf_counter()
    nw = 0
    while nw < max(args.warmup, 3) or time.perf_counter() - t_w < 0.5:
        step()
        nw += 1
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    w0 = time.perf_counter()
    e0.record(stream)
    for _ in range(args.steps):
        step()
        t = eng.last_timing()
        sweep_ms.append(t["sweep_ms"])
        fin_ms.append(t["finalize_ms"])
        launches += t["launches"]
    e1.record(stream)
    torch.cuda.synchronize()
    w1 = time.perf_counter()
    if world > 1:
        dist.barrier()
    dev_ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    tmax = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_total = float(tmax.item())
    ms_per_step = ms_total / args.steps
    value = world * args.genes / (ms_per_step * 1e-3)

    out = None
    if rank == 0:
        res = np.frombuffer(d_res.cpu().numpy().tobytes(), dtype=rvtests_b200.engine.RESULT_DTYPE)
        pk = peaks()
        hbm_peak = pk["hbm_gbs"] if pk else 6650.0
        alg_bytes = float(args.genes) * args.samples * args.variants  # 1 byte per genotype
        sweep_s = float(np.mean(sweep_ms)) * 1e-3
        achieved = alg_bytes / sweep_s / 1e9
        # DRAM bytes of the dominant kernel from the committed ncu --set full capture (per launch,
        # scaled from the captured launch by its traffic/algorithmic ratio)
        traffic, traffic_src = None, None
        try:
            import glob
            latest = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_summary.json")))[-1]
            prof = json.load(open(latest))
            ratio = prof["k_sweep_tc"]["traffic_over_algorithmic"]
            traffic = ratio * alg_bytes
            traffic_src = "profiles/%s: dram__bytes_read+write = %.4f x algorithmic bytes (ncu --set full, 512-gene launch)" % (os.path.basename(latest), ratio)
        except Exception:
            pass
        # "always print both fractions" -- the same launch against the tensor pipe.  Algorithmic work per gene
        # = 2 N M^2 (Gram) + 2 N M (C + 1) (G'X, G'r); the s8 x s8 -> s32 tcgen05 operations are counted as flops and held
        # against the MEASURED dense bf16 throughput (the int8 pipe's own peak is twice that); sustained figure because the
        # kernel is timed inside a long step.
        alg_flops = float(args.genes) * (2.0 * args.samples * args.variants ** 2
                                         + 2.0 * args.samples * args.variants * (args.covariates + 1))
        tens_peak = (pk.get("bf16_tflops_sustained") or pk.get("bf16_tflops")) if pk else 1590.0
        tens_ach = alg_flops / sweep_s / 1e12
        out = {
            "metric": METRIC, "value": value, "unit": "gene-sets/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "s8 x s8 -> s32 (exact) + f64 tail", "data": "synthetic",
            "config": workload_config(args),
            "engine": {"name": {1: "dp4a", 2: "tcgen05.kind::i8"}.get(int(eng.info("last_engine")), "?"),
                       "splits": int(eng.info("last_splits")),
                       "parallelism": f"genes sharded over {world} rank(s), one NCCL all_gather of result records",
                       "numa": numa},
            "wall_ms_per_step": (w1 - w0) * 1e3 / args.steps,
            "kernel_ms_per_step": {"sweep": float(np.mean(sweep_ms)), "finalize": float(np.mean(fin_ms))},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "k_sweep_tc" if int(eng.info("last_engine")) == 2 else "k_sweep_simt",
                         "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if pk else "fallback 6.65 TB/s (of fallback)",
                         "algorithmic_bytes_per_launch": alg_bytes, "traffic": traffic, "traffic_source": traffic_src,
                         "note": "peak is the driver's copy measurement (half reads, half writes); this kernel only reads, "
                                 "and a read-only stream can run a few % above a copy, hence frac may exceed 1",
                         "tensor": {"achieved": tens_ach, "peak": tens_peak, "unit": "TFLOP/s", "frac": tens_ach / tens_peak,
                                    "algorithmic_flops_per_launch": alg_flops,
                                    "peak_source": ("MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if pk
                                                    else "fallback 1.59 PFLOP/s (of fallback)"),
                                    "note": "secondary: int8 storage at 1 byte per genotype makes the sweep HBM-bound "
                                            "(2M/1 = 100 op/B at M = 50, below the ridge); s8 tcgen05 ops counted as flops "
                                            "against the dense bf16 peak (the int8 peak is 2x that)"}},
            "sanity": {"genes_ok": int((res["status"] == 0).sum()), "median_p_skat": float(np.median(res["p_skat"])),
                       "davies_fault_frac": float((res["davies_fault"] != 0).mean())},
        }
    # ---- the same step with SKAT-O on: BASELINE configs[2] is `--kernel skat,skato --burden cmc,zeggini`
    skato = None if args.no_skato else run_skato_arm(args, eng, torch, dist, world, rank, dev, stream, d_res, d_all)
    res_skato = None
    if rank == 0 and skato is not None:
        res_skato = np.frombuffer(d_res.cpu().numpy().tobytes(), dtype=rvtests_b200.engine.RESULT_DTYPE).copy()
    # ---- e2e: HOST buffers through the C ABI, H2D + D2H inside the timed region (rank-local, all ranks)
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, eng, torch, dist, world, rank, dev, "bed")
        e2e_i8 = run_e2e(args, eng, torch, dist, world, rank, dev, "i8")
        e2e_f64 = run_e2e(args, eng, torch, dist, world, rank, dev, "f64")
        e2e_miss = run_e2e(args, eng, torch, dist, world, rank, dev, "bed_missing")
        e2e_bin = run_e2e(args, eng, torch, dist, world, rank, dev, "bed_binary")
    if rank == 0:
        out["e2e"] = e2e
        if e2e is not None:
            out["e2e_int8"] = e2e_i8
            out["e2e_f64"] = e2e_f64
            out["e2e_missing_calls"] = e2e_miss
            out["e2e_binary_trait"] = e2e_bin
        out["skato"] = skato
        if not args.no_cpu and world == 1:
            # the oracle as CHECKER of the very records the timed steps produced (parity gates measured in
            # the same run), then as the timed CPU baseline
            out["parity"] = parity_gate(args, eng, res, res_skato)
            out["cpu_baseline"] = cpu_baseline(args, threads=os.cpu_count())
            if skato is not None:
                skato["cpu_baseline"] = cpu_baseline_skato(args)
        print(json.dumps(out))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def run_skato_arm(args, eng, torch, dist, world, rank, dev, stream, d_res, d_all):
    """K more steps over the same resident genes with SKAT-O enabled (skat + skato + cmc + zeggini per gene)."""
    eng.set_option("skato", 1)
    try:
        def step():
            eng.run_loaded(d_res.data_ptr())
            if world > 1:
                dist.all_gather_into_tensor(d_all, d_res)
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        steps = max(2, min(args.steps, 5))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        sweep_ms, fin_ms = [], []
        for _ in range(steps):
            step()
            t = eng.last_timing()
            sweep_ms.append(t["sweep_ms"])
            fin_ms.append(t["finalize_ms"])
        e1.record(stream)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        tmax = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms = float(tmax.item()) / steps
    finally:
        eng.set_option("skato", 0)
    return {"value": world * args.genes / (ms * 1e-3), "unit": "gene-sets/s", "ms_per_step": ms, "steps": steps,
            "kernel_flags": "skat[nPerm=0] + skato + cmc + zeggini (BASELINE configs[2] flags)",
            "kernel_ms_per_step": {"sweep": float(np.mean(sweep_ms)), "finalize+skato": float(np.mean(fin_ms))}}


def parity_gate(args, eng, res, res_skato, n_check=8):
    """oracle (oracle/skat_oracle.c, oracle/skato_oracle.py -- pinned on the reference build) on the first n_check resident
    genes vs the records of the timed steps: NonRefSite exact, Q 1e-6, p 1e-4, identical Davies fault flags."""
    from oracle import oracle as O
    from oracle import skato_oracle as SO
    O.build()
    from rvtests_b200.synth import covariates
    M, N = args.variants, args.samples
    n_check = min(n_check, args.genes)
    X, y = covariates(SEED, N, args.covariates)
    nm = O.fit_null_linear(X, y)

    def rel(a, b):
        a, b = float(a), float(b)
        return 0.0 if a == b else abs(a - b) / max(abs(a), abs(b), 1e-300)

    worst = {"Q": 0.0, "p_skat": 0.0, "cmc_p": 0.0, "zeg_p": 0.0, "skato_Q": 0.0, "skato_p": 0.0}
    exact = {"nonref": True, "davies_fault": True, "m_poly": True, "skato_rho": True}
    for g in range(n_check):
        Gg = eng.loaded_read(g * M, M)                      # (M, N) int8, the resident genotypes themselves
        Gd = Gg.T.astype(np.float64)
        af = 0.5 * Gg.sum(axis=1, dtype=np.int64) / N
        ref, lam = O.gene(Gd, af, X, nm["resid"], nm["sigma2"])
        r = res[g]
        exact["nonref"] &= int(r["cmc_nonref"]) == ref.cmc_nonref
        exact["m_poly"] &= int(r["m_poly"]) == ref.m_poly
        exact["davies_fault"] &= int(r["davies_fault"]) == ref.skat.fault
        worst["Q"] = max(worst["Q"], rel(r["Q"], ref.skat.Q))
        worst["p_skat"] = max(worst["p_skat"], rel(r["p_skat"], ref.skat.pvalue))
        worst["cmc_p"] = max(worst["cmc_p"], rel(r["cmc_p"], ref.cmc_p))
        worst["zeg_p"] = max(worst["zeg_p"], rel(r["zeg_p"], ref.zeg_p))
        if res_skato is not None:
            so = SO.skato_gene(Gd, af, X, nm["resid"])
            rs = res_skato[g]
            exact["skato_rho"] &= bool(so["ok"]) and int(rs["skato_ok"]) == 1 and float(rs["skato_rho"]) == float(so["rho"])
            worst["skato_Q"] = max(worst["skato_Q"], rel(rs["skato_Q"], so["Q"]))
            worst["skato_p"] = max(worst["skato_p"], rel(rs["skato_p"], so["pvalue"]))
            worst["Q"] = max(worst["Q"], rel(rs["Q"], ref.skat.Q))
    ok = (all(exact.values()) and worst["Q"] <= 1e-6 and worst["skato_Q"] <= 1e-6
          and all(worst[k] <= 1e-4 for k in ("p_skat", "cmc_p", "zeg_p", "skato_p")))
    return {"genes_checked": n_check, "pass": bool(ok), "max_rel_err": worst, "exact": exact,
            "tolerances": {"Q": 1e-6, "p": 1e-4, "counts_and_fault_flags": "exact"},
            "checker": "oracle/skat_oracle.c + oracle/skato_oracle.py (pinned on the reference's own Skat.cpp / SkatO.cpp build)"}


def cpu_baseline_skato(args, n_genes=4):
    """SKAT-O on the host, 1 thread: the numpy restatement of SkatO::Fit with the reference's own Davies (C) and GSL
    quadrature -- a bounded sample of the same workload."""
    from oracle import oracle as O
    from oracle import skato_oracle as SO
    O.build()
    from rvtests_b200.synth import variant_params, covariates
    M, N = args.variants, args.samples
    keys, t0, t1 = variant_params(SEED, 0, n_genes * M)
    G = O.synth_rows_f64(keys, t0, t1, N, threads=0).reshape(n_genes, M, N)
    X, y = covariates(SEED, N, args.covariates)
    nm = O.fit_null_linear(X, y)
    t = time.perf_counter()
    for g in range(n_genes):
        Gd = np.ascontiguousarray(G[g].T)
        SO.skato_gene(Gd, 0.5 * Gd.sum(axis=0) / N, X, nm["resid"])
    dt = time.perf_counter() - t
    return {"value": n_genes / dt, "unit": "gene-sets/s", "cores": 1, "kind": "port",
            "sample": f"{n_genes} genes of N={N} x M={M} in {dt:.1f} s: SKAT-O only (oracle/skato_oracle.py: numpy N x M algebra, "
                      "the reference's qfc.c Davies and GSL 1.16 qags), single thread"}


def run_e2e(args, eng, torch, dist, world, rank, dev, fmt="bed"):
    """same metric through the host-buffer C ABI: rvt_gene_push_bed (PLINK 2-bit rows, the form the
    reference keeps large cohorts in) or rvt_gene_push_i8, pinned host memory -> rvt_flush (records
    copied back).  H2D of every gene and D2H of the records are inside the timed region."""
    import rvtests_b200
    from rvtests_b200.synth import pack_bed
    ng, M, N = (args.e2e_genes if fmt in ("bed", "bed_missing", "bed_binary") else max(16, args.e2e_genes // 4)), args.variants, args.samples
    if fmt == "f64":
        ng = max(4, args.e2e_genes // 32)               # 200 MB per gene: the reference's own Matrix boundary
    ng = min(ng, args.genes)
    nd = min(ng, {"f64": 4, "bed_missing": 16}.get(fmt, 64))   # distinct genes held on the host, cycled
    calls = eng.loaded_read(0, nd * M)                  # the same genotypes as the first resident genes
    af = 0.5 * calls.reshape(nd, M, N).sum(axis=2, dtype=np.int64) / N
    src_eng = eng
    if fmt == "bed_binary":
        # a binary trait (logistic null on the device, p(1-p)-weighted statistics through the fp64 path, enqueued behind the
        # copies: option binary_stream) on an engine of its own, so that the resident cohort keeps its quantitative null
        from rvtests_b200.synth import covariates
        X, _y = covariates(SEED, N, args.covariates)
        yb = (np.random.default_rng(SEED).random(N) < 0.3).astype(np.float64)
        eng = rvtests_b200.GeneEngine(torch.cuda.current_device())
        eng.set_stream(torch.cuda.current_stream().cuda_stream)
        eng.set_null_model(X, yb, binary=True)
    if fmt in ("bed", "bed_missing", "bed_binary"):
        host = torch.empty((nd * M, (N + 3) // 4), dtype=torch.uint8, pin_memory=True)
        if fmt in ("bed", "bed_binary"):
            host.numpy()[:] = pack_bed(calls)
        else:   # 1 % of the calls missing (PLINK code 01): mean-imputed on the device, augmented tensor-core sweep
            rng = np.random.default_rng(SEED + rank)
            for b0 in range(0, nd * M, M):
                host.numpy()[b0:b0 + M] = pack_bed(calls[b0:b0 + M], rng.random((M, N)) < 0.01)
            af = None   # (the engine derives the frequencies from the observed calls)
    elif fmt == "i8":
        host = torch.empty((nd * M, N), dtype=torch.int8, pin_memory=True)
        host.numpy()[:] = calls
    else:   # dc->getGenotype(): N x M column-major doubles == (M, N) row-major
        host = torch.empty((nd * M, N), dtype=torch.float64, pin_memory=True)
        host.numpy()[:] = calls
    del calls
    hn = host.numpy()

    def step():
        for g in range(ng):
            k = g % nd
            blk = hn[k * M:(k + 1) * M]
            if fmt == "bed_missing":
                eng.push_bed(blk, None)
            elif fmt in ("bed", "bed_binary"):
                eng.push_bed(blk, af[k])
            elif fmt == "i8":
                eng.push_i8(blk, af[k])
            else:
                eng.push_f64(blk.T, af[k])
        return eng.flush()

    # kernels of every 64 pushed genes are enqueued at once and run under the following H2D copies
    eng.set_option("stream_batch", 64)
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    reps = max(2, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(reps):
        r = step()
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    sec = float(dt.item()) / reps
    eng.set_option("stream_batch", 0)
    assert int((r["status"] == 0).sum()) == ng
    if eng is not src_eng:
        eng.close()
    return {"value": world * ng / sec, "unit": "gene-sets/s",
            "h2d_bytes_per_step": int(world * ng * M * hn.shape[1] * hn.itemsize), "d2h_bytes_per_step": int(world * ng * rvtests_b200.engine.RESULT_DTYPE.itemsize),
            "genes_per_step_per_gpu": ng,
            "host_format": {"bed": "PLINK .bed 2-bit SNP-major rows, pinned host memory (rvt_gene_push_bed)",
                            "bed_missing": "PLINK .bed 2-bit rows with 1 % of the calls missing (mean-imputed on the device: augmented "
                                           "tensor-core sweep), pinned host memory (rvt_gene_push_bed)",
                            "bed_binary": "PLINK .bed 2-bit rows, BINARY trait (logistic null; every gene takes the p(1-p)-weighted fp64 "
                                          "statistics, enqueued behind the copies), pinned host memory (rvt_gene_push_bed)",
                            "i8": "int8 variant-major hard calls, pinned host memory (rvt_gene_push_i8)",
                            "f64": "N x M column-major doubles = dc->getGenotype(), the ModelFitter::fit boundary itself, "
                                   "pinned host memory (rvt_gene_push_f64)"}[fmt],
            "pcie_gbs": world * ng * M * hn.shape[1] * hn.itemsize / sec / 1e9,
            "timing": "host wall clock around push+flush (copies inside), max over ranks"}


# ---------------------------------------------------------------------------------------------------
def cpu_problem(args):
    from oracle import oracle as O
    try:
        O.build(native=True)
        native = True
    except Exception:
        native = False
    from rvtests_b200.synth import variant_params, covariates
    nd, M, N = args.cpu_genes, args.variants, args.samples
    keys, t0, t1 = variant_params(SEED, 0, nd * M)
    G = O.synth_rows_f64(keys, t0, t1, N, threads=0, native=native).reshape(nd, M, N)
    X, y = covariates(SEED, N, args.covariates)
    nm = O.fit_null_linear(X, y)
    af = 0.5 * G.sum(axis=2) / N
    return O, native, G, af, X, nm


def time_cpu(O, native, G, af, X, nm, threads, seconds):
    nd = G.shape[0]
    # calibrate on one pass of `threads` tasks, then size the sample for ~`seconds` of work
    idx = np.arange(max(threads, 1)) % nd
    t = time.perf_counter()
    O.gene_batch_idx(G, af, idx, X, nm["resid"], nm["sigma2"], threads=threads, native=native)
    one = time.perf_counter() - t
    reps = int(max(1, min(50, seconds / max(one, 1e-3))))
    idx = np.arange(len(idx) * reps) % nd
    t = time.perf_counter()
    out = O.gene_batch_idx(G, af, idx, X, nm["resid"], nm["sigma2"], threads=threads, native=native)
    dt = time.perf_counter() - t
    return len(idx) / dt, len(idx), dt, out


def reference_wall(O, M, C, sizes=(2000, 4000, 8000)):
    """the reference's OWN Skat::Fit (regression/Skat.cpp:29-105, compiled unmodified into
    oracle/_ref/libskat_ref.so against oracle/eigen_standin) forms the N x N float matrix P0 and costs 2 M N^2 flops
    per gene: timed at small N to show the O(N^2) wall and why it cannot run at the metric's N.  The stand-in's
    products are plain loops (Eigen's blocked GEMM would be a few times faster): the exponent and the 4 N^2 bytes of
    P0 are the point, not the constant."""
    if O.ref_skat() is None:
        return {"unavailable": "oracle/_ref/libskat_ref.so not built"}
    rng = np.random.default_rng(SEED)
    rows = []
    for N in sizes:
        maf = rng.uniform(0.005, 0.05, M)
        G = np.asfortranarray((rng.random((N, M)) < maf).astype(np.float64) + (rng.random((N, M)) < maf))
        X = np.c_[np.ones(N), rng.normal(size=(N, C - 1))]
        nm = O.fit_null_linear(X, rng.normal(size=N))
        af = G.mean(axis=0) / 2
        w = np.array([O.lib().orc_skat_weight(float(a), 1.0, 25.0, 1) for a in af])
        t = time.perf_counter()
        r = O.ref_skat_fit(nm["resid"], np.full(N, nm["sigma2"]), X, G, w)
        dt = time.perf_counter() - t
        rows.append({"N": N, "s_per_gene": dt, "p0_bytes": 4 * N * N, "Q": r["Q"]})
    slope = float(np.polyfit(np.log([r_["N"] for r_ in rows]), np.log([r_["s_per_gene"] for r_ in rows]), 1)[0])
    last = rows[-1]
    return {"what": "Skat::Fit of the reference build (float32, explicit N x N P0), 1 thread, M=%d, C=%d" % (M, C),
            "rows": rows, "fitted_exponent_in_N": slope,
            "extrapolated_s_per_gene_at_500k": last["s_per_gene"] * (500_000 / last["N"]) ** 2,
            "p0_bytes_at_500k": 4 * 500_000 ** 2}


def cpu_baseline(args, threads):
    O, native, G, af, X, nm = cpu_problem(args)
    try:
        wall = reference_wall(O, args.variants, args.covariates)
    except Exception as e:  # a reported side figure must never cost the bench line
        wall = {"unavailable": repr(e)}
    v, ntask, dt, _ = time_cpu(O, native, G, af, X, nm, threads, args.cpu_seconds)
    v1, ntask1, dt1, _ = time_cpu(O, native, G, af, X, nm, 1, min(args.cpu_seconds, 6.0))
    return {"value": v, "unit": "gene-sets/s", "cores": threads, "kind": "port",
            "sample": f"{ntask} gene-tasks over {G.shape[0]} distinct genes of N={args.samples} x M={args.variants} "
                      f"(fp64 column-major Matrix, {args.samples * args.variants * 8 / 1e6:.0f} MB each) in {dt:.1f} s, OpenMP over genes, "
                      f"oracle/skat_oracle.c {'-march=native' if native else 'x86-64-v3'}: reduced O(N M^2) algebra "
                      "(the reference's own Skat.cpp is O(N^2) and cannot run at this N: SURVEY.md F2)",
            "single_thread_value": v1, "single_thread_note": "the reference's gene loop is serial (src/Main.cpp:1221-1254)",
            "reference_algorithm_wall": wall}


def run_reference(args):
    """reference arm: the reference algorithm on all host cores; rank 0 only.  The oracle port (reduced O(N M^2)
    algebra) is timed: the reference's own Skat::Fit is built (oracle/_ref/libskat_ref.so) but forms an N x N matrix
    -- 1 TB at N = 500 000 -- see cpu_baseline.reference_algorithm_wall of the main arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count()
    O, native, G, af, X, nm = cpu_problem(args)
    per_step = []
    for k in range(args.warmup + args.steps):
        v, ntask, dt, _ = time_cpu(O, native, G, af, X, nm, threads, max(2.0, args.cpu_seconds / max(args.steps, 1)))
        if k >= args.warmup:
            per_step.append((v, ntask, dt))
    value = float(np.mean([p[0] for p in per_step]))
    ms = float(np.mean([p[2] for p in per_step])) * 1e3
    sample = (f"each step = {per_step[0][1]} gene-tasks over {G.shape[0]} distinct genes of N={args.samples} x "
              f"M={args.variants}, OpenMP over genes on {threads} threads, oracle port "
              f"({'-march=native' if native else 'x86-64-v3'})")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "gene-sets/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": "gene-sets/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "gene-sets/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ---------------------------------------------------------------------------------------------------
META_METRIC = "variants/sec (--meta score,cov, 500k samples, MetaCov window 1 Mb)"


def meta_config(args):
    return {"workload": f"--meta score,cov: N={args.samples} samples, {args.meta_variants} variants per GPU and step, one variant "
                        f"every {args.meta_spacing} bp, window 1 Mb (BASELINE configs[3] in blocks + one-window halo), C={args.covariates}",
            "samples": args.samples, "variants_per_gpu": args.meta_variants, "spacing_bp": args.meta_spacing,
            "window_bp": 1_000_000, "covariates_incl_intercept": args.covariates,
            "l2_policy": "the covariance band is walked sample-chunk by sample-chunk so that a window's tiles are re-used from "
                         "L2 BY DESIGN (that reuse is the algorithm); every step starts from HBM: 4.6 GB of genotypes per step "
                         ">> 126 MB L2"}


def run_meta(args):
    """--meta score,cov: per rank a block of consecutive variants plus a one-window halo (the variants of the next block
    that the last variants of this one pair with), all resident as 64-variant tiles; a step = score statistics + exact HWE
    of every variant + the covariance band of the block; N > 1: one NCCL all_gather of the bands."""
    import torch
    import torch.distributed as dist
    import rvtests_b200
    from rvtests_b200.synth import variant_params, covariates

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    numa = bind_numa(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    stream = torch.cuda.current_stream()
    N, nv, window = args.samples, args.meta_variants, 1_000_000
    halo = ((window // args.meta_spacing + 63) // 64) * 64          # variants of the next block inside the last window
    nall = nv + halo
    keys, t0, t1 = variant_params(SEED, rank * nv, nall)            # (the halo IS the next rank's first variants)
    X, y = covariates(SEED, N, args.covariates)
    eng = rvtests_b200.GeneEngine(local)
    eng.set_stream(stream.cuda_stream)
    eng.set_null_model(X, y)
    eng.synth_load(keys, t0, t1, nall // 64, 64)
    pos = (args.meta_spacing * (rank * nv + np.arange(nall))).astype(np.int32)
    chrom = np.ones(nall, dtype=np.int32)
    eng.push_loaded()
    _, _, wmax = eng.meta_flush(nall, pos, chrom, window)             # plan + first-use allocations (host outputs)
    rec = rvtests_b200.engine.VARIANT_DTYPE.itemsize
    d_v = torch.empty(nall * rec, dtype=torch.uint8, device=dev)
    d_band = torch.empty(nall * (wmax + 1), dtype=torch.float64, device=dev)
    d_all = torch.empty(world * nv * (wmax + 1), dtype=torch.float64, device=dev) if world > 1 else None

    def step():
        eng.push_loaded()
        eng.meta_flush_dev(nall, pos, chrom, window, d_v.data_ptr(), d_band.data_ptr(), d_band.numel())
        if world > 1:
            dist.all_gather_into_tensor(d_all, d_band[: nv * (wmax + 1)])

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    t_w, nw = time.perf_counter(), 0
    while nw < max(args.warmup, 3) or time.perf_counter() - t_w < 0.5:
        step()
        nw += 1
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pair_ms, diag_ms, pairs = [], [], 0
    e0.record(stream)
    for _ in range(args.steps):
        step()
        t = eng.last_timing()
        pair_ms.append(t["sweep_ms"])
        diag_ms.append(t["finalize_ms"])
        pairs = t["launches"]
    e1.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop() if rank == 0 else None
    tmax = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms = float(tmax.item()) / args.steps
    # ---- e2e: 2-bit rows from pinned host memory in, records + band to pinned host memory out
    from rvtests_b200.synth import pack_bed
    calls = eng.loaded_read(0, nall)
    host = torch.empty((nall, (N + 3) // 4), dtype=torch.uint8, pin_memory=True)
    host.numpy()[:] = pack_bed(calls)
    hn = host.numpy()
    h_v = torch.empty(nall * rec, dtype=torch.uint8, pin_memory=True)
    h_band = torch.empty(nall * (wmax + 1), dtype=torch.float64, pin_memory=True)

    def e2e_step():
        for b0 in range(0, nall, 64):
            eng.push_bed(hn[b0:b0 + 64], None)
        eng.meta_flush_dev(nall, pos, chrom, window, h_v.data_ptr(), h_band.data_ptr(), h_band.numel())

    e2e_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    reps = 2
    tw = time.perf_counter()
    for _ in range(reps):
        e2e_step()
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - tw], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    e2e_s = float(dt.item()) / reps
    if rank == 0:
        vout = np.frombuffer(d_v.cpu().numpy().tobytes(), dtype=rvtests_b200.engine.VARIANT_DTYPE)
        band = d_band.cpu().numpy().reshape(nall, wmax + 1)
        pk = peaks()
        pair_s = float(np.mean(pair_ms)) * 1e-3
        ops = 2.0 * 64 * 64 * N * pairs                         # every tile pair is a 64 x 64 x N int8 GEMM
        tens_peak = (pk.get("bf16_tflops_sustained") or pk.get("bf16_tflops")) if pk else 1590.0
        operand_bytes = 2.0 * 64 * N * pairs
        alg_bytes = float(nall) * N                             # every genotype byte once
        out = {
            "metric": META_METRIC, "value": world * nv / (ms * 1e-3), "unit": "variants/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "s8 x s8 -> s32 (exact) + f64 tail", "data": "synthetic", "config": meta_config(args),
            "kernel_ms_per_step": {"tile pairs (k_sweep_tc PAIR) + band assembly": float(np.mean(pair_ms)),
                                   "diagonal tiles + score statistics + exact HWE": float(np.mean(diag_ms))},
            "tile_pairs_per_step": int(pairs), "cov_entries_per_step": int(np.sum(~np.isnan(band[:nv]))),
            "gpu_launches": int(args.steps * (2 * ((nall // 64 + 1023) // 1024) + 1 + 2 * ((pairs + 1023) // 1024))),
            "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": "k_sweep_tc (PAIR mode)", "achieved": ops / pair_s / 1e12, "peak": tens_peak,
                         "unit": "TFLOP/s", "frac": ops / pair_s / 1e12 / tens_peak,
                         "peak_source": ("MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if pk else "fallback (of fallback)"),
                         "note": "s8 tcgen05 operations counted as flops against the dense bf16 peak (the int8 pipe's own peak is "
                                 "twice that); 64 x 64 x 32 UMMAs, whose issue cost (~76 clk) is the limit of this tiling",
                         "algorithmic_flops_per_launch": ops,
                         "operand_tb_per_s": operand_bytes / pair_s / 1e12,
                         "operand_note": "bytes of A and B tiles fed to the tensor core per second; above the HBM copy peak "
                                         "means the band is served from L2 (split-major walk of the pair units)",
                         "traffic": None},
            "engine": {"splits": int(eng.info("last_splits")), "numa": numa,
                       "parallelism": f"variant blocks + one-window halo over {world} rank(s), one NCCL all_gather of the band"},
            "sanity": {"variants_ok": int(vout["ok"][:nv].sum()), "median_p": float(np.median(vout["pvalue"][:nv][vout["ok"][:nv] == 1])),
                       "median_hwe_p": float(np.median(vout["hwe_p"][:nv]))},
            "e2e": {"value": world * nv / e2e_s, "unit": "variants/s",
                    "h2d_bytes_per_step": int(world * nall * hn.shape[1]),
                    "d2h_bytes_per_step": int(world * (nall * rec + nall * (wmax + 1) * 8)),
                    "host_format": "PLINK .bed 2-bit rows in pinned host memory -> rvt_gene_push_bed; records and the covariance band "
                                   "copied back to pinned host memory", "timing": "host wall clock, max over ranks"},
        }
        if not args.no_cpu and world == 1:
            out["parity"], out["cpu_baseline"] = meta_cpu_leg(args, eng, vout, band, pos, chrom, window, X, y)
        print(json.dumps(out))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def meta_cpu_leg(args, eng, vout, band, pos, chrom, window, X, y, n_check=96):
    """oracle/meta_oracle.py (numpy restatement of MetaScoreTest / MetaCovTest, pinned on the reference model layer) on the
    first n_check variants with a window cut to them: checker of the records of the timed steps AND the timed CPU baseline."""
    from oracle import oracle as O
    from oracle import meta_oracle as MO
    O.build()
    N = args.samples
    G = eng.loaded_read(0, n_check).T.astype(np.float64)          # (N, n_check)
    nm = O.fit_null_linear(X, y)
    t = time.perf_counter()
    sc = [MO.meta_score(G[:, j], X, nm["resid"], nm["sigma2"]) for j in range(n_check)]
    cv = MO.meta_cov(G, pos[:n_check], chrom[:n_check], X, nm["sigma2"], window)
    dt = time.perf_counter() - t

    def rel(a, b):
        a, b = float(a), float(b)
        return 0.0 if a == b else abs(a - b) / max(abs(a), abs(b), 1e-300)

    worst = {"U": 0.0, "sqrtV": 0.0, "pvalue": 0.0, "hwe_p": 0.0, "cov": 0.0}
    exact = True
    for j in range(n_check):
        r, o = vout[j], sc[j]
        exact &= (int(r["n_ref"]), int(r["n_het"]), int(r["n_alt"])) == (o["n_ref"], o["n_het"], o["n_alt"])
        worst["hwe_p"] = max(worst["hwe_p"], rel(r["hwe_p"], o["hwe_p"]))
        if o.get("ok"):
            for k in ("U", "sqrtV", "pvalue"):
                worst[k] = max(worst[k], rel(r[k], o[k]))
        if cv[j] is not None:
            for pj, val in zip(*cv[j]):
                d = (pj - int(pos[j])) // args.meta_spacing
                worst["cov"] = max(worst["cov"], abs(band[j, d] - val) / max(abs(val), abs(cv[j][1][0]) * 1e-6, 1e-300))
    ok = exact and worst["U"] <= 1e-6 and worst["sqrtV"] <= 1e-6 and worst["pvalue"] <= 1e-4 and worst["hwe_p"] <= 1e-9 and worst["cov"] <= 1e-5
    parity = {"variants_checked": n_check, "pass": bool(ok), "max_rel_err": worst, "counts_exact": bool(exact),
              "tolerances": {"U, sqrtV": 1e-6, "pvalue": 1e-4, "hwe_p": 1e-9, "cov (relative to the variant's variance)": 1e-5},
              "checker": "oracle/meta_oracle.py (pinned on the reference's MetaScoreTest / MetaCovTest text)"}
    cpu = {"value": n_check / dt, "unit": "variants/s", "cores": 1, "kind": "port",
           "sample": f"{n_check} variants of N={N} with all their pairs inside the sample ({n_check * (n_check + 1) // 2} covariances) in {dt:.1f} s: "
                     "numpy restatement of MetaScoreTest + MetaCovTest (oracle/meta_oracle.py), single thread; the full window "
                     f"({window // args.meta_spacing} partners per variant) would cost ~{window // args.meta_spacing / (n_check / 2):.0f}x more per variant"}
    return parity, cpu


def run_reference_meta(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    from oracle import meta_oracle as MO
    O.build()
    from rvtests_b200.synth import variant_params, covariates
    N, n_s = args.samples, 64
    keys, t0, t1 = variant_params(SEED, 0, n_s)
    G = O.synth_rows_f64(keys, t0, t1, N, threads=0).reshape(n_s, N).T.copy()
    X, y = covariates(SEED, N, args.covariates)
    nm = O.fit_null_linear(X, y)
    pos = (args.meta_spacing * np.arange(n_s)).astype(np.int32)
    chrom = np.ones(n_s, dtype=np.int32)
    vals = []
    for k in range(args.warmup + args.steps):
        t = time.perf_counter()
        [MO.meta_score(G[:, j], X, nm["resid"], nm["sigma2"]) for j in range(n_s)]
        MO.meta_cov(G, pos, chrom, X, nm["sigma2"], 1_000_000)
        dt = time.perf_counter() - t
        if k >= args.warmup:
            vals.append((n_s / dt, dt))
    value = float(np.mean([v[0] for v in vals]))
    sample = (f"each step = {n_s} variants of N={N} with the {n_s * (n_s + 1) // 2} covariances among them "
              f"(the full 1 Mb window has {1_000_000 // args.meta_spacing} partners per variant, ~{1_000_000 // args.meta_spacing / (n_s / 2):.0f}x the work), "
              "oracle/meta_oracle.py (numpy), single thread")
    print(json.dumps({
        "impl": "reference", "metric": META_METRIC, "value": value, "unit": "variants/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": float(np.mean([v[1] for v in vals])) * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": meta_config(args),
        "cpu_baseline": {"value": value, "unit": "variants/s", "cores": 1, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "variants/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


if __name__ == "__main__":
    a = parse()
    if a.workload == "bolt":
        from tools import bolt_bench
        if "--steps" not in sys.argv:
            a.steps = 2                       # a step is a whole null fit
        if "--warmup" not in sys.argv:
            a.warmup = 1
        bolt_bench.run_reference_bolt(a) if a.impl == "reference" else bolt_bench.run_bolt(a, ClockSampler, bind_numa)
    elif a.impl == "reference":
        run_reference_meta(a) if a.workload == "meta" else run_reference(a)
    elif a.workload == "meta":
        run_meta(a)
    else:
        run_ours(a)
