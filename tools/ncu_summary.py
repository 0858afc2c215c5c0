#!/usr/bin/env python
"""Turn one gpurun evidence run (tools/gpu_profile.sh) into the tracked files under profiles/.

  python tools/ncu_summary.py r01c            # reads gpurun_out/, writes profiles/r01c_*

Inputs (all written by tools/gpu_profile.sh on the GPU box):
  gpurun_out/prof.ncu-rep    ncu --set full of one k_sweep_tc + one k_finalize launch (512 genes)
  gpurun_out/launches.csv    ncu --metrics gpu__time_duration.sum launch list of bench.py
  gpurun_out/bench_full.json, bench_ref.json, phases.log, ablation.log
"""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")


def raw_page(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units, body = rows[0], rows[1], rows[2:]
    return txt, [{h: (r[i], units[i]) for i, h in enumerate(hdr)} for r in body]


def fnum(cell):
    v, u = cell
    v = float(v.replace(",", ""))
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Tbyte": 1e12, "ms": 1, "us": 1e-3, "ns": 1e-6, "s": 1e3}
    return v * scale.get(u, 1)


def main():
    tag = sys.argv[1]
    genes = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    N, M = 500000, 50
    txt, kernels = raw_page(os.path.join(OUT, "prof.ncu-rep"))
    with open(os.path.join(PROF, f"{tag}_ncu_raw_sweep_finalize.csv"), "w") as f:
        f.write(txt)
    summ = {"capture": "ncu --set full --clock-control none --import-source on -k regex:k_sweep_tc|k_finalize -s 2 -c 2 "
                       f"python bench.py --genes {genes} --steps 1 --warmup 3 --no-cpu --no-e2e  (B200)"}
    for k in kernels:
        name = k["Kernel Name"][0]
        short = "k_sweep_tc" if "k_sweep_tc" in name else "k_finalize" if "k_finalize" in name else name[:40]
        d = {
            "kernel": name[:80],
            "genes_in_launch": genes,
            "duration_ms_under_ncu": fnum(k["gpu__time_duration.sum"]),
            "dram_bytes_read": fnum(k["dram__bytes_read.sum"]),
            "dram_bytes_write": fnum(k["dram__bytes_write.sum"]),
            "dram_read_pct_of_ncu_peak": float(k["dram__bytes_read.sum.pct_of_peak_sustained_elapsed"][0]),
            "issue_active_pct": float(k["smsp__issue_active.avg.pct_of_peak_sustained_active"][0]),
            "sm_throughput_pct": float(k["sm__throughput.avg.pct_of_peak_sustained_elapsed"][0]),
            "registers_per_thread": int(k["launch__registers_per_thread"][0]),
            "grid": int(k["launch__grid_size"][0]),
            "block": int(k["launch__block_size"][0]),
            "ctas_per_sm_limit": {"registers": float(k["launch__occupancy_limit_registers"][0]),
                                  "shared_mem": float(k["launch__occupancy_limit_shared_mem"][0])},
        }
        if short == "k_sweep_tc":
            d["algorithmic_bytes"] = genes * N * M
            d["traffic_over_algorithmic"] = (d["dram_bytes_read"] + d["dram_bytes_write"]) / d["algorithmic_bytes"]
            d["achieved_gbs_under_ncu"] = d["algorithmic_bytes"] / d["duration_ms_under_ncu"] / 1e6
        summ[short] = d
    # launch-list shares of the bench step
    lpath = os.path.join(OUT, "launches.csv")
    if os.path.exists(lpath):
        shutil.copy(lpath, os.path.join(PROF, f"{tag}_launches_bench.csv"))
        tot = {}
        with open(lpath) as f:
            lines = [l for l in f if l.startswith('"')]
        for r in csv.DictReader(lines):
            nm = r["Kernel Name"]
            key = "k_sweep_tc" if "k_sweep_tc" in nm else "k_finalize" if "k_finalize" in nm else "other (setup: null model, synthetic genotypes, flags)"
            tot[key] = tot.get(key, 0.0) + float(r["Metric Value"])
        hot = tot.get("k_sweep_tc", 0) + tot.get("k_finalize", 0)
        summ["launch_list_shares_of_step"] = {k: v / hot for k, v in tot.items() if k.startswith("k_")}
        summ["launch_list_total_ns"] = tot
    for src, dst in (("bench_full.json", f"{tag}_bench_n1.json"), ("bench_ref.json", f"{tag}_bench_reference_arm.json"),
                     ("phases.log", f"{tag}_finalize_phases.txt"), ("ablation.log", f"{tag}_sweep_ablation.txt"),
                     ("pytest_gpu.log", f"{tag}_pytest_gpu.txt")):
        p = os.path.join(OUT, src)
        if os.path.exists(p):
            with open(p) as f:
                body = f.read()
            if src.endswith(".json"):
                body = "\n".join(l for l in body.splitlines() if l.startswith("{")) + "\n"
                try:
                    b = json.loads(body)
                    if "kernel_ms_per_step" in b:
                        km = b["kernel_ms_per_step"]
                        s = sum(km.values())
                        summ["bench_cuda_event_shares_of_step"] = {k: v / s for k, v in km.items()}
                except Exception:
                    pass
            with open(os.path.join(PROF, dst), "w") as f:
                f.write(body)
    with open(os.path.join(PROF, f"{tag}_ncu_summary.json"), "w") as f:
        json.dump(summ, f, indent=1)
    print(json.dumps(summ, indent=1))


if __name__ == "__main__":
    main()
