#!/usr/bin/env python
"""Copies the round's evidence from gpurun_out/ into profiles/ and derives the summaries
(traffic.json, raw-metric CSV, per-function breakdown).  usage: tools/make_profiles.py r01"""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
KERNEL = "_ZN4orca17step_small_kernelILi10ELb1ELi1EEEvNS_8StepArgsE"
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__waves_per_multiprocessor", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "launch__grid_size", "launch__block_size", "sm__cycles_elapsed.avg"]

shutil.copy(os.path.join(G, f"launches_{tag}.csv"), os.path.join(P, f"{tag}_launches.csv"))
shutil.copy(os.path.join(G, f"bench_{tag}.json"), os.path.join(P, f"{tag}_bench_n1.json"))
shutil.copy(os.path.join(G, f"bench_{tag}_ref.json"), os.path.join(P, f"{tag}_bench_reference_arm.json"))
if os.path.exists(os.path.join(G, f"configs_{tag}.jsonl")):
    shutil.copy(os.path.join(G, f"configs_{tag}.jsonl"), os.path.join(P, f"{tag}_configs_1gpu.jsonl"))
rep = os.path.join(G, f"prof_{tag}_step.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, u = rows[0], rows[1]
with open(os.path.join(P, f"{tag}_step_small_ncu_raw.csv"), "w") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit"] + [f"launch{i}" for i in range(len(rows) - 2)])
    for k in KEYS:
        if k in h:
            i = h.index(k)
            w.writerow([k, u[i]] + [r[i] for r in rows[2:]])
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def nbytes(r, k):
    return float(r[h.index(k)].replace(",", "")) * scale[u[h.index(k)]]


tr = [nbytes(r, "dram__bytes_read.sum") + nbytes(r, "dram__bytes_write.sum") for r in rows[2:]]
json.dump({"step_small_kernel_dram_bytes_per_launch": sum(tr) / len(tr),
           "source": f"profiles/{tag}_step_small_ncu_raw.csv (ncu --set full on bench.py, step ~100, {len(tr)} launches)",
           "algorithmic_bytes_per_launch": 41 * 65536 * 16}, open(os.path.join(P, "traffic.json"), "w"), indent=1)
with open(os.path.join(P, f"{tag}_step_small_by_function.txt"), "w") as f:
    subprocess.call([sys.executable, os.path.join(ROOT, "tools", "ncu_regions.py"), rep, KERNEL], stdout=f)
print(open(os.path.join(P, "traffic.json")).read())
