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
KEYS += ["sm__icc_request_hit_rate.pct", "sm__icc_requests.sum",
         "gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed"]  # instruction caches: SM, L1.5
KEYS += [f"smsp__average_warps_issue_stalled_{r}_per_issue_active.ratio" for r in
         ("barrier", "wait", "no_instruction", "branch_resolving", "long_scoreboard", "short_scoreboard",
          "math_pipe_throttle", "not_selected")]


def sym(pattern):
    """mangled name of the first kernel in the library whose name contains every piece of `pattern`"""
    lib = os.path.join(ROOT, "collision_avoidance_b200", "liborca_b200.so")
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    for ln in out.splitlines():
        if "Function :" in ln and all(p in ln for p in pattern):
            return ln.split("Function :")[1].strip()
    raise SystemExit(f"no kernel matches {pattern}")


scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
# config -> (report suffix, kernel symbol pieces, algorithmic bytes per launch)
CAPTURES = {
    "cfg2": ("step", ("step_small_kernelILi10ELb1ELi1ELi2E",), 41 * 65536 * 16),
    "cfg3": ("cfg3", ("step_small_kernelILi10ELb1ELi3ELi2E",), 82 * 32768 * 32),
    "cfg4": ("cfg4", ("step_small_kernelILi10ELb1ELi1ELi6E",), 48 * 2048 * 256),
    "cfg5": ("grid", ("step_grid_kernelILi10ELb1ELi1E",), 44 * 1_000_000),
}
traffic = {}
for cfg, (suffix, pieces, alg_bytes) in CAPTURES.items():
    rep = os.path.join(G, f"prof_{tag}_{suffix}.ncu-rep")
    if not os.path.exists(rep):
        continue
    kernel = sym(pieces)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, u = rows[0], rows[1]
    with open(os.path.join(P, f"{tag}_{cfg}_ncu_raw.csv"), "w") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [f"launch{i}" for i in range(len(rows) - 2)])
        w.writerow(["kernel", ""] + [r[h.index("Kernel Name")] for r in rows[2:]])
        for k in KEYS:
            if k in h:
                i = h.index(k)
                w.writerow([k, u[i]] + [r[i] for r in rows[2:]])

    def val(r, k):
        return float(r[h.index(k)].replace(",", ""))

    def nbytes(r, k):
        return val(r, k) * scale[u[h.index(k)]]

    n = len(rows) - 2
    traffic[cfg] = {
        "kernel": rows[2][h.index("Kernel Name")],
        "dram_bytes_per_launch": sum(nbytes(r, "dram__bytes_read.sum") + nbytes(r, "dram__bytes_write.sum") for r in rows[2:]) / n,
        "algorithmic_bytes_per_launch": alg_bytes,
        "kernel_us_under_ncu": sum(val(r, "gpu__time_duration.sum") for r in rows[2:]) / n,
        "issue_slots_busy_frac": sum(val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active") for r in rows[2:]) / n / 100.0,
        "warp_instructions_per_launch": sum(val(r, "smsp__inst_executed.sum") for r in rows[2:]) / n,
        "active_lanes_per_instruction": sum(val(r, "smsp__thread_inst_executed_per_inst_executed.ratio") for r in rows[2:]) / n,
        "episode_step": {"cfg2": 160, "cfg3": 160, "cfg4": 90, "cfg5": 90}[cfg],
        "source": f"profiles/{tag}_{cfg}_ncu_raw.csv (ncu --set full --clock-control none on bench.py --config {cfg}, {n} launch(es))",
    }
    with open(os.path.join(P, f"{tag}_{cfg}_by_function.txt"), "w") as f:
        subprocess.call([sys.executable, os.path.join(ROOT, "tools", "ncu_regions.py"), rep, kernel], stdout=f)
    if cfg == "cfg2":
        with open(os.path.join(P, f"{tag}_{cfg}_by_callpath.txt"), "w") as f:  # same, keyed by inline call path
            subprocess.call([sys.executable, os.path.join(ROOT, "tools", "ncu_phases.py"), rep, kernel, "3"], stdout=f)
        with open(os.path.join(P, f"{tag}_{cfg}_hot_code.txt"), "w") as f:  # instruction-cache working set
            subprocess.call([sys.executable, os.path.join(ROOT, "tools", "ncu_hot_code.py"), rep, kernel, "2"], stdout=f)
json.dump(traffic, open(os.path.join(P, "traffic.json"), "w"), indent=1)
for name in (f"launches_{tag}.csv", f"launches_{tag}_grid.csv", f"configs_{tag}.jsonl"):
    if os.path.exists(os.path.join(G, name)):
        shutil.copy(os.path.join(G, name), os.path.join(P, f"{tag}_" + name.replace(f"_{tag}", "")))
print(json.dumps(traffic, indent=1))
