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
def fnum(r, k):
    return float(r[h.index(k)].replace(",", "")) if k in h else None


issue = [fnum(r, "smsp__issue_active.avg.pct_of_peak_sustained_active") for r in rows[2:]]
winst = [fnum(r, "smsp__inst_executed.sum") for r in rows[2:]]
lanes = [fnum(r, "smsp__thread_inst_executed_per_inst_executed.ratio") for r in rows[2:]]
json.dump({"step_small_kernel_dram_bytes_per_launch": sum(tr) / len(tr),
           "step_small_kernel_issue_active_pct": sum(issue) / len(issue),
           "step_small_kernel_warp_instructions_per_launch": sum(winst) / len(winst),
           "step_small_kernel_active_lanes_per_instruction": sum(lanes) / len(lanes),
           "source": f"profiles/{tag}_step_small_ncu_raw.csv (ncu --set full on bench.py, step ~100, {len(tr)} launches)",
           "algorithmic_bytes_per_launch": 41 * 65536 * 16}, open(os.path.join(P, "traffic.json"), "w"), indent=1)
with open(os.path.join(P, f"{tag}_step_small_by_function.txt"), "w") as f:
    subprocess.call([sys.executable, os.path.join(ROOT, "tools", "ncu_regions.py"), rep, KERNEL], stdout=f)
with open(os.path.join(P, f"{tag}_step_small_by_callpath.txt"), "w") as f:  # same, keyed by inline call path
    subprocess.call([sys.executable, os.path.join(ROOT, "tools", "ncu_phases.py"), rep, KERNEL, "3"], stdout=f)
with open(os.path.join(P, f"{tag}_step_small_hot_code.txt"), "w") as f:  # instruction-cache working set
    subprocess.call([sys.executable, os.path.join(ROOT, "tools", "ncu_hot_code.py"), rep, KERNEL, "2"], stdout=f)
for n in (2, 4, 8):
    for arm, suffix in (("", ""), ("_ref", "_reference_arm")):
        src = os.path.join(G, f"bench_{tag}_n{n}{arm}.json")
        if os.path.exists(src):
            shutil.copy(src, os.path.join(P, f"{tag}_bench_n{n}{suffix}.json"))
# other kernels of the round: raw metrics + per-function breakdown + tensor-core SASS evidence
EXTRA = {"policy_tc": ("_ZN4orca20policy_mlp_tc_kernelENS_7MlpArgsE", "policy_mlp_tc_kernel"),
         "obs": ("_ZN4orca14observe_kernelENS_7ObsArgsEi", "observe_kernel"),
         "grid": ("_ZN4orca16step_grid_kernelILi10ELb1ELi1EEEvNS_8StepArgsEPK6float2S4_PKiS6_PKNS_10GridParamsE",
                  "step_grid_kernel")}
for short, (mangled, nice) in EXTRA.items():
    rep2 = os.path.join(G, f"prof_{tag}_{short}.ncu-rep")
    if not os.path.exists(rep2):
        continue
    raw2 = subprocess.run(["ncu", "-i", rep2, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows2 = list(csv.reader(raw2.splitlines()))
    h2, u2 = rows2[0], rows2[1]
    with open(os.path.join(P, f"{tag}_{short}_ncu_raw.csv"), "w") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [f"launch{i}" for i in range(len(rows2) - 2)])
        for k in KEYS + ["sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
                         "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
                         "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"]:
            if k in h2:
                i = h2.index(k)
                w.writerow([k, u2[i]] + [r[i] for r in rows2[2:]])
    with open(os.path.join(P, f"{tag}_{short}_by_function.txt"), "w") as f:
        subprocess.call([sys.executable, os.path.join(ROOT, "tools", "ncu_regions.py"), rep2, mangled], stdout=f)
if os.path.exists(os.path.join(G, f"policy_{tag}.jsonl")):
    shutil.copy(os.path.join(G, f"policy_{tag}.jsonl"), os.path.join(P, f"{tag}_policy_kernels.jsonl"))
# SASS mnemonics of the tensor-core kernel (tcgen05.mma / tcgen05.ld / tcgen05.st / commit / alloc)
lib = os.path.join(ROOT, "collision_avoidance_b200", "liborca_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
on, counts = False, {}
for ln in sass.splitlines():
    if "Function :" in ln:
        on = "policy_mlp_tc" in ln
    elif on:
        for m in ("UTCMMA", "UTCBAR", "LDTM", "STTM", "UTCATOMSWS", "SYNCS", "UTCHMMA", "UTCQMMA"):
            if m in ln:
                op = [t for t in ln.split() if t.startswith(m)]
                if op:
                    counts[op[0].rstrip(";")] = counts.get(op[0].rstrip(";"), 0) + 1
with open(os.path.join(P, f"{tag}_policy_tc_sass_mnemonics.txt"), "w") as f:
    f.write("cuobjdump -sass liborca_b200.so, function policy_mlp_tc_kernel: tensor-core / tensor-memory instructions\n")
    for k in sorted(counts):
        f.write(f"{counts[k]:4d}  {k}\n")
print(open(os.path.join(P, "traffic.json")).read())
