#!/usr/bin/env python
"""Aggregate an ncu SASS-page CSV by CALL PATH (dev tool): the same source line inlined at different
places of the kernel (LP2 of the front half vs LP2 inside LP3) is counted separately.
usage: tools/ncu_phases.py prof.ncu-rep <mangled kernel> [depth=3] [lib.so]

Each SASS instruction is keyed by the outermost `depth` frames of its inline chain
(nvdisasm --print-line-info-inline), printed as function@line > function@line > ..."""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
CSRC = os.path.join(ROOT, "collision_avoidance_b200", "csrc")


def regions(path):
    out = []
    for i, l in enumerate(open(path), 1):
        mm = (re.match(r"^(?:ORCA_HD|__global__|__device__)\s.*?(\w+)\s*\(", l) or re.match(r"^struct (\w+)", l)
              or re.match(r"^\s+ORCA_HD\s.*?(\w+)\s*\(", l))
        if mm:
            out.append((i, mm.group(1)))
    return out


def parse_chains(path, kernel):
    """offset -> list of (file, line) frames, innermost first"""
    out = {}
    cur = []
    pending = []
    on = False
    with open(path) as f:
        for ln in f:
            if ln.startswith("//--------------------- .text."):
                on = (".text." + kernel + " ") in ln or ln.rstrip().endswith(".text." + kernel)
                continue
            if not on:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                pending.append((m.group(1).split("/")[-1], int(m.group(2))))
                continue
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(\S.*);", ln)
            if m:
                if pending:
                    cur = pending
                    pending = []
                out[int(m.group(1), 16)] = cur
    return out


def main():
    rep, kernel = sys.argv[1:3]
    depth = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    lib = sys.argv[4] if len(sys.argv) > 4 else os.path.join(ROOT, "collision_avoidance_b200", "liborca_b200.so")
    tmp = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = os.path.join(tmp, "all.sass")
    with open(dis, "w") as f:
        subprocess.check_call(["nvdisasm", "--print-line-info-inline", os.path.join(tmp, cubin)], stdout=f)
    sass = os.path.join(tmp, "sass.csv")
    with open(sass, "w") as f:
        subprocess.call(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], stdout=f,
                        stderr=subprocess.DEVNULL)
    chains = parse_chains(dis, kernel)
    R = {f: regions(os.path.join(CSRC, f)) for f in os.listdir(CSRC) if f.endswith((".cuh", ".cu"))}

    def region(f, line):
        r = R.get(f)
        if not r:
            return f.split(".")[0]
        name = "?"
        for i, n in r:
            if i <= line:
                name = n
        return name

    rows = list(csv.reader(open(sass)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    col = {h: i for i, h in enumerate(rows[hi])}
    stall_cols = ["stall_barrier", "stall_wait", "stall_no_inst", "stall_branch_resolving", "stall_short_sb", "stall_long_sb"]
    agg = defaultdict(lambda: [0, 0, 0] + [0] * len(stall_cols))
    tot = [0, 0, 0]
    base = None
    for r in rows[hi + 1:]:
        if not r or not r[0].startswith("0x"):
            break
        a = int(r[0], 16)
        base = base if base is not None else a
        frames = chains.get(a - base, [])
        outer = list(reversed(frames))[:depth]  # outermost first
        key = " > ".join(f"{region(f, l)}@{l}" for f, l in outer) or "?"
        inst = int(r[col["Instructions Executed"]] or 0)
        ti = int(r[col["Thread Instructions Executed"]] or 0)
        sm = int(r[col["# Samples"]] or 0)
        v = agg[key]
        v[0] += inst
        v[1] += ti
        v[2] += sm
        for i, c in enumerate(stall_cols):
            v[3 + i] += int(r[col[c]] or 0)
        tot[0] += inst
        tot[1] += ti
        tot[2] += sm
    print(f"total warp-inst {tot[0]:,} avg active threads {tot[1] / max(1, tot[0]):.1f} samples {tot[2]:,}")
    print("   inst%   act  samples%  [" + " ".join(c.replace("stall_", "") for c in stall_cols) + "]  path")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][2]):
        if v[0] == 0 and v[2] == 0:
            continue
        if v[2] * 400 < tot[2] and v[0] * 400 < tot[0]:
            continue
        st = " ".join(f"{100 * x / max(1, tot[2]):4.1f}" for x in v[3:])
        print(f"{100 * v[0] / tot[0]:7.2f} {v[1] / max(1, v[0]):5.1f} {100 * v[2] / max(1, tot[2]):8.2f}  [{st}]  {k}")


if __name__ == "__main__":
    main()
