#!/usr/bin/env python
"""Aggregate an ncu SASS-page CSV by source function (dev tool).
usage: tools/ncu_regions.py prof.ncu-rep <mangled kernel> [lib.so]"""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ncu_by_line import parse_disasm  # noqa: E402

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
CSRC = os.path.join(ROOT, "collision_avoidance_b200", "csrc")


def regions(path):
    out = []
    for i, l in enumerate(open(path), 1):
        mm = re.match(r"^(?:ORCA_HD|__global__|__device__)\s.*?(\w+)\s*\(", l) or re.match(r"^struct (\w+)", l)
        if mm:
            out.append((i, mm.group(1)))
    return out


def main():
    rep, kernel = sys.argv[1:3]
    lib = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "collision_avoidance_b200", "liborca_b200.so")
    tmp = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = os.path.join(tmp, "all.sass")
    with open(dis, "w") as f:
        subprocess.check_call(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], stdout=f)
    sass = os.path.join(tmp, "sass.csv")
    with open(sass, "w") as f:
        subprocess.call(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], stdout=f,
                        stderr=subprocess.DEVNULL)
    table = parse_disasm(dis, kernel)
    R = {f: regions(os.path.join(CSRC, f)) for f in os.listdir(CSRC) if f.endswith((".cuh", ".cu"))}

    def region(f, line):
        r = R.get(f)
        if not r:
            return f
        name = "?"
        for i, n in r:
            if i <= line:
                name = n
        return name

    helpers = {"v2", "add", "sub", "neg", "mul", "dot", "det", "abs_sq", "sqr", "div_s", "unit", "left_of",
               "fmin_first", "fmax_first", "get", "set", "pt", "dr", "Lines", "LocalLines"}
    rows = list(csv.reader(open(sass)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    col = {h: i for i, h in enumerate(rows[hi])}
    agg = defaultdict(lambda: [0, 0, 0])
    tot = [0, 0, 0]
    base = None
    for r in rows[hi + 1:]:
        if not r or not r[0].startswith("0x"):
            break
        a = int(r[0], 16)
        base = base if base is not None else a
        f, line, chain, text = table.get(a - base, ("?", 0, "", ""))
        key = region(f, line)
        if key in helpers and chain:
            for ff, ll in re.findall(r'inlined at "([^"]+)", line (\d+)', chain):
                k2 = region(ff.split("/")[-1], int(ll))
                if k2 not in helpers:
                    key = k2
                    break
        inst = int(r[col["Instructions Executed"]] or 0)
        ti = int(r[col["Thread Instructions Executed"]] or 0)
        sm = int(r[col["# Samples"]] or 0)
        v = agg[key]
        v[0] += inst
        v[1] += ti
        v[2] += sm
        tot[0] += inst
        tot[1] += ti
        tot[2] += sm
    print(f"total warp-inst {tot[0]:,} avg active threads {tot[1] / max(1, tot[0]):.1f} samples {tot[2]:,}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][2]):
        if v[0] == 0 and v[2] == 0:
            continue
        print(f"{k:28s} inst {v[0]:12,d} {100 * v[0] / tot[0]:6.2f}%  act {v[1] / max(1, v[0]):5.1f}  "
              f"samples {100 * v[2] / max(1, tot[2]):6.2f}%")


if __name__ == "__main__":
    main()
