#!/usr/bin/env python
"""Static size of the HOT code of a kernel (dev tool): SASS instructions that were executed at least
once in an ncu capture, grouped by inline call path -- the instruction-cache working set.
usage: tools/ncu_hot_code.py prof.ncu-rep <mangled kernel> [depth=2] [lib.so]"""
import csv
import os
import subprocess
import sys
import tempfile
from collections import defaultdict

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ncu_phases import CSRC, parse_chains, regions  # noqa: E402

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
rep, kernel = sys.argv[1:3]
depth = int(sys.argv[3]) if len(sys.argv) > 3 else 2
lib = sys.argv[4] if len(sys.argv) > 4 else os.path.join(ROOT, "collision_avoidance_b200", "liborca_b200.so")
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = os.path.join(tmp, "all.sass")
with open(dis, "w") as f:
    subprocess.check_call(["nvdisasm", "--print-line-info-inline", os.path.join(tmp, cubin)], stdout=f)
sass = os.path.join(tmp, "sass.csv")
with open(sass, "w") as f:
    subprocess.call(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], stdout=f, stderr=subprocess.DEVNULL)
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
agg = defaultdict(lambda: [0, 0, 0])
base = None
for r in rows[hi + 1:]:
    if not r or not r[0].startswith("0x"):
        break
    a = int(r[0], 16)
    base = base if base is not None else a
    outer = list(reversed(chains.get(a - base, [])))[:depth]
    key = " > ".join(f"{region(f, l)}@{l}" for f, l in outer) or "?"
    n = int(r[col["Instructions Executed"]] or 0)
    agg[key][0] += 1
    agg[key][1] += 1 if n > 0 else 0
    agg[key][2] += n
tot = sum(v[0] for v in agg.values())
hot = sum(v[1] for v in agg.values())
print(f"static instructions {tot} ({tot * 16 / 1024:.1f} KB), executed at least once {hot} ({hot * 16 / 1024:.1f} KB)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:24]:
    print(f"{v[0]:6d} static {v[1]:6d} hot {v[2]:12,d} executed  {k}")
