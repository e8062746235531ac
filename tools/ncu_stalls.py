#!/usr/bin/env python
"""Top source lines by stall reason (dev tool).  usage: ncu_stalls.py prof.ncu-rep <kernel> <stall col> [top]"""
import csv
import os
import subprocess
import sys
import tempfile
from collections import defaultdict

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ncu_by_line import parse_disasm  # noqa: E402

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
rep, kernel, colname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
lib = os.environ.get("PROF_LIB", os.path.join(ROOT, "collision_avoidance_b200", "liborca_b200.so"))
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
rows = list(csv.reader(open(sass)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
col = {h: i for i, h in enumerate(rows[hi])}
agg = defaultdict(lambda: [0, 0, set()])
tot = 0
base = None
for r in rows[hi + 1:]:
    if not r or not r[0].startswith("0x"):
        break
    a = int(r[0], 16)
    base = base if base is not None else a
    f, line, chain, text = table.get(a - base, ("?", 0, "", ""))
    v = int(r[col[colname]] or 0)
    key = (f, line)
    agg[key][0] += v
    agg[key][1] += int(r[col["Instructions Executed"]] or 0)
    if v:
        agg[key][2].add(text.split()[0] if text else "?")
    tot += v
print(f"{colname}: total {tot}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{k[0]}:{k[1]:<5d} {v[0]:6d} {100 * v[0] / max(1, tot):6.2f}%  inst={v[1]:10,d} ops={','.join(sorted(v[2]))[:60]}")
