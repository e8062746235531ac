#!/usr/bin/env python
"""Basic-block level view of an ncu capture (dev tool): executions, active lanes, samples.
usage: tools/ncu_blocks.py prof.ncu-rep <mangled kernel> [top]"""
import csv
import os
import subprocess
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ncu_by_line import parse_disasm  # noqa: E402

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")


def main():
    rep, kernel = sys.argv[1:3]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 45
    lib = os.path.join(ROOT, "collision_avoidance_b200", "liborca_b200.so")
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
    base = None
    cur = None
    blocks = []
    for r in rows[hi + 1:]:
        if not r or not r[0].startswith("0x"):
            break
        a = int(r[0], 16)
        base = base if base is not None else a
        f, line, chain, text = table.get(a - base, ("?", 0, "", ""))
        inst = int(r[col["Instructions Executed"]] or 0)
        ti = int(r[col["Thread Instructions Executed"]] or 0)
        smp = int(r[col["# Samples"]] or 0)
        if cur is None or inst != cur["inst"]:
            cur = {"start": a - base, "inst": inst, "n": 0, "ti": 0, "smp": 0, "lines": set()}
            blocks.append(cur)
        cur["n"] += 1
        cur["ti"] += ti
        cur["smp"] += smp
        cur["lines"].add((f, line))
    tot = sum(b["inst"] * b["n"] for b in blocks)
    tsmp = sum(b["smp"] for b in blocks)
    big = sorted(blocks, key=lambda b: -b["inst"] * b["n"])[:top]
    for b in sorted(big, key=lambda b: b["start"]):
        ls = sorted(b["lines"])
        lines = ",".join(f"{f.split('.')[0][-5:]}:{l}" for f, l in ls[:7])
        print(f"@{b['start']:6x} n={b['n']:4d} exec={b['inst']:9,d} inst%={100 * b['inst'] * b['n'] / tot:5.2f} "
              f"act={b['ti'] / max(1, b['inst'] * b['n']):5.1f} smp%={100 * b['smp'] / max(1, tsmp):5.2f} {lines}")


if __name__ == "__main__":
    main()
