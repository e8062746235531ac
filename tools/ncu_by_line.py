#!/usr/bin/env python
"""Attribute an ncu SASS-page CSV to source lines.

  ncu -i prof.ncu-rep --page source --csv --print-source sass > sass.csv
  cuobjdump -xelf all liborca_b200.so ; nvdisasm --print-line-info x.cubin > all.sass
  python tools/ncu_by_line.py sass.csv all.sass <mangled kernel name> [--top 40]

Joins the per-instruction counters of the profile with nvdisasm's line table (built with
-lineinfo) and prints the executed warp-instructions, thread efficiency and stall samples per
source line and per labelled region.
"""
import csv
import re
import sys
from collections import defaultdict


def parse_disasm(path, kernel):
    """offset -> (file, line, inline chain string)"""
    out = {}
    cur = ("?", 0, "")
    on = False
    with open(path) as f:
        for ln in f:
            if ln.startswith("//--------------------- .text."):
                on = (".text." + kernel + " ") in ln
                continue
            if not on:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
            if m:
                cur = (m.group(1).split("/")[-1], int(m.group(2)), m.group(3).strip())
                continue
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(\S.*);", ln)
            if m:
                out[int(m.group(1), 16)] = cur + (m.group(2).strip(),)
    return out


def main():
    sass_csv, disasm, kernel = sys.argv[1:4]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
    table = parse_disasm(disasm, kernel)
    rows = list(csv.reader(open(sass_csv)))
    # find the header row of the requested kernel (first kernel in the file is used)
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    col = {h: i for i, h in enumerate(hdr)}
    base = None
    per_line = defaultdict(lambda: [0, 0, 0, 0])  # inst, thread inst, samples, not-issued samples
    per_op = defaultdict(int)
    total = [0, 0, 0]
    for r in rows[hdr_i + 1:]:
        if not r or not r[0].startswith("0x"):
            break
        addr = int(r[0], 16)
        if base is None:
            base = addr
        off = addr - base
        inst = int(r[col["Instructions Executed"]] or 0)
        tinst = int(r[col["Thread Instructions Executed"]] or 0)
        samp = int(r[col["# Samples"]] or 0)
        f, line, chain, text = table.get(off, ("?", 0, "", ""))
        key = (f, line)
        per_line[key][0] += inst
        per_line[key][1] += tinst
        per_line[key][2] += samp
        total[0] += inst
        total[1] += tinst
        total[2] += samp
        per_op[text.split()[0].split(".")[0] if text else "?"] += inst
    print(f"total warp-inst {total[0]:,}  thread-inst {total[1]:,}  avg active {total[1] / max(1, total[0]):.1f}  samples {total[2]:,}")
    print(f"{'file:line':38s} {'warp-inst':>12s} {'%':>6s} {'act':>5s} {'samples':>8s} {'%':>6s}")
    for key, v in sorted(per_line.items(), key=lambda kv: -kv[1][2])[:top]:
        print(f"{key[0] + ':' + str(key[1]):38s} {v[0]:12,d} {100 * v[0] / total[0]:6.2f} {v[1] / max(1, v[0]):5.1f} "
              f"{v[2]:8,d} {100 * v[2] / max(1, total[2]):6.2f}")
    print("\nopcode mix (warp-inst):")
    for op, n in sorted(per_op.items(), key=lambda kv: -kv[1])[:25]:
        print(f"  {op:12s} {n:12,d} {100 * n / total[0]:6.2f}%")


if __name__ == "__main__":
    main()
