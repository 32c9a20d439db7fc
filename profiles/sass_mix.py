#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source sass` export by opcode: executed warp-instructions and
stall samples per opcode class, per kernel.  usage: sass_mix.py src.csv [units_per_kernel]"""
import collections
import csv
import re
import sys

FP64 = ("DFMA", "DADD", "DMUL", "DSETP", "DMNMX")
MEM = ("LDG", "STG", "LD", "ST", "RED", "ATOM", "ATOMG", "REDG", "LDL", "STL", "LDS", "STS", "LDC", "ULDC", "SHFL", "MUFU")
CTRL = ("BRA", "BSSY", "BSYNC", "CALL", "RET", "EXIT", "WARPSYNC", "BAR", "NOP", "BREAK", "YIELD")


def main(path, units=None):
    kernels, cur, hdr = [], None, None
    for row in csv.reader(open(path)):
        if not row:
            continue
        if row[0] == "Kernel Name":
            cur = {"name": row[1], "rows": []}
            kernels.append(cur)
            hdr = None
        elif row[0] == "Address":
            hdr = row
        elif cur is not None and hdr is not None and len(row) == len(hdr):
            cur["rows"].append(dict(zip(hdr, row)))
    for k in kernels:
        ex, st = collections.Counter(), collections.Counter()
        for r in k["rows"]:
            src = r["Source"].strip()
            m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", src)
            base = m.group(2) if m else src.split()[0]
            ex[base] += int(float(r["Instructions Executed"] or 0))
            st[base] += int(float(r["# Samples"] or 0))
        tot, tots = sum(ex.values()), sum(st.values())
        print("=" * 100)
        print(k["name"][:110], " total warp-instr:", tot, " samples:", tots)
        grp = collections.Counter()
        for c, n in ex.items():
            g = "fp64" if c in FP64 else (c if c in MEM else ("ctrl" if c in CTRL else "int/mov/cvt"))
            grp[g] += n
        for g, n in grp.most_common():
            extra = f"  per-unit {n / units:8.1f}" if units else ""
            print(f"  {g:12s} {n:14d} {100.0 * n / tot:6.2f}%{extra}")
        print("  -- detail (top 30 by executed) --")
        for c, n in ex.most_common(30):
            extra = f"  per-unit {n / units:8.1f}" if units else ""
            print(f"  {c:12s} {n:14d} {100.0 * n / tot:6.2f}%  stall-samples {100.0 * st[c] / max(tots, 1):6.2f}%{extra}")


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else None)
