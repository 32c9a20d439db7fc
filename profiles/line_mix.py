#!/usr/bin/env python
"""Aggregate `ncu --page source --csv --print-source cuda,sass` by CUDA source line: static SASS instructions, executed
warp-instructions, stall samples (top reasons).  usage: line_mix.py src.csv [kernel-substring] [top]"""
import collections
import csv
import sys


def main(path, kfilter="", top=40):
    cur, hdr, fn, fpath = None, None, None, None
    agg = {}
    for row in csv.reader(open(path)):
        if not row:
            continue
        if row[0] == "File Path":
            fpath = row[1].split("/")[-1]
        elif row[0] == "Function Name":
            fn = row[1]
        elif row[0] == "Line No":
            hdr = row
        elif hdr is not None and len(row) == len(hdr) and fn is not None and kfilter in fn:
            d = dict(zip(hdr, row))   # duplicate "Source" header: the second (SASS) wins
            key = (fn[:60], fpath, int(d["Line No"]) if d["Line No"].isdigit() else -1)
            a = agg.setdefault(key, collections.Counter())
            if d.get("Address"):
                a["static"] += 1
                a["exec"] += int(float(d["Instructions Executed"] or 0))
                a["samples"] += int(float(d["# Samples"] or 0))
                for k in ("stall_long_sb", "stall_no_inst", "stall_wait", "stall_short_sb", "stall_math", "stall_lg", "stall_mio", "stall_not_selected"):
                    a[k] += int(float(d.get(k) or 0))
    byfn = collections.defaultdict(list)
    for (f, p, ln), a in agg.items():
        byfn[f].append((p, ln, a))
    for f, lst in byfn.items():
        tot = collections.Counter()
        for _, _, a in lst:
            tot.update(a)
        print("=" * 120)
        print(f, " static", tot["static"], " exec", tot["exec"], " samples", tot["samples"],
              {k: v for k, v in tot.items() if k.startswith("stall")})
        lst.sort(key=lambda t: -t[2]["samples"])
        for p, ln, a in lst[:top]:
            st = " ".join(f"{k[6:]}={a[k]}" for k in a if k.startswith("stall") and a[k] > 0.15 * max(a["samples"], 1))
            print(f"  {p}:{ln:<5d} static={a['static']:5d} exec={a['exec']:>11d} samples={a['samples']:>7d} ({100 * a['samples'] / max(tot['samples'], 1):4.1f}%)  {st}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "", int(sys.argv[3]) if len(sys.argv) > 3 else 40)
