#!/usr/bin/env python
"""Walk the SASS of one kernel in address order (ncu --page source --csv --print-source sass) in chunks of N instructions:
executed warp-instructions, stall samples with the main reasons, and the opcode mix -- shows which REGION of a long
fully unrolled kernel the time goes to.  usage: addr_mix.py src.csv kernel-substring [chunk]"""
import collections
import csv
import re
import sys

KEYS = ("LDG", "STG", "LDS", "STS", "LDL", "STL", "SHFL", "RED", "MUFU", "DFMA", "DADD", "DMUL", "BRA", "CALL", "SEL", "FSEL")
STALLS = ("stall_long_sb", "stall_no_inst", "stall_wait", "stall_short_sb", "stall_math", "stall_lg", "stall_mio", "stall_not_selected", "stall_branch_resolving", "stall_dispatch")


def main(path, kfilter, chunk=500):
    fn, hdr, rows = None, None, []
    for row in csv.reader(open(path)):
        if not row:
            continue
        if row[0] == "Kernel Name" or row[0] == "Function Name":
            fn = row[1]; hdr = None
        elif row[0] in ("Address", "Line No"):
            hdr = row
        elif hdr is not None and fn and kfilter in fn and len(row) >= len(hdr) - 1:
            d = dict(zip(hdr, row))
            addr = d.get("Address", "")
            if addr.startswith("0x"):
                rows.append(d)
    seen, uniq = set(), []
    for d in rows:
        if d["Address"] not in seen:
            seen.add(d["Address"]); uniq.append(d)
    uniq.sort(key=lambda d: int(d["Address"], 16))
    tot_s = sum(int(float(d["# Samples"] or 0)) for d in uniq)
    tot_e = sum(int(float(d["Instructions Executed"] or 0)) for d in uniq)
    print(f"{kfilter}: {len(uniq)} SASS instructions, executed {tot_e}, samples {tot_s}")
    for c0 in range(0, len(uniq), chunk):
        blk = uniq[c0:c0 + chunk]
        ex = sum(int(float(d["Instructions Executed"] or 0)) for d in blk)
        sm = sum(int(float(d["# Samples"] or 0)) for d in blk)
        st = collections.Counter()
        ops = collections.Counter()
        for d in blk:
            for k in STALLS:
                st[k[6:]] += int(float(d.get(k) or 0))
            m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", d["Source"])
            op = m.group(2) if m else "?"
            for k in KEYS:
                if op.startswith(k):
                    ops[k] += 1
                    break
        sts = " ".join(f"{k}={100 * v // max(sm, 1)}%" for k, v in st.most_common(4))
        print(f"[{c0:6d}] exec={ex:>11d} ({100 * ex / max(tot_e, 1):4.1f}%) samples={sm:>7d} ({100 * sm / max(tot_s, 1):4.1f}%) {sts:58s} " +
              " ".join(f"{k}:{v}" for k, v in ops.most_common(7)))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 500)
