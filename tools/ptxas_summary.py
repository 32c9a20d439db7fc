#!/usr/bin/env python
"""Summarise `nvcc -Xptxas -v` output read from stdin: one line per kernel (registers, spills, stack, smem)."""
import re
import subprocess
import sys

cur = None
mine = False
rows = []
for line in sys.stdin:
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        name = m.group(1)
        try:
            name = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
        except OSError:
            pass
        name = re.sub(r"\(anonymous namespace\)::|uapic::|void ", "", name)
        name = re.sub(r"\(.*\)$", "", name)
        cur = {"name": name, "mangled": m.group(1)}
        rows.append(cur)
        mine = False
        continue
    if "Function properties for" in line:
        mine = cur is not None and cur["mangled"] in line
        continue
    if cur is None:
        continue
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m and mine:
        cur["stack"], cur["sst"], cur["sld"] = map(int, m.groups())
    m = re.search(r"Used (\d+) registers", line)
    if m:
        cur["regs"] = int(m.group(1))
        s = re.search(r"(\d+) bytes smem", line)
        cur["smem"] = int(s.group(1)) if s else 0
for r in rows:
    print(f"{r['name']:<60} regs {r.get('regs', '?'):>4}  spill st/ld {r.get('sst', 0):>5}/{r.get('sld', 0):<5} stack {r.get('stack', 0)}")
