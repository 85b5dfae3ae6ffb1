"""Executed (dynamic) SASS mix of one kernel from an ncu report's source page:
   python tools/ncu_exec_mix.py report.ncu-rep > profiles/rN_<kernel>_exec_mix.txt
Warp-level "Instructions Executed" per opcode, with the stall samples that landed on them."""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
lines = txt.splitlines()
name = next(csv.reader([lines[0]]))[1]
rd = csv.DictReader(io.StringIO("\n".join(lines[1:])))
ex, st = collections.Counter(), collections.Counter()
for r in rd:
    m = re.match(r"\s*(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*(?:\.[A-Z0-9_]+)*)", r["Source"])
    if not m:
        continue
    full = m.group(1)
    op = full.split(".")[0]
    if op in ("LDS", "STS", "LDG", "STG", "LDL", "STL", "SYNCS", "UTMALDG", "UBLKCP", "LD", "ST"):
        op = ".".join([op] + [x for x in full.split(".")[1:] if x in ("128", "64", "U8", "U16", "ARRIVE", "TRYWAIT", "EXCH", "2D")])
    ex[op] += int(r["Instructions Executed"] or 0)
    st[op] += int(r["# Samples"] or 0)
tot, tots = sum(ex.values()), sum(st.values())
print(f"kernel {name}: {tot} warp instructions executed, {tots} stall samples   ({rep})")
print(f"{'opcode':18s} {'executed':>12s} {'%':>6s} {'samples':>9s} {'%':>6s}")
for op, n in ex.most_common(45):
    print(f"{op:18s} {n:12d} {100.0 * n / tot:6.2f} {st[op]:9d} {100.0 * st[op] / max(tots, 1):6.2f}")
