"""SASS opcode histogram of an `ncu --page source --csv` dump (executed warp-instructions per opcode).
usage: ncu -i X.ncu-rep --page source --csv > s.csv ; python tools/ncu_ops.py s.csv [units]   (units: divide counts, e.g. chunks)"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hdr = rows[1] if "Source" in rows[1] else rows[0]
ci = {h: i for i, h in enumerate(hdr)}
ops = collections.Counter()
tot = 0
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    try:
        ex = int(r[ci["Instructions Executed"]])
    except ValueError:
        continue
    src = r[ci["Source"]].strip()
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
    op = m.group(2).split(".")[0] if m else src[:10]
    ops[op] += ex
    tot += ex
print("total warp-instructions", tot, " per unit", tot / units)
for k, v in ops.most_common(30):
    print(f"{k:12s} {v:12d} {100 * v / tot:5.1f}%  per-unit {v / units:.2f}")
