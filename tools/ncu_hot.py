"""Summarise an ncu --page source --csv dump: top SASS lines by stall samples with their dominant stall reason.
usage: ncu -i X.ncu-rep --page source --csv > s.csv ; python tools/ncu_hot.py s.csv [topN]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
total = 0
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    try:
        s = int(r[ci["# Samples"]])
    except ValueError:
        continue
    total += s
    data.append((s, r))
data.sort(key=lambda t: -t[0])
print("total samples", total)
for s, r in data[:top]:
    reasons = sorted(((int(r[ci[c]] or 0), c) for c in stall_cols), reverse=True)[:3]
    print(f"{s:7d} {100.0 * s / max(total, 1):5.1f}%  {r[ci['Source']].strip()[:70]:70s} exec={r[ci['Instructions Executed']]:>8s}  "
          + " ".join(f"{c[6:]}={v}" for v, c in reasons if v))
