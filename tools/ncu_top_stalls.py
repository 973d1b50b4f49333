"""Print the hottest SASS instructions (by warp-stall samples) from `ncu --page source --csv` output."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
data = [r for r in rows[h + 1:] if len(r) == len(hdr) and r[hdr.index("# Samples")].isdigit()]
iS, iSrc, iE = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
stall_cols = [i for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
tot = sum(int(r[iS] or 0) for r in data)
print("total samples", tot, "instructions", len(data))
agg = {}
for r in data:
    for i in stall_cols:
        agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i] or 0)
print("stall totals:", sorted(((v, k) for k, v in agg.items() if v), reverse=True)[:8])
for r in sorted(data, key=lambda r: -int(r[iS] or 0))[:n]:
    st = sorted(((int(r[i] or 0), hdr[i]) for i in stall_cols), reverse=True)[:2]
    print(f"{int(r[iS]):6d} {100 * int(r[iS]) / max(tot, 1):5.1f}%  exec={r[iE]:>9}  {r[iSrc][:80]:80s} {st}")
