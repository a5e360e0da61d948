"""Summarise an `ncu --page source --csv` export: top SASS instructions by stall samples."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
iS = hdr.index("# Samples"); iSrc = hdr.index("Source"); iEx = hdr.index("Instructions Executed")
def num(x):
    try: return int(float(x))
    except Exception: return 0
data = [r for r in rows[2:] if len(r) > iS and r[iS] != "# Samples"]
tot = sum(num(r[iS]) for r in data)
print("rows", len(data), "total samples", tot)
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
agg = {}
for r in data:
    for i in stall_cols:
        agg[hdr[i]] = agg.get(hdr[i], 0) + num(r[i])
print("stall totals:", sorted(((v, k) for k, v in agg.items() if v), reverse=True)[:8])
for idx, r in sorted(enumerate(data), key=lambda t: -num(t[1][iS]))[:n]:
    st = sorted(((num(r[i]), hdr[i]) for i in stall_cols), reverse=True)[:2]
    print(f"#{idx:5d} {num(r[iS]):7d} {100*num(r[iS])/max(tot,1):5.1f}% ex={r[iEx]:>9}  {r[iSrc][:90]:90s} {st}")
