"""Segment an `ncu --page source --csv` export at sync landmarks; print samples per segment."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; iS = hdr.index("# Samples"); iSrc = hdr.index("Source")
data = [r for r in rows[2:] if len(r) > iS and r[iS] != "# Samples"]
data = data[:len(data)//2]
def num(x):
    try: return int(float(x))
    except: return 0
tot = sum(num(r[iS]) for r in data)
print("total samples", tot, "sass rows", len(data))
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
bucket = 0; bstart = 0
for i, r in enumerate(data):
    n = num(r[iS]); bucket += n; s = r[iSrc]
    if any(m in s for m in ["SYNCS.PHASECHK", "SYNCS.ARRIVE", "LDTM", "UTCBAR", "MEMBAR", "BAR.SYNC", "WARPSYNC", "UTMALDG", "UTCHMMA"]):
        if bucket >= max(1, tot // 200):
            st = sorted(((num(r[j]), hdr[j]) for j in stall_cols), reverse=True)[:1]
            print(f"[{bstart:5d}-{i:5d}] {bucket:6d} ({100*bucket/tot:4.1f}%) | {s[:78]} own={n} {st[0] if st else ''}")
        bucket = 0; bstart = i + 1
print("tail", bucket, f"({100*bucket/tot:.1f}%)")
