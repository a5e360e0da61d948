"""Warp-stall samples of an `ncu --page source --csv` export aggregated per CUDA source line.
usage: ncu_lines.py <source.csv> <nvdisasm -gi output> <mangled-name substring> [lo hi]
Every SASS row (in address order, 16 bytes apart) is attributed through the line table of the matching function:
to the frame of its inline chain that falls inside [lo, hi] (default: the outermost frame)."""
import csv, re, sys
src, dis, fn = sys.argv[1:4]
lo, hi = (int(sys.argv[4]), int(sys.argv[5])) if len(sys.argv) > 5 else (0, 0)
lines = open(dis).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and fn in l and l.rstrip().endswith(":"))
off2line, chain = {}, []
pending = []
for l in lines[start + 1:]:
    if l.startswith(".text.") or l.startswith("//-----"):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        pending.append((m.group(1).split("/")[-1], int(m.group(2))))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", l)
    if m:
        if pending:
            chain, pending = pending, []
        off2line[int(m.group(1), 16)] = chain
rows = list(csv.reader(open(src)))
hdr = rows[1]
iA, iS, iSrc = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Source")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = [r for r in rows[2:] if len(r) > iS and r[iA].startswith("0x")]
base = int(data[0][iA], 16)
agg, tot = {}, 0
for r in data:
    off = int(r[iA], 16) - base
    ch = off2line.get(off, [])
    pick = ch[-1] if ch else ("?", 0)
    if lo:
        for f, n in ch:
            if f == "conv_tc.cu" and lo <= n <= hi:
                pick = (f, n); break
    n = int(float(r[iS] or 0)); tot += n
    a = agg.setdefault(pick, [0, {}])
    a[0] += n
    for i in stall_cols:
        v = int(float(r[i] or 0))
        if v: a[1][hdr[i]] = a[1].get(hdr[i], 0) + v
srcfile = {}
print("total samples", tot)
for (f, n), (s, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[6]) if len(sys.argv) > 6 else 40]:
    text = ""
    if f == "conv_tc.cu":
        if f not in srcfile:
            srcfile[f] = open("learning-driven-image-compression-algorithm_b200/csrc/conv_tc.cu").read().split("\n")
        text = srcfile[f][n - 1].strip()[:90]
    top = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print(f"{f}:{n:5d} {s:6d} {100 * s / max(tot, 1):5.1f}%  {text:90s} {top}")
