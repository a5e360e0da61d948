"""One step of the `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py, grouped by kernel.
usage: launchlist_summary.py gpurun_out/launches.csv  (a step = the launches between two conv_first_kernel launches)"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr, data = None, []
for r in rows:
    if "Kernel Name" in r:
        hdr = r; continue
    if hdr and len(r) == len(hdr):
        data.append(dict(zip(hdr, r)))
first = [i for i, d in enumerate(data) if "conv_first" in d["Kernel Name"]]
# the timed region of bench.py replays graphs: take the LAST complete step before the eager profiling pass is not
# distinguishable here, so use the second-to-last pair of conv_first launches
a, b = first[-2], first[-1]
step = data[a:b]
agg = collections.OrderedDict()
tot = 0.0
for d in step:
    t = float(d["Metric Value"]) / (1e3 if d["Metric Unit"] in ("ns", "nsecond") else 1.0)
    name = d["Kernel Name"].split("(")[0][:90]
    e = agg.setdefault(name, [0.0, 0])
    e[0] += t; e[1] += 1; tot += t
print("one step (batch 16 x 768x512), ncu --metrics gpu__time_duration.sum --clock-control none (cold caches, serialised launches)")
print(f"kernels {len(step)} total {tot:.1f} us   (launch ids {step[0]['ID']}..{step[-1]['ID']} of {len(data)})")
for name, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"  {t:8.1f} us {100 * t / tot:5.1f}%  x{n:3d}  {name}")
