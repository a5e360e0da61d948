"""Per-kernel summary of an ncu report exported with `ncu -i X.ncu-rep --page raw --csv`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
want = [("name", "Kernel Name"), ("grid", "Grid Size"), ("dur_us", "gpu__time_duration.sum"),
        ("tensor_pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        ("dram_rd_MB", "dram__bytes_read.sum"), ("dram_wr_MB", "dram__bytes_write.sum"),
        ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        ("l2_to_sm_GB", "l1tex__m_xbar2l1tex_read_bytes.sum"),
        ("regs", "launch__registers_per_thread"), ("sm_cycles", "sm__cycles_elapsed.avg"),
        ("sm_mhz", "sm__cycles_elapsed.avg.per_second"), ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active")]
idx = [(k, hdr.index(h)) for k, h in want if h in hdr]
def conv(v, u):
    try: f = float(v.replace(",", ""))
    except Exception: return v
    scale = {"ms": 1e3, "us": 1, "ns": 1e-3, "s": 1e6, "Gbyte": 1e3, "Mbyte": 1, "Kbyte": 1e-3, "byte": 1e-6, "Tbyte": 1e6}.get(u)
    return round(f * scale, 3) if scale else round(f, 3)
print("\t".join(k for k, _ in idx))
for d in data:
    out = []
    for k, i in idx:
        v = d[i]
        if k == "name": v = v.split("(")[0][-40:]
        elif k == "l2_to_sm_GB": v = conv(v, units[i]); v = round(v / 1e3, 3) if isinstance(v, float) else v
        else: v = conv(v, units[i])
        out.append(str(v))
    print("\t".join(out))
