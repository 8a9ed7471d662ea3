"""Condense `ncu --page raw --csv` output into the per-kernel table committed under profiles/."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr, data = rows[0], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
def f(d, k, default=0.0):
    try:
        return float(d[ix[k]])
    except Exception:
        return default
agg = collections.OrderedDict()
for d in data:
    name = d[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("sps::", "")
    if "<" in d[ix["Kernel Name"]]:
        name = d[ix["Kernel Name"]].split("(sps_")[0].split("(const")[0].replace("void ", "").replace("sps::", "").replace("(int)", "")
    a = agg.setdefault(name, collections.defaultdict(float))
    a["launches"] += 1
    a["time_us"] += f(d, "gpu__time_duration.sum") * (1e-3 if hdr and rows[1][ix["gpu__time_duration.sum"]] in ("ns", "nsecond") else 1e3 if rows[1][ix["gpu__time_duration.sum"]] in ("ms", "msecond") else 1.0)
    a["dram_rd_MB"] += f(d, "dram__bytes_read.sum") * {"Gbyte": 1e3, "Mbyte": 1.0, "Kbyte": 1e-3, "byte": 1e-6}.get(rows[1][ix["dram__bytes_read.sum"]], 1.0)
    a["dram_wr_MB"] += f(d, "dram__bytes_write.sum") * {"Gbyte": 1e3, "Mbyte": 1.0, "Kbyte": 1e-3, "byte": 1e-6}.get(rows[1][ix["dram__bytes_write.sum"]], 1.0)
    for key, col in (("tensor_pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                     ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                     ("l1_pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
                     ("lts_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
                     ("warps_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
                     ("regs", "launch__registers_per_thread")):
        a[key] += f(d, col)
tot = sum(a["time_us"] for a in agg.values())
print(f"| kernel | launches | time us | share | DRAM rd MB | DRAM wr MB | dram % | l1tex % | lts % | tensor % | warps % | regs |")
print("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["time_us"]):
    n = a["launches"]
    print(f"| `{name}` | {int(n)} | {a['time_us']:.1f} | {100*a['time_us']/tot:.1f}% | {a['dram_rd_MB']:.1f} | {a['dram_wr_MB']:.1f} | "
          f"{a['dram_pct']/n:.1f} | {a['l1_pct']/n:.1f} | {a['lts_pct']/n:.1f} | {a['tensor_pct']/n:.2f} | {a['warps_pct']/n:.1f} | {a['regs']/n:.0f} |")
print(f"\ntotal kernel time {tot:.1f} us over {int(sum(a['launches'] for a in agg.values()))} launches (cold-cache, serialised ncu replays: compare shares, not absolutes)")
