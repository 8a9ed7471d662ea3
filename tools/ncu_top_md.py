"""`ncu --set full ... --page raw --csv` -> markdown: per captured launch the headline counters and the top warp-stall reasons.
usage: ncu_top_md.py gpurun_out/r2_top_full.csv [title] > profiles/r2_top_kernels_ncu_full.md"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
title = sys.argv[2] if len(sys.argv) > 2 else "ncu --set full captures"
cols = [("us", "gpu__time_duration.sum"), ("tensor %", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        ("issue %", "smsp__issue_active.avg.pct_of_peak_sustained_active"), ("l1tex %", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("lts %", "lts__throughput.avg.pct_of_peak_sustained_elapsed"), ("dram %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        ("L1 hit %", "l1tex__t_sector_hit_rate.pct"), ("L2 hit %", "lts__t_sector_hit_rate.pct"),
        ("DRAM rd MB", "dram__bytes_read.sum"), ("DRAM wr MB", "dram__bytes_write.sum"),
        ("warps %", "sm__warps_active.avg.pct_of_peak_sustained_active"), ("regs", "launch__registers_per_thread"),
        ("smem KB/CTA", "launch__shared_mem_per_block_dynamic"), ("grid", "launch__grid_size")]
scale = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]


def val(d, col):
    if col not in ix or d[ix[col]] in ("", "n/a"):
        return None
    v = float(d[ix[col]].replace(",", ""))
    u = units[ix[col]]
    if col.startswith("gpu__time") or col.startswith("dram__bytes"):
        v *= scale.get(u, 1.0)
    if col.startswith("launch__shared"):
        v *= {"byte": 1 / 1024, "Kbyte": 1.0}.get(u, 1.0)
    return v


print(f"# {title}\n")
print("One row per captured launch (`ncu --set full --clock-control none --import-source on`, second forward of the workload; "
      "cold caches and serialised replays: compare shapes, not absolute times).  Stalls = warps stalled per issue-active cycle, top four.\n")
print("| # | kernel | " + " | ".join(c for c, _ in cols) + " | top stalls |")
print("|---:|---|" + "---:|" * len(cols) + "---|")
for n, d in enumerate(data):
    name = d[ix["Kernel Name"]].replace("void ", "").split("(")[0]
    cells = []
    for c, col in cols:
        v = val(d, col)
        cells.append("-" if v is None else (f"{v:.0f}" if c in ("regs", "grid") else f"{v:.1f}"))
    st = sorted(((float(d[ix[h]] or 0), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for h in stall),
                reverse=True)[:4]
    print(f"| {n} | `{name}` | " + " | ".join(cells) + " | " + ", ".join(f"{k} {v:.1f}" for v, k in st) + " |")
