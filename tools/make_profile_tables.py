"""Turn the ncu CSVs of tools/ncu_capture.sh into the tables committed under profiles/:
   forward_metrics_r1.csv (metrics pass over one forward) -> per-kernel table (markdown) + per-stage DRAM traffic (json)
usage: make_profile_tables.py gpurun_out/forward_metrics_r1.csv profiles/r1_forward_kernels_ncu.md profiles/traffic.json"""
import csv, json, sys, collections
src, out_md, out_json = sys.argv[1:4]
rows = [r for r in csv.reader(open(src)) if len(r) >= 15 and r[0].isdigit()]
launch = collections.OrderedDict()
for r in rows:
    d = launch.setdefault(int(r[0]), {"name": r[4]})
    v = float(r[14].replace(",", "")) if r[14] not in ("", "n/a") else 0.0
    unit = r[13]
    scale = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3,
             "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
    d[r[12]] = v * scale
def short(n):
    n = n.replace("void ", "").replace("sps::", "").replace("(int)", "")
    return n.split("(")[0]
agg = collections.OrderedDict()
for d in launch.values():
    a = agg.setdefault(short(d["name"]), collections.defaultdict(float))
    a["n"] += 1
    a["us"] += d.get("gpu__time_duration.sum", 0.0)
    a["rd"] += d.get("dram__bytes_read.sum", 0.0)
    a["wr"] += d.get("dram__bytes_write.sum", 0.0)
    for k, m in (("dram", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), ("l1", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
                 ("lts", "lts__throughput.avg.pct_of_peak_sustained_elapsed"), ("tensor", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                 ("warps", "sm__warps_active.avg.pct_of_peak_sustained_active"), ("regs", "launch__registers_per_thread"),
                 ("l1hit", "l1tex__t_sector_hit_rate.pct"), ("l2hit", "lts__t_sector_hit_rate.pct"), ("issue", "smsp__issue_active.avg.pct_of_peak_sustained_active")):
        a[k] += d.get(m, 0.0) * d.get("gpu__time_duration.sum", 0.0)   # time-weighted
tot = sum(a["us"] for a in agg.values())
with open(out_md, "w") as f:
    f.write(f"ncu metrics pass over ONE forward of the bench workload (batch 8, 2.47 M rows, {len(launch)} launches; second forward of tools/profile_forward.py;\n"
            "`tools/ncu_capture.sh`).  Per-launch times are cold-cache and serialised: compare shares.  Percentages are time-weighted means.\n\n")
    f.write("| kernel | launches | time us | share | DRAM rd MB | DRAM wr MB | dram % | l1tex % | lts % | L1 hit % | L2 hit % | issue % | tensor % | warps % | regs |\n")
    f.write("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|\n")
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        t = max(a["us"], 1e-9)
        f.write(f"| `{name}` | {int(a['n'])} | {a['us']:.1f} | {100*a['us']/tot:.1f}% | {a['rd']/1e6:.1f} | {a['wr']/1e6:.1f} | {a['dram']/t:.1f} | "
                f"{a['l1']/t:.1f} | {a['lts']/t:.1f} | {a['l1hit']/t:.1f} | {a['l2hit']/t:.1f} | {a['issue']/t:.1f} | {a['tensor']/t:.2f} | {a['warps']/t:.1f} | {a['regs']/t:.0f} |\n")
    f.write(f"\ntotal kernel time {tot:.1f} us over {len(launch)} launches\n")
# per-stage traffic: the n-th launch of a kernel family maps onto a bench stage name
conv_names = ["conv1p1s2", "block1.conv1", "block1.conv2", "conv2p2s2", "block2.conv1", "block2.conv2", "conv3p4s2", "block3.conv1",
              "block3.conv2", "conv4p8s2", "block4.conv1", "block4.conv2", "convtr4p16s2", "block5.conv1", "block5.conv2", "convtr5p8s2",
              "block6.conv1", "block6.conv2", "convtr6p4s2", "block7.conv1", "block7.conv2", "convtr7p2s2", "block8.conv1", "block8.conv2+final"]
fam = {"k_conv_umma6": conv_names, "k_kernel_map_blk3": ["kmap3"], "k_conv0_const": ["conv0+kmap5"],
       "k_insert_points": ["vox.insert"], "k_assign_points": ["vox.assign"], "k_tile_masks_perm": ["slices"], "k_up_order": ["up_order"]}
seen = collections.Counter()
traffic = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu metrics pass over one forward of the bench workload "
                       "(tools/ncu_capture.sh)"}
for d in launch.values():
    for key, names in fam.items():
        if short(d["name"]).startswith(key):
            i = seen[key]; seen[key] += 1
            if i < len(names):
                traffic[names[i]] = int(d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0))
json.dump(traffic, open(out_json, "w"), indent=1)
print(open(out_md).read())
