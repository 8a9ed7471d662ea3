"""Aggregate `ncu --page source --csv --print-source cuda,sass` per CUDA source line (SASS rows are
attributed to the CUDA line that precedes them).  usage: ncu_source_lines.py dump.csv [kernel_index] [top]"""
import csv, sys, collections
path = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
rows = list(csv.reader(open(path)))
# split into (file, function) sections
secs = []
cur = None
fpath = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fpath = r[1]
    elif r[0] == "Function Name":
        cur = {"file": fpath, "name": r[1], "hdr": None, "rows": []}
        secs.append(cur)
    elif r[0] == "Line No" and cur is not None:
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] is not None:
        cur["rows"].append(r)
# kernels = runs of sections; a kernel's first section is the main .cu file
kernels = []
for s in secs:
    if not kernels or (s["file"] or "").endswith(".cu") and (not kernels or kernels[-1][0]["name"] != s["name"] or True):   # a kernel's first section is its .cu file
        kernels.append([s])
    else:
        kernels[-1].append(s)
print(len(secs), "sections,", len(kernels), "kernels")
k = kernels[which]
print(k[0]["name"][:90])
agg = collections.OrderedDict()
tot_s = tot_i = 0
for s in k:
    hdr = s["hdr"]
    i_samp = hdr.index("# Samples"); i_inst = hdr.index("Instructions Executed")
    stall = [(i, n) for i, n in enumerate(hdr) if n.startswith("stall_")]
    line = ("?", "")
    for r in s["rows"]:
        if r[0] not in ("", None):
            line = (s["file"].split("/")[-1] + ":" + r[0], r[1][:90])
            continue
        if len(r) != len(hdr) or not r[2].startswith("0x"):
            continue
        try:
            sm = int(r[i_samp]); ins = int(r[i_inst])
        except ValueError:
            continue
        a = agg.setdefault(line, [0, 0, collections.Counter(), collections.Counter()])
        a[0] += ins; a[1] += sm; tot_s += sm; tot_i += ins
        a[3][r[3].split()[0] if r[3].split() else "?"] += sm
        for i, n in stall:
            try:
                v = int(r[i])
            except ValueError:
                v = 0
            if v: a[2][n[6:]] += v
print("instructions", tot_i, "samples", tot_s)
for (ln, src), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    st = ",".join(f"{n}:{v}" for n, v in a[2].most_common(3))
    op = ",".join(f"{n}:{v}" for n, v in a[3].most_common(2))
    print(f"{ln:>22} inst {100*a[0]/max(tot_i,1):5.1f}% samp {100*a[1]/max(tot_s,1):5.1f}% {st:44s} {op:28s}| {src.strip()[:70]}")
