"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line."""
import csv, sys, collections
def num(x):
    try:
        return int(float(x))
    except ValueError:
        return 0
path, which = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = list(csv.reader(open(path)))
secs, cur = [], None
for r in rows:
    if r and r[0] == "Function Name":
        cur = {"name": r[1], "rows": []}; secs.append(cur)
    elif r and r[0] == "Line No" and cur is not None:
        cur["hdr"] = r
    elif cur is not None and "hdr" in cur and len(r) == len(cur["hdr"]):
        cur["rows"].append(r)
print(len(secs), "kernel sections")
s = secs[which]
hdr = s["hdr"]
i_line, i_src = 0, 1
i_inst = hdr.index("Instructions Executed"); i_samp = hdr.index("# Samples")
stall_cols = [i for i, n in enumerate(hdr) if n.startswith("stall_") and "Not Issued" not in n]
agg = collections.OrderedDict()
for r in s["rows"]:
    key = (r[i_line], r[i_src][:100])
    a = agg.setdefault(key, [0, 0, collections.Counter()])
    a[0] += num(r[i_inst]); a[1] += num(r[i_samp])
    for c in stall_cols:
        if num(r[c]): a[2][hdr[c]] += num(r[c])
tot_i = sum(a[0] for a in agg.values()); tot_s = sum(a[1] for a in agg.values())
print(s["name"][:70], "inst", tot_i, "samples", tot_s)
for (ln, src), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    top = ",".join(f"{k[6:]}:{v}" for k, v in a[2].most_common(3))
    print(f"{ln:>5} inst {100*a[0]/tot_i:5.1f}% samp {100*a[1]/tot_s:5.1f}%  {top:40s} | {src.strip()[:80]}")
