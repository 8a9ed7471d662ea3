import json, sys
r = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k: r[k] for k in ("value", "ms_per_step", "gpu_launches", "clocks")}, "e2e", r["e2e"]["value"])
print("roofline", r.get("roofline"))
tot = 0
for k, v in r.get("stages", {}).items():
    tot += v["ms"]
    print(f"{k:22s} {v['ms']:8.4f} ms  {v.get('GB/s',0):8.1f} GB/s  {v.get('TFLOP/s',0):7.3f} TF/s")
print("sum stages", tot, r.get("cpu_baseline", {}).get("value"))
