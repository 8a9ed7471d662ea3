"""Side-by-side per-stage times of several bench.py JSON lines: bench_compare.py a.json b.json ..."""
import json, sys
runs = [json.loads(open(p).read().strip().splitlines()[-1]) for p in sys.argv[1:]]
names = [p.split("/")[-1].replace(".json", "") for p in sys.argv[1:]]
print(f"{'':22s}" + "".join(f"{n:>12s}" for n in names))
print(f"{'scans/s':22s}" + "".join(f"{r['value']:12.1f}" for r in runs))
print(f"{'e2e scans/s':22s}" + "".join(f"{r['e2e']['value']:12.1f}" for r in runs))
print(f"{'ms/step':22s}" + "".join(f"{r['ms_per_step']:12.3f}" for r in runs))
for k in runs[0].get("stages", {}):
    print(f"{k:22s}" + "".join(f"{r['stages'].get(k, {}).get('ms', float('nan')):12.4f}" for r in runs))
print(f"{'sum stages':22s}" + "".join(f"{sum(v['ms'] for v in r['stages'].values()):12.3f}" for r in runs))
