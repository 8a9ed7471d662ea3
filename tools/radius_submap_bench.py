"""Offline-loader submap selection (blt_dataset.py:258-271): GPU ball query vs scipy's query_ball_tree on the bench
workload's shapes (os1-64 scans against the synthetic base map).  Prints one JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scipy.spatial import cKDTree
from sps_b200 import synth
from sps_b200.datasets import RadiusSubmap

world = synth.World(0)
base = synth.base_map(world, "os1-64", n_poses=12, seed=0, voxel=0.1).astype(np.float32)
scans = [synth.scan(world, "os1-64", pose=(1.0 * i, -0.5 * i, 0.1 * i), seed=i).astype(np.float32) for i in range(8)]
d_base = torch.as_tensor(base).cuda()
t0 = time.perf_counter(); sub = RadiusSubmap(d_base, 0.1); torch.cuda.synchronize(); t_build = time.perf_counter() - t0
d_scans = [torch.as_tensor(s).cuda() for s in scans]
for s in d_scans[:2]:
    sub.select_closest_points(s)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
hits = 0
for s in d_scans:
    hits += len(sub.select_closest_points(s))
e1.record(); torch.cuda.synchronize()
gpu_ms = e0.elapsed_time(e1) / len(d_scans)
t0 = time.perf_counter(); tree = cKDTree(base); t_tree = time.perf_counter() - t0
t0 = time.perf_counter()
ref_hits = 0
for s in scans[:2]:
    lists = cKDTree(s).query_ball_tree(tree, 0.1)
    ref_hits += sum(len(l) for l in lists)
cpu_ms = (time.perf_counter() - t0) / 2 * 1e3
print(json.dumps({"map_points": len(base), "scan_points": len(scans[0]), "hits_per_scan": hits // len(scans),
                  "gpu_ms_per_scan": round(gpu_ms, 3), "gpu_index_build_ms": round(t_build * 1e3, 1),
                  "scipy_ms_per_scan": round(cpu_ms, 1), "scipy_tree_build_ms": round(t_tree * 1e3, 1),
                  "speedup": round(cpu_ms / gpu_ms, 1), "host_cpus": os.cpu_count()}))
