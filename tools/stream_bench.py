"""Config 3 of BASELINE.json (streamed ROS path): HDL-32-like scans (~57.6k points) against a 1M-voxel
base map -- per scan: prune against the replicated map hash -> assemble -> forward (sps_infer_scan,
no host synchronisation inside).  Prints scans/s and the mean submap size."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sps_b200 import synth, engine
import bench

n_scans = int(sys.argv[1]) if len(sys.argv) > 1 else 200
if os.environ.get("SPS_PATTERN_SORT"):      # 0 never, 1 inputs >= 400k rows (default), 2 always
    engine.set_defaults(pattern_sort=int(os.environ["SPS_PATTERN_SORT"]))
world = synth.World(0)
traj = synth.loop_trajectory(radius=35.0, n=400)
t0 = time.time()
base = synth.base_map(world, "hdl-32", n_poses=400, seed=0, voxel=0.1, trajectory=traj, target_voxels=1_000_000)
print(f"map: {len(base)} voxels in {time.time() - t0:.1f} s", file=sys.stderr)
scans = [torch.as_tensor(synth.scan(world, "hdl-32", traj(7 * i), seed=i)).cuda() for i in range(32)]
mh = engine.MapHash(torch.as_tensor(base).cuda(), 0.1)
net = engine.Net(bench.random_state_dict())
n = len(scans[0])
eng = engine.Engine(2 * n)
outs = [torch.empty(n, device="cuda") for _ in range(2)]
counts = torch.zeros(2, dtype=torch.int32, device="cuda")
for i in range(5):
    mh.infer_scan(eng, net, scans[i % 32], 0.1, out=outs[i & 1], counts=counts)
eng.status()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
subs = []
e0.record()
for i in range(n_scans):
    mh.infer_scan(eng, net, scans[i % 32], 0.1, out=outs[i & 1], counts=counts)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n_scans
t0 = time.perf_counter()
for i in range(n_scans):
    mh.infer_scan(eng, net, scans[i % 32], 0.1, out=outs[i & 1], counts=counts)
cpu_ms = (time.perf_counter() - t0) / n_scans * 1e3
torch.cuda.synchronize()
streamer = engine.ScanStreamer(mh, eng, net, n, 0.1)
ref = mh.infer_scan(eng, net, scans[3], 0.1)[0].clone()
got = streamer.infer(scans[3]).clone()
torch.cuda.synchronize()
assert torch.equal(ref, got), float((ref - got).abs().max())
for i in range(5):
    streamer.infer(scans[i % 32])
torch.cuda.synchronize()
e0.record()
for i in range(n_scans):
    streamer.infer(scans[i % 32])
e1.record()
torch.cuda.synchronize()
graph_ms = e0.elapsed_time(e1) / n_scans
print(json.dumps({"graph_ms_per_scan": graph_ms, "graph_scans_per_s": 1e3 / graph_ms}))
print(json.dumps({"workload": "config3: hdl-32 scan (57600 pts) vs 1M-voxel map, crop+assemble+forward per scan",
                  "map_voxels": len(base), "scans_per_s": 1e3 / ms, "ms_per_scan": ms, "host_enqueue_ms_per_scan": cpu_ms,
                  "submap_voxels_last": int(counts[0].item()), "scan_voxels_last": int(counts[1].item()),
                  "launches_per_scan": eng.launch_count() + 7}))
