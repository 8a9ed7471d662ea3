"""Small driver for ncu: N forwards of the bench workload (one batch), nothing else."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from sps_b200.models import SPSModel

n_fwd = int(sys.argv[1]) if len(sys.argv) > 1 else 3
batch = int(sys.argv[2]) if len(sys.argv) > 2 else bench.BATCH
rows = bench.make_batches(0, n_distinct=1, batch=batch)[0]
pts = torch.as_tensor(np.ascontiguousarray(rows[:, :5])).cuda()
model = SPSModel(bench.VOXEL, max_points=len(pts))
model.MinkUNet.load_state_dict({k: torch.as_tensor(v) for k, v in bench.random_state_dict().items()})
model = model.cuda().eval()
for _ in range(n_fwd):
    s = model(pts)
torch.cuda.synchronize()
model.check()
print("ok", float(s.mean()))
