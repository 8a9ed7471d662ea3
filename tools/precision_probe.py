"""Score error of the storage formats against the C oracle (hdl-32 case): backend 1 exact fp32, 2 TF32 on fp32
rows, 0/3 fp16 rows; plain and head-amplified weights."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from conftest import make_case
from oracle import sps_oracle as O, me_cpu
from sps_b200 import engine, _cabi
lib = _cabi.load()
rows = make_case("hdl-32", seed=2, submap="voxel", n_map_poses=6)
pts = rows[:, :5]
d = torch.as_tensor(pts).cuda()
base = O.make_state_dict(seed=0, randomize_bn=True)
for gain in (1.0, 8.0, 40.0):
    sd = dict(base); sd["final.kernel"] = sd["final.kernel"] * np.float32(gain)
    ref, _, _ = me_cpu.forward(pts, 0.1, me_cpu.pack_weights(sd))
    net = engine.Net(sd); eng = engine.Engine(len(pts))
    line = [f"gain {gain:5.1f} spread [{ref.min():.3f},{ref.max():.3f}]"]
    for b in (1, 2, 0):
        eng.set_conv_backend(b)
        got = eng.forward(net, d, 0.1).cpu().numpy(); eng.status()
        e = np.abs(got - ref)
        line.append(f"backend {b}: max {e.max():.2e} mean {e.mean():.2e} label agree {np.mean((got < 0.84) == (ref < 0.84)):.5f}")
    print(" | ".join(line))
