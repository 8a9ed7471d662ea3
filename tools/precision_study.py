"""CPU study of where the score error of reduced-precision operands comes from (no GPU needed).

Emulates the fused forward's numerics on the numpy oracle: BatchNorm folded into the kernels in fp32, weights and/or
stored activations rounded (fp16, or an fp16 hi+lo pair = ~21 bits), products accumulated exactly (fp64), and
compares the sigmoid scores with the fp64 truth for a head gain that spreads the scores over (0,1).

    python tools/precision_study.py [sensor] [gain]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

from conftest import make_case  # noqa: E402
from oracle import sps_oracle as O  # noqa: E402

f64 = np.float64


def r16(x):
    return np.asarray(x, np.float32).astype(np.float16).astype(f64)


def r16x2(x):   # fp16 hi + fp16 lo
    x = np.asarray(x, np.float32)
    hi = x.astype(np.float16).astype(np.float32)
    lo = (x - hi).astype(np.float16).astype(np.float32)
    return (hi.astype(f64) + lo.astype(f64))


def ident(x):
    return np.asarray(x, f64)


def fold(sd, kname, bn):
    w = np.asarray(sd[kname], np.float32)
    g, b, m, v = (np.asarray(sd[f"{bn}.bn.{s}"], np.float32) for s in ("weight", "bias", "running_mean", "running_var"))
    inv = np.float32(1.0) / np.sqrt(v + np.float32(1e-5))
    scale = (g * inv).astype(f64)
    shift = b.astype(f64) - m.astype(f64) * scale
    return (w.astype(f64) * scale).astype(np.float32), shift


def forward(lv, sd, wr, ar, per_layer=None):
    """wr / ar: rounding of weights / stored activations; per_layer: {layer name: (wr, ar_out)} overrides."""
    per_layer = per_layer or {}

    def rules(name):
        return per_layer.get(name, (wr, ar))
    x0 = np.full((len(lv.coords[0]), 1), 0.5, f64)
    w, sh = fold(sd, "conv0p1s1.kernel", "bn0")
    x = rules("conv0")[1](O.relu(O.conv(x0, lv.nbr5, w.astype(f64), f64) + sh))   # conv0 runs in fp32 on CUDA cores
    skips = [x]
    enc = [("conv1p1s2", "bn1", "block1"), ("conv2p2s2", "bn2", "block2"), ("conv3p4s2", "bn3", "block3"),
           ("conv4p8s2", "bn4", "block4")]

    def block(x, nbr, name, last=False):
        w1, s1 = fold(sd, f"{name}.0.conv1.kernel", f"{name}.0.norm1")
        wr1, ar1 = rules(name + ".conv1")
        h = ar1(O.relu(O.conv(x, nbr, wr1(w1), f64) + s1))
        w2, s2 = fold(sd, f"{name}.0.conv2.kernel", f"{name}.0.norm2")
        wr2, ar2 = rules(name + ".conv2")
        out = O.conv(h, nbr, wr2(w2), f64) + s2
        if f"{name}.0.downsample.0.kernel" in sd:
            wd, sdn = fold(sd, f"{name}.0.downsample.0.kernel", f"{name}.0.downsample.1")
            out = out + x @ wr2(wd.reshape(x.shape[1], -1)) + sdn
        else:
            out = out + x
        out = O.relu(out)
        return out if last else ar2(out)
    for L, (cname, bname, blk) in enumerate(enc):
        w, sh = fold(sd, cname + ".kernel", bname)
        wrc, arc = rules(cname)
        x = arc(O.relu(O.conv_stride2(x, lv.parent[L], lv.koff[L], wrc(w), len(lv.coords[L + 1]), f64) + sh))
        x = block(x, lv.nbr3[L + 1], blk)
        skips.append(x)
    dec = [("convtr4p16s2", "bntr4", "block5"), ("convtr5p8s2", "bntr5", "block6"), ("convtr6p4s2", "bntr6", "block7"),
           ("convtr7p2s2", "bntr7", "block8")]
    for j, (cname, bname, blk) in enumerate(dec):
        L = 3 - j
        w, sh = fold(sd, cname + ".kernel", bname)
        wrc, arc = rules(cname)
        x = arc(O.relu(O.conv_transpose2(x, lv.parent[L], lv.koff[L], wrc(w), f64) + sh))
        x = np.concatenate([x, skips[L]], axis=1)
        x = block(x, lv.nbr3[L], blk, last=(j == 3))
    wf = np.asarray(sd["final.kernel"], f64).reshape(x.shape[1], -1)
    return x @ wf + np.asarray(sd["final.bias"], f64).reshape(1, -1)


def main():
    sensor = sys.argv[1] if len(sys.argv) > 1 else "tiny"
    gain = float(sys.argv[2]) if len(sys.argv) > 2 else 8.0
    rows = make_case(sensor, seed=2, submap="voxel", n_map_poses=6)
    pts = rows[:, :5]
    sd = O.make_state_dict(seed=0, randomize_bn=True)
    sd["final.kernel"] = sd["final.kernel"] * np.float32(gain)
    c0, inv = O.voxelize(pts, 0.1)
    lv = O.Levels(c0)
    truth = forward(lv, sd, ident, ident)
    s_truth = O.sigmoid(truth[inv, 0])
    print(f"{sensor}: {len(pts)} rows, {len(c0)} voxels, gain {gain}, score spread [{s_truth.min():.3f}, {s_truth.max():.3f}]")

    def report(tag, logits):
        s = O.sigmoid(logits[inv, 0])
        e = np.abs(s - s_truth)
        lab = np.mean((s < 0.84) == (s_truth < 0.84))
        print(f"{tag:58s} max {e.max():.2e} mean {e.mean():.2e} p99.9 {np.quantile(e, 0.999):.2e} labels {lab:.5f}")
    report("fp16 weights + fp16 activations (round-1 default)", forward(lv, sd, r16, r16))
    report("fp16 hi+lo weights, fp16 activations", forward(lv, sd, r16x2, r16))
    report("fp16 weights, exact activations", forward(lv, sd, r16, ident))
    report("exact weights, fp16 activations", forward(lv, sd, ident, r16))
    hi = (r16x2, r16x2)
    layers = ["conv1p1s2", "block1.conv1", "block1.conv2", "conv2p2s2", "block2.conv1", "block2.conv2", "conv3p4s2",
              "block3.conv1", "block3.conv2", "conv4p8s2", "block4.conv1", "block4.conv2", "convtr4p16s2", "block5.conv1",
              "block5.conv2", "convtr5p8s2", "block6.conv1", "block6.conv2", "convtr6p4s2", "block7.conv1", "block7.conv2",
              "convtr7p2s2", "block8.conv1", "block8.conv2", "conv0"]
    # error contribution per layer: ONLY that layer reduced (weights + its stored output), the rest exact
    exact = (ident, ident)
    for name in layers:
        pl = {n: exact for n in layers}
        pl[name] = (r16, r16)
        report(f"only {name} in fp16", forward(lv, sd, ident, ident, pl))
    # candidates: everything fp16 except the level-0/1 tail in hi+lo
    for keep in (["block8.conv1", "block8.conv2", "convtr7p2s2"],
                 ["block8.conv1", "block8.conv2", "convtr7p2s2", "block7.conv1", "block7.conv2", "convtr6p4s2"],
                 ["conv0", "block8.conv1", "block8.conv2", "convtr7p2s2"]):
        pl = {n: hi for n in keep}
        report("fp16 except hi+lo: " + ",".join(keep), forward(lv, sd, r16, r16, pl))
        pl = {n: (r16x2, r16) for n in keep}
        report("fp16 except hi+lo WEIGHTS only: " + ",".join(keep), forward(lv, sd, r16, r16, pl))


if __name__ == "__main__":
    main()
