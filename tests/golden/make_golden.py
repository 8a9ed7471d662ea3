"""Writes the golden fixtures under tests/golden/ from the CPU oracle.

The reference cannot run here (MinkowskiEngine is not vendored/installable), so
these vectors come from oracle/sps_oracle.py -- "parity unpinned" -- and serve as a regression
pin for both the oracle and the CUDA path.  Inputs are regenerated from seeds; only compact
outputs are stored: canonical coordinates per level (as a checksum + head), kernel-map pair
counts and a digest, per-point scores.

    python tests/golden/make_golden.py
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CASES = {
    "tiny_voxel_s3": dict(sensor="tiny", seed=3, submap="voxel", batch=1, n_map_poses=4),
    "tiny_batch2_s7": dict(sensor="tiny", seed=7, submap="voxel", batch=2, n_map_poses=4),
    "hdl32_voxel_s2": dict(sensor="hdl-32", seed=2, submap="voxel", batch=1, n_map_poses=6),
}


def digest(a: np.ndarray) -> np.ndarray:
    h = hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest()
    return np.frombuffer(h[:16], dtype=np.uint8).copy()


def case_points(name):
    from sps_b200 import synth
    kw = CASES[name]
    rows = synth.make_batch(sensor=kw["sensor"], batch=kw["batch"], seed=kw["seed"], submap=kw["submap"],
                            n_map_poses=kw["n_map_poses"])
    return rows


def build_case(name):
    from oracle import sps_oracle as O
    rows = case_points(name)
    pts = rows[:, :5]
    sd = O.make_state_dict(seed=0, randomize_bn=True)
    scores, lv, inv = O.sps_forward(pts, 0.1, sd, return_levels=True)
    out = {"n_points": np.array([len(pts)]), "points_digest": digest(pts), "scores": scores.astype(np.float32),
           "inverse_digest": digest(inv.astype(np.int64))}
    for L in range(5):
        canon = O.canonical(lv.coords[L])
        out[f"coords{L}_count"] = np.array([len(canon)])
        out[f"coords{L}_digest"] = digest(canon.astype(np.int32))
        out[f"coords{L}_head"] = canon[:8].astype(np.int32)
        km = O.canonical_kernel_map(lv.nbr3[L], lv.coords[L], lv.coords[L])
        out[f"kmap3_{L}_pairs"] = np.array([len(km)])
        out[f"kmap3_{L}_digest"] = digest(km.astype(np.int64))
    km5 = O.canonical_kernel_map(lv.nbr5, lv.coords[0], lv.coords[0])
    out["kmap5_pairs"] = np.array([len(km5)])
    out["kmap5_digest"] = digest(km5.astype(np.int64))
    return out


def main():
    index = {}
    for name in CASES:
        data = build_case(name)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **data)
        index[name] = {"n_points": int(data["n_points"][0]), "voxels": [int(data[f"coords{L}_count"][0]) for L in range(5)],
                       **CASES[name]}
        print(name, index[name])
    json.dump(index, open(os.path.join(HERE, "index.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
