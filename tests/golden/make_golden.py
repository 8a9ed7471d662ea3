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


# ---- "next" rows: radius submap selection, 4DMOS window, MapMOS scan + map pair ----
def next_rows_inputs():
    """Seeded inputs shared by the generator and the tests."""
    from sps_b200 import synth
    world = synth.World(11)
    base = synth.base_map(world, "tiny", n_poses=6, seed=11, voxel=0.1).astype(np.float32)
    scans = [synth.scan(world, "tiny", pose=(0.3 * i, -0.25 * i, 0.07 * i), seed=40 + i).astype(np.float32) for i in range(4)]
    window = np.vstack([np.hstack([np.zeros((len(s), 1), np.float32), s, np.full((len(s), 1), 200 + i, np.float32)])
                        for i, s in enumerate(scans)])                               # mos4d_node.py:98-117
    rng = np.random.default_rng(11)
    map_idx = rng.integers(0, 6, len(base)).astype(np.float32)
    return base, scans, window, map_idx


def next_rows_state_dicts():
    from oracle import sps_oracle as O
    sd = O.make_state_dict(seed=5, randomize_bn=True)
    sd3 = dict(sd)
    rng = np.random.default_rng(6)
    sd3["final.kernel"] = (rng.standard_normal((8, 3)) * np.sqrt(2.0 / 3)).astype(np.float32)
    sd3["final.bias"] = rng.uniform(-0.3, 0.3, (1, 3)).astype(np.float32)
    return sd, sd3


def build_next_rows():
    """Ball-query lists come from the call the REFERENCE makes (scipy cKDTree.query_ball_tree, blt_dataset.py:258-262);
    the 4DMOS / MapMOS logits from the oracle's restatements (mos4d.py:17-32, mapmos.py:59-83)."""
    from scipy.spatial import cKDTree
    from oracle import sps_oracle as O
    base, scans, window, map_idx = next_rows_inputs()
    sd, sd3 = next_rows_state_dicts()
    lists = cKDTree(scans[0].astype(np.float64)).query_ball_tree(cKDTree(base.astype(np.float64)), 0.1)
    counts = np.array([len(l) for l in lists], np.int32)
    flat = np.concatenate([np.sort(np.asarray(l, np.int64)) for l in lists]) if counts.sum() else np.zeros(0, np.int64)
    shifted = window.copy()
    shifted[:, 4] -= 200
    coords = np.vstack([np.hstack([np.zeros((len(scans[0]), 1), np.float32), scans[0], np.ones((len(scans[0]), 1), np.float32)]),
                        np.hstack([np.zeros((len(base), 1), np.float32), base, np.zeros((len(base), 1), np.float32)])])
    idx = np.concatenate([np.full(len(scans[0]), 6.0, np.float32), map_idx])
    return {"inputs_digest": digest(np.concatenate([base.ravel(), window.ravel(), map_idx])),
            "ball_counts": counts, "ball_sorted_digest": digest(flat), "ball_total": np.array([len(flat)]),
            "mos4d_logits": O.mos4d_forward(shifted, 0.1, sd3).astype(np.float32),
            "mapmos_logits": O.mapmos_forward(coords, idx, 0.1, sd).astype(np.float32)}


def main():
    np.savez_compressed(os.path.join(HERE, "next_rows_s11.npz"), **build_next_rows())
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
