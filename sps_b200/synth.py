"""Deterministic synthetic LiDAR-shaped inputs (no dataset is reachable offline).

World = ground plane + tall boundary walls + random axis-aligned "buildings"; a scan is an exact
ray cast from a sensor pose (beams x azimuths) with 2 cm range noise, so that point density,
surface structure and voxel occupancy resemble a real spinning LiDAR.  A base map is the
voxel-downsampled union of scans from jittered poses of the *static* world; scans may see extra
"dynamic" boxes that the map does not contain.  Everything is numpy, seeded, and cheap enough
to regenerate on the GPU box (nothing here is read from /root/reference).

Input row layout follows the reference: ``[b, x, y, z, t, label]`` with t = 1 scan / 0 map
(src/sps/datasets/blt_dataset.py:173-182, 209-244; src/sps/datasets/util.py:20-21).
"""
from __future__ import annotations

import numpy as np

SENSORS = {
    # name: (beams, azimuths, elevation low, elevation high) in degrees
    "os1-128": (128, 1024, -22.5, 22.5),   # config 1: 131 072 points
    "os1-64": (64, 1024, -22.5, 22.5),     # config 2: 65 536 points (BLT-shaped)
    "hdl-32": (32, 1800, -30.67, 10.67),   # config 3: 57 600 points (NCLT-shaped)
    "dense-128x4096": (128, 4096, -22.5, 22.5),  # config 5 stress: 524 288 points
    "tiny": (16, 128, -15.0, 15.0),        # unit tests
}


class World:
    """Static boxes + optional dynamic boxes inside a walled square arena."""

    def __init__(self, seed=0, half_extent=70.0, n_static=40, n_dynamic=12):
        rng = np.random.default_rng(seed)
        self.half = float(half_extent)

        def boxes(n, hmin, hmax, smin, smax):
            c = rng.uniform(-0.85 * half_extent, 0.85 * half_extent, (n, 2))
            keep = np.hypot(c[:, 0], c[:, 1]) > 6.0          # keep the start area free
            c = c[keep]
            sz = rng.uniform(smin, smax, (len(c), 2))
            h = rng.uniform(hmin, hmax, len(c))
            lo = np.column_stack([c - sz / 2, np.zeros(len(c))])
            hi = np.column_stack([c + sz / 2, h])
            return lo, hi
        self.static = boxes(n_static, 2.5, 12.0, 3.0, 16.0)
        self.dynamic = boxes(n_dynamic, 1.4, 2.2, 0.6, 4.5)


def _ray_boxes(origin, dirs, lo, hi):
    """Slab-test distance of every ray to the nearest of the boxes (inf if none)."""
    best = np.full(len(dirs), np.inf, dtype=np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = [(1.0 / dirs[:, a]).astype(np.float32) for a in range(3)]
        o = origin.astype(np.float32)
        for b in range(len(lo)):
            tmin = None
            for a in range(3):
                t0 = np.float32(lo[b, a] - o[a]) * inv[a]
                t1 = np.float32(hi[b, a] - o[a]) * inv[a]
                mn, mx = np.minimum(t0, t1), np.maximum(t0, t1)
                tmin = mn if tmin is None else np.maximum(tmin, mn)
                tmax = mx if a == 0 else np.minimum(tmax, mx)
            hit = (tmax >= tmin) & (tmin > 0.05) & (tmin < best)
            best = np.where(hit, tmin, best)
    return best.astype(np.float64)


def scan(world: World, sensor="os1-64", pose=(0.0, 0.0, 0.0), seed=0, dynamic=True,
         sensor_height=1.8, noise=0.02):
    """One sweep -> fp32 ``[beams*azimuths, 3]`` points in the WORLD (map) frame, as the
    reference feeds scans already transformed to the map frame (blt_dataset.py:57-75,
    sps_node.py:103).  Row order is beam-major."""
    beams, naz, el_lo, el_hi = SENSORS[sensor]
    rng = np.random.default_rng([seed, 17])
    el = np.deg2rad(np.linspace(el_lo, el_hi, beams))
    az = np.linspace(0.0, 2 * np.pi, naz, endpoint=False) + pose[2]
    ce, se = np.cos(el)[:, None], np.sin(el)[:, None]
    dirs = np.stack([ce * np.cos(az)[None, :], ce * np.sin(az)[None, :], np.broadcast_to(se, (beams, naz))],
                    axis=-1).reshape(-1, 3)
    origin = np.array([pose[0], pose[1], sensor_height])
    H = world.half
    # ground plane z = 0 and the arena walls (an inward-facing box of height 60 m)
    with np.errstate(divide="ignore", invalid="ignore"):
        tg = np.where(dirs[:, 2] < 0, -origin[2] / dirs[:, 2], np.inf)
        tw = np.full(len(dirs), np.inf)
        for ax in (0, 1):
            for side in (-H, H):
                t = (side - origin[ax]) / dirs[:, ax]
                tw = np.where((t > 0) & (t < tw), t, tw)
    t = np.minimum(tg, tw)
    t = np.minimum(t, _ray_boxes(origin, dirs, *world.static))
    if dynamic and len(world.dynamic[0]):
        t = np.minimum(t, _ray_boxes(origin, dirs, *world.dynamic))
    t = t + rng.normal(0.0, noise, len(t))
    pts = origin[None, :] + dirs * t[:, None]
    return pts.astype(np.float32)


def voxel_downsample(points, voxel):
    """Keep the first point met in every ``voxel``-sized cell (floor lattice)."""
    c = np.floor(points.astype(np.float64) / voxel).astype(np.int64)
    key = (c[:, 0] + (1 << 20)) << 42 | (c[:, 1] + (1 << 20)) << 21 | (c[:, 2] + (1 << 20))
    _, first = np.unique(key, return_index=True)
    return points[np.sort(first)]


def base_map(world: World, sensor="os1-64", n_poses=20, seed=0, voxel=0.1, jitter=1.5,
             trajectory=None, target_voxels=None):
    """Base map = voxel-downsampled union of static-world scans.
    ``trajectory`` (callable i -> pose) overrides the jittered poses; ``target_voxels`` stops
    as soon as the map holds that many points."""
    rng = np.random.default_rng([seed, 99])
    clouds, total = [], None
    for i in range(n_poses):
        if trajectory is not None:
            pose = trajectory(i)
        else:
            pose = (rng.uniform(-jitter, jitter), rng.uniform(-jitter, jitter), rng.uniform(0, 2 * np.pi))
        clouds.append(scan(world, sensor, pose, seed=1000 + i, dynamic=False))
        if target_voxels is not None and (i % 8 == 7 or i == n_poses - 1):
            total = voxel_downsample(np.vstack(clouds), voxel)
            clouds = [total]
            if len(total) >= target_voxels:
                return total[:target_voxels]
    return voxel_downsample(np.vstack(clouds), voxel)


def loop_trajectory(radius=35.0, n=1000):
    """Closed loop through the arena used by the streamed configuration (config 3)."""
    def pose(i):
        a = 2 * np.pi * (i % n) / n
        return (radius * np.cos(a), 0.6 * radius * np.sin(a), a + np.pi / 2)
    return pose


def submap_voxel_overlap(map_xyz, scan_xyz, ds):
    """Host-side statement of ``util.prune`` (src/sps/datasets/util.py:85-114) used only to
    PREPARE synthetic inputs: map voxels (truncation lattice) that also hold a scan point,
    returned as voxel corners ``coords * ds`` in fp32."""
    def coords(x):
        return np.trunc(x.astype(np.float32) / np.float32(ds)).astype(np.int64)
    def key(c):
        return (c[:, 0] + (1 << 20)) << 42 | (c[:, 1] + (1 << 20)) << 21 | (c[:, 2] + (1 << 20))
    cm = coords(map_xyz)
    km = key(cm)
    _, first = np.unique(km, return_index=True)
    first = np.sort(first)
    keep = np.isin(km[first], key(coords(scan_xyz)))
    return (cm[first][keep].astype(np.float32) * np.float32(ds)).astype(np.float32)


def submap_radius(map_xyz, scan_xyz, radius):
    """Offline-path submap (blt_dataset.py:222-226, 258-271): every map point within
    ``radius`` of a scan point, once PER scan point (duplicates kept), original coordinates."""
    from scipy.spatial import cKDTree
    lists = cKDTree(scan_xyz).query_ball_tree(cKDTree(map_xyz), radius)
    idx = np.concatenate([np.asarray(l, dtype=np.int64) for l in lists]) if len(lists) else np.zeros(0, np.int64)
    return map_xyz[idx]


def assemble(scan_xyz, submap_xyz, batch_index=0, labels=None, seed=0):
    """Rows ``[b, x, y, z, t, label]``: scan rows (t=1) first, then submap rows (t=0, label 1)
    (blt_dataset.py:209-244 + collate_fn 173-182)."""
    ns, nm = len(scan_xyz), len(submap_xyz)
    if labels is None:
        labels = np.random.default_rng([seed, 5]).uniform(0, 1, ns).astype(np.float32)
    out = np.empty((ns + nm, 6), np.float32)
    out[:, 0] = batch_index
    out[:ns, 1:4] = scan_xyz
    out[ns:, 1:4] = np.asarray(submap_xyz, np.float32).reshape(-1, 3)
    out[:ns, 4] = 1.0
    out[ns:, 4] = 0.0
    out[:ns, 5] = labels
    out[ns:, 5] = 1.0
    return out


def make_batch(sensor="os1-64", batch=8, seed=0, voxel=0.1, submap="radius", world=None, map_xyz=None,
               n_map_poses=12):
    """A collated batch for ``SPSNet.forward`` / ``predict_step``: ``batch`` scans from
    different poses of one world, each with its submap (``radius``: kd-tree ball query as the
    offline path; ``voxel``: prune semantics as the ROS path)."""
    world = world or World(seed)
    if map_xyz is None:
        map_xyz = base_map(world, sensor, n_poses=n_map_poses, seed=seed, voxel=voxel)
    rng = np.random.default_rng([seed, 3])
    rows = []
    for b in range(batch):
        pose = (rng.uniform(-1.5, 1.5), rng.uniform(-1.5, 1.5), rng.uniform(0, 2 * np.pi))
        s = scan(world, sensor, pose, seed=seed * 1000 + b, dynamic=True)
        if submap == "radius":
            m = submap_radius(map_xyz, s, voxel)
        else:
            m = submap_voxel_overlap(map_xyz, s, voxel)
        rows.append(assemble(s, m, b, seed=seed * 1000 + b))
    return np.vstack(rows)
