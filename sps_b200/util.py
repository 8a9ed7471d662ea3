"""Host-side mirror of the hot-path helpers of ``sps.datasets.util`` (reference:
src/sps/datasets/util.py) -- same names, argument meaning and return values; ROS message
converters and SE(3) helpers are out of scope.

All set/crop arithmetic runs in the CUDA library through ``engine.MapHash``; nothing here
computes on the CPU except the scalar metric formulas of ``calculate_metrics``.
"""
from __future__ import annotations

import time

import numpy as np
import torch

SCAN_TIMESTAMP = 1   # util.py:20
MAP_TIMESTAMP = 0    # util.py:21


class CoordsFeatStruct:
    """util.py:23-26.  ``cloud`` keeps the metric points so that the device library can
    redo the quantisation bit-exactly; ``map_hash`` caches the replicated map hash."""

    def __init__(self, cloud_coords, features, cloud=None, ds=None):
        self.cloud_coords = cloud_coords
        self.features = features
        self.cloud = cloud
        self.ds = ds
        self.map_hash = None


def load_model(cfg=None, weights_pth=None, device="cuda"):
    """util.py:29-46: strip the Lightning prefix ``model.MinkUNet.``, drop ``MOSLoss`` keys,
    load into ``SPSNet(cfg).model.MinkUNet``, move to the GPU, eval + freeze."""
    assert cfg is not None, "cfg is None!"
    assert weights_pth is not None, "weights_pth is None!"
    from . import models
    ckpt = torch.load(weights_pth, map_location="cpu")
    state_dict = {k.replace("model.MinkUNet.", ""): v for k, v in ckpt["state_dict"].items()}
    state_dict = {k: v for k, v in state_dict.items() if "MOSLoss" not in k}
    model = models.SPSNet(cfg)
    model.model.MinkUNet.load_state_dict(state_dict)
    model = model.to(device)
    model.eval()
    model.freeze()
    return model


def to_coords_features(cloud, feature_type="map", ds=0.1, device="cuda"):
    """util.py:67-82: ``(xyz / ds).int()`` (fp32 division, truncation) + one-hot 2-channel
    feature (scan -> column 0, map -> column 1)."""
    assert feature_type == "map" or feature_type == "scan", "feature_type need to be either 'map' or 'scan'"
    feature_axis = 0 if feature_type == "scan" else 1
    cloud_xyz = torch.as_tensor(cloud)[:, :3].to(device=device, dtype=torch.float32)
    quantization = torch.tensor([ds, ds, ds], dtype=torch.float32, device=cloud_xyz.device)
    cloud_coords = torch.div(cloud_xyz, quantization).int()
    features = torch.zeros(cloud_xyz.shape[0], 2, device=cloud_xyz.device)
    features[:, feature_axis] = 1
    return CoordsFeatStruct(cloud_coords, features, cloud=cloud_xyz.contiguous(), ds=ds)


def prune(map_coords_feat=None, scan_coords_feat=None, ds=0.1):
    """util.py:85-114: map voxels that also hold a scan point, returned as fp32 voxel corners
    ``coordinates * ds`` plus the number of unique scan voxels.  The base-map hash is built on
    the first call and reused (the reference re-hashes the whole map every scan)."""
    from .engine import MapHash
    if map_coords_feat.map_hash is None or map_coords_feat.map_hash.ds != float(ds):
        map_coords_feat.map_hash = MapHash(map_coords_feat.cloud, ds)
    return map_coords_feat.map_hash.crop_voxel(scan_coords_feat.cloud)


def add_timestamp(data, stamp, device=None):
    """util.py:156-160."""
    ones = torch.ones(len(data), 1, dtype=data.dtype, device=data.device)
    return torch.hstack([data, ones * stamp])


def infer(scan_points, submap_points, model, device="cuda"):
    """util.py:163-184: rows [0, x,y,z, t] with the scan rows (t=1) first, one forward,
    ``scores[:len(scan_points)]``; returns (scan_scores, elapsed_seconds)."""
    start_time = time.time()
    assert scan_points.size(-1) == 3, f"Expected 3 columns, but the scan tensor has {scan_points.size(-1)} columns."
    assert submap_points.size(-1) == 3, f"Expected 3 columns, but the submap tensor has {submap_points.size(-1)} columns."
    scan = add_timestamp(scan_points.to(device=device, dtype=torch.float32), SCAN_TIMESTAMP)
    sub = add_timestamp(submap_points.to(device=device, dtype=torch.float32), MAP_TIMESTAMP)
    data = torch.vstack([scan, sub])
    batch = torch.zeros(len(data), 1, dtype=data.dtype, device=data.device)
    tensor = torch.hstack([batch, data]).reshape(-1, 5)
    with torch.no_grad():
        scores = model.forward(tensor)
    checker = getattr(getattr(model, "model", model), "check", None)
    if checker is not None:
        checker()          # synchronises (the reference's elapsed time is a wall clock too); raises on out-of-range coordinates
    scan_scores = scores[: len(scan_points)]
    elapsed_time = time.time() - start_time
    return scan_scores.to(device), elapsed_time


def calculate_metrics(true_labels, predicted_labels):
    """util.py:285-299: class 1 = unstable (score >= epsilon).  Returns
    (precision, recall, f1, accuracy, dIoU) with dIoU = TP / (TP + FN + FP)."""
    t = np.asarray(true_labels)
    p = np.asarray(predicted_labels)
    tp = np.sum((t == 1) & (p == 1))
    tn = np.sum((t == 0) & (p == 0))
    fp = np.sum((t == 0) & (p == 1))
    fn = np.sum((t == 1) & (p == 0))
    precision = tp / (tp + fp) if (tp + fp) != 0 else 0
    recall = tp / (tp + fn) if (tp + fn) != 0 else 0
    f1 = 2 * (precision * recall) / (precision + recall) if (precision + recall) != 0 else 0
    accuracy = (tp + tn) / (tp + tn + fp + fn)
    dIoU = tp / (tp + fn + fp)
    return precision, recall, f1, accuracy, dIoU
