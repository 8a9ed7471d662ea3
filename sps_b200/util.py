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


# --------------------------------------------------------------------------------------------------------------
# ROS-path scan I/O on the device: the per-scan host work of sps_node.py:89-107,146-149.
# Messages are duck-typed (sensor_msgs/PointCloud2 and nav_msgs/Odometry attribute names): rospy is not a dependency.
# --------------------------------------------------------------------------------------------------------------
_PF_SIZES = {1: 1, 2: 1, 3: 2, 4: 2, 5: 4, 6: 4, 7: 4, 8: 8}   # sensor_msgs/PointField datatypes -> bytes


class PointField:
    """sensor_msgs/PointField (name, offset, datatype, count)."""
    INT8, UINT8, INT16, UINT16, INT32, UINT32, FLOAT32, FLOAT64 = range(1, 9)

    def __init__(self, name, offset, datatype, count=1):
        self.name, self.offset, self.datatype, self.count = name, offset, datatype, count


class PointCloud2:
    """The attributes of sensor_msgs/PointCloud2 that util.to_numpy / util.to_rosmsg touch (util.py:117-153)."""

    def __init__(self):
        self.header = None
        self.height = self.width = 0
        self.fields = []
        self.is_bigendian = False
        self.point_step = self.row_step = 0
        self.data = b""
        self.is_dense = True


def pointcloud2_to_tensor(pointcloud_msg, device="cuda"):
    """``util.to_numpy`` (util.py:146-153) with the result left on the device: fp32 ``[height*width, nfields]``, every
    field cast to float32, fields in message order.  ``pointcloud_msg.data`` may be bytes / bytearray / numpy uint8 or
    a uint8 tensor (already on the device: no copy)."""
    import ctypes as C
    from . import _cabi
    from .engine import _ptr, _stream
    lib = _cabi.load()
    m = pointcloud_msg
    fields = list(m.fields)
    if any(getattr(f, "count", 1) != 1 for f in fields):
        raise NotImplementedError("PointField.count != 1")
    data = m.data
    if int(m.height) * int(m.width) == 0:
        return torch.empty((0, len(fields)), dtype=torch.float32, device=device)
    if not isinstance(data, torch.Tensor):
        data = torch.frombuffer(bytearray(data), dtype=torch.uint8) if not isinstance(data, np.ndarray) else torch.as_tensor(data)
    dev = torch.device(device)
    with torch.cuda.device(dev):
        data = data.to(device=dev, non_blocking=True)
        n = int(m.height) * int(m.width)
        out = torch.empty((n, len(fields)), dtype=torch.float32, device=dev)
        offs = (C.c_int32 * len(fields))(*[int(f.offset) for f in fields])
        dts = (C.c_int32 * len(fields))(*[int(f.datatype) for f in fields])
        _cabi.check(lib.sps_pointcloud2_unpack(_ptr(data), int(m.width), int(m.height), int(m.point_step), int(m.row_step),
                                               len(fields), offs, dts, int(bool(m.is_bigendian)), _ptr(out), _stream()),
                    "sps_pointcloud2_unpack")
    return out


def to_numpy(pointcloud_msg):
    """util.py:146-153, same return type as the reference (host fp32 array); the unpacking itself runs on the GPU."""
    return pointcloud2_to_tensor(pointcloud_msg).cpu().numpy()


def transform_point_cloud(point_cloud, transformation_matrix):
    """util.py:187-194 on the device: CUDA fp32 ``[N, >=3]`` points x float64 4x4 matrix -> fp32 ``[N,3]`` (the reference
    returns the float64 product and its caller casts it to float32, sps_node.py:106; the kernel does both)."""
    import ctypes as C
    from . import _cabi
    from .engine import _ptr, _stream
    pts = point_cloud if isinstance(point_cloud, torch.Tensor) else torch.as_tensor(np.asarray(point_cloud))
    if not pts.is_cuda:
        raise RuntimeError("transform_point_cloud needs a CUDA tensor: sps_b200 has no CPU path")
    pts = pts.to(torch.float32)
    if pts.stride(-1) != 1:
        pts = pts.contiguous()
    T = np.ascontiguousarray(np.asarray(transformation_matrix, dtype=np.float64).reshape(4, 4))
    out = torch.empty((pts.shape[0], 3), dtype=torch.float32, device=pts.device)
    with torch.cuda.device(pts.device):
        _cabi.check(_cabi.load().sps_transform_points(_ptr(pts), pts.stride(0), pts.shape[0], T.ctypes.data_as(C.c_void_p), _ptr(out),
                                                      _stream()), "sps_transform_points")
    return out


def inverse_transform_point_cloud(transformed_point_cloud, transformation_matrix):
    """util.py:197-206."""
    return transform_point_cloud(transformed_point_cloud, np.linalg.inv(np.asarray(transformation_matrix, dtype=np.float64)))


def quaternion_matrix(quaternion):
    """tf.transformations.quaternion_matrix for a (x, y, z, w) quaternion: homogeneous 4x4 rotation (float64)."""
    q = np.array(quaternion[:4], dtype=np.float64, copy=True)
    nq = np.dot(q, q)
    if nq < np.finfo(float).eps * 4.0:
        return np.identity(4)
    q *= np.sqrt(2.0 / nq)
    q = np.outer(q, q)
    return np.array(((1.0 - q[1, 1] - q[2, 2], q[0, 1] - q[2, 3], q[0, 2] + q[1, 3], 0.0),
                     (q[0, 1] + q[2, 3], 1.0 - q[0, 0] - q[2, 2], q[1, 2] - q[0, 3], 0.0),
                     (q[0, 2] - q[1, 3], q[1, 2] + q[0, 3], 1.0 - q[0, 0] - q[1, 1], 0.0),
                     (0.0, 0.0, 0.0, 1.0)), dtype=np.float64)


def to_tr_matrix(odom_msg):
    """util.py:209-232: translation x rotation of a nav_msgs/Odometry pose (host scalar arithmetic, one 4x4 per scan)."""
    p, o = odom_msg.pose.pose.position, odom_msg.pose.pose.orientation
    translation = np.array([[1, 0, 0, p.x], [0, 1, 0, p.y], [0, 0, 1, p.z], [0, 0, 0, 1]], dtype=np.float64)
    return np.dot(translation, quaternion_matrix([o.x, o.y, o.z, o.w]))


def filter_scan(scan, scores, epsilon):
    """sps_node.py:148: ``scan[scores <= epsilon]`` on the device -- the rows (sensor-frame x, y, z, intensity) of the
    published cloud, in scan order.  Returns (fp32 ``[n,4]`` buffer, device int32 count): no host synchronisation."""
    from . import _cabi
    from .engine import _ptr, _stream
    lib = _cabi.load()
    assert scan.is_cuda and scan.dtype == torch.float32 and scan.shape[1] >= 4 and (scan.stride(1) == 1 or scan.shape[0] == 0)
    n = scan.shape[0]
    scores = scores.reshape(-1).to(torch.float32).contiguous()
    assert scores.numel() == n
    if n == 0:
        return torch.empty((0, 4), dtype=torch.float32, device=scan.device), torch.zeros(1, dtype=torch.int32, device=scan.device)
    with torch.cuda.device(scan.device):
        nbytes = lib.sps_pointcloud2_pack_scratch_bytes(n)
        scratch = torch.empty(nbytes, dtype=torch.uint8, device=scan.device)
        out = torch.empty((max(n, 1), 4), dtype=torch.float32, device=scan.device)
        count = torch.zeros(1, dtype=torch.int32, device=scan.device)
        _cabi.check(lib.sps_pointcloud2_pack(_ptr(scan), scan.stride(0), n, _ptr(scores), float(epsilon), _ptr(out), _ptr(count),
                                             _ptr(scratch), nbytes, _stream()), "sps_pointcloud2_pack")
    return out, count


def to_rosmsg(data, header, frame_id=None):
    """util.py:117-143: x, y, z, intensity FLOAT32, point_step 16.  ``data``: fp32 ``[M,4]`` (tensor or array)."""
    cloud = PointCloud2()
    cloud.header = header
    if frame_id and header is not None:
        cloud.header.frame_id = frame_id
    cloud.fields = [PointField("x", 0, PointField.FLOAT32, 1), PointField("y", 4, PointField.FLOAT32, 1),
                    PointField("z", 8, PointField.FLOAT32, 1), PointField("intensity", 12, PointField.FLOAT32, 1)]
    arr = data.detach().cpu().numpy() if isinstance(data, torch.Tensor) else np.asarray(data)
    arr = np.array(arr, dtype=np.float32)
    cloud.is_bigendian = False
    cloud.point_step = 16
    cloud.row_step = cloud.point_step * len(arr)
    cloud.is_dense = True
    cloud.width = len(arr)
    cloud.height = 1
    cloud.data = arr.tobytes()
    return cloud
