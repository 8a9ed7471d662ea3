"""ctypes wrapper of oracle/me_cpu.c (C/OpenMP restatement of ME's CPU algorithm).

TEST INFRASTRUCTURE / CPU BASELINE ONLY -- see the header of sps_oracle.py.  PARITY UNPINNED."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import sps_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libme_cpu.so")
_lib = None


def build(force=False):
    src = os.path.join(HERE, "me_cpu.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "-B", "libme_cpu.so"], check=True, capture_output=True)
    return LIB


def load():
    global _lib
    if _lib is None:
        build()
        lib = C.CDLL(LIB)
        lib.me_cpu_forward.restype = C.c_int
        lib.me_cpu_forward.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_float, C.c_void_p, C.c_void_p, C.c_int,
                                       C.c_void_p, C.c_void_p]
        lib.me_cpu_voxelize.restype = C.c_int64
        lib.me_cpu_voxelize.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_float, C.c_void_p, C.c_void_p]
        lib.me_cpu_max_threads.restype = C.c_int
        _lib = lib
    return _lib


def pack_weights(sd) -> np.ndarray:
    """state_dict -> flat fp32 blob in ``sps_oracle.layer_shapes()`` order."""
    parts = []
    for name, kind, shape in O.layer_shapes():
        if kind == "bn":
            for suffix in ("weight", "bias", "running_mean", "running_var"):
                parts.append(np.asarray(sd[f"{name}.bn.{suffix}"], np.float32).reshape(-1))
        else:
            a = np.asarray(sd[name], np.float32)
            assert a.size == int(np.prod(shape)), name
            parts.append(a.reshape(-1))
    return np.ascontiguousarray(np.concatenate(parts))


def forward(points, voxel_size, blob, nthreads=0):
    """Returns (scores fp32 [n], voxel counts per level [5], timings [maps, unet, slice] seconds)."""
    lib = load()
    pts = np.ascontiguousarray(points, np.float32)
    n, ld = pts.shape
    scores = np.empty(n, np.float32)
    counts = np.zeros(5, np.int64)
    timings = np.zeros(3, np.float64)
    rc = lib.me_cpu_forward(pts.ctypes.data, n, ld, float(voxel_size), blob.ctypes.data, scores.ctypes.data,
                            int(nthreads), counts.ctypes.data, timings.ctypes.data)
    assert rc == 0
    return scores, counts, timings


def voxelize(points, voxel_size):
    lib = load()
    pts = np.ascontiguousarray(points, np.float32)
    n, ld = pts.shape
    coords = np.empty((max(n, 1), 5), np.int32)
    inv = np.empty(max(n, 1), np.int32)
    v = lib.me_cpu_voxelize(pts.ctypes.data, n, ld, float(voxel_size), coords.ctypes.data, inv.ctypes.data)
    return coords[:v].copy(), inv[:n].copy()


def max_threads():
    return int(load().me_cpu_max_threads())
