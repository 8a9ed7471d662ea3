"""CPU oracle for the SPS inference hot path (numpy restatement).

TEST INFRASTRUCTURE ONLY.  Nothing under ``sps_b200/`` may import this module; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs use it, and only as the checker / the CPU baseline.

PARITY UNPINNED: the arithmetic of this path lives in MinkowskiEngine (NVIDIA/MinkowskiEngine,
unpinned master ~= v0.5.4, installed by the reference's ``Dockerfile:38-40``), which is neither
vendored under /root/reference nor installable here, and the reference ships no tests, golden
vectors or fixtures.  This file therefore restates ME's *published*
semantics at the reference's call sites and is pinned only by self-consistency known-answer
tests (dense-conv equivalence against ``torch.nn.functional.conv3d``, hand-computed tiny cases;
see ``tests/test_oracle.py``) -- not by outputs of ME itself.

Every function cites the reference file:line it follows (paths under /root/reference).
Conventions:
  * points are fp32 ``[N,5] = [b, x, y, z, t]`` (metres; t = 1 scan / 0 map, ``util.py:20-21``),
  * coordinates stay in units of the stride-1 lattice at every level (tensor stride
    ``[s,s,s,1]``), kernel maps are dense tables ``nbr[K, V_out]`` (-1 = absent),
  * kernel offset index ``k = i0 + k0*(i1 + k1*(i2 + k2*i3))`` (first spatial axis fastest),
    odd sizes centred, even sizes 0..k-1 (ME HYPER_CUBE region).
"""
from __future__ import annotations

import numpy as np

SCAN_TIMESTAMP = 1  # src/sps/datasets/util.py:20
MAP_TIMESTAMP = 0   # src/sps/datasets/util.py:21
BN_EPS = 1e-5       # nn.BatchNorm1d default wrapped by ME.MinkowskiBatchNorm

PLANES = (8, 16, 32, 64, 64, 32, 16, 8)  # customminkunet.py:11
INIT_DIM = 8                             # customminkunet.py:12


# --------------------------------------------------------------------------------------
# a1/a2  quantise + voxelise   (src/sps/models/models.py:21-25)
# --------------------------------------------------------------------------------------
def quantize(points: np.ndarray, voxel_size: float) -> np.ndarray:
    """models.py:16,21 -- ``coords / Tensor([1,vs,vs,vs,1])`` in IEEE fp32, then ME's
    ``TensorField.sparse()`` floors every column to int32 (models.py:24-25)."""
    p = np.asarray(points, dtype=np.float32)
    q = np.array([1.0, voxel_size, voxel_size, voxel_size, 1.0], dtype=np.float32)
    return np.floor(p / q).astype(np.int32)


def _pack(coords: np.ndarray, lo: np.ndarray, rng: np.ndarray) -> np.ndarray:
    """Mixed-radix int64 key from data-dependent extents (independent of the product's
    fixed bit packing)."""
    c = coords.astype(np.int64) - lo
    key = np.zeros(len(c), dtype=np.int64)
    for d in range(c.shape[1]):
        key = key * rng[d] + c[:, d]
    return key


def _extent(*sets, margin=0):
    lo = np.min([s.min(axis=0) for s in sets if len(s)], axis=0).astype(np.int64) - margin
    hi = np.max([s.max(axis=0) for s in sets if len(s)], axis=0).astype(np.int64) + margin
    rng = hi - lo + 1
    assert np.prod(rng.astype(np.float64)) < 2.0 ** 62, "coordinate extent too large for oracle"
    return lo, rng


def unique_first(coords: np.ndarray):
    """Unique rows in FIRST-OCCURRENCE order plus inverse map (ME CPU ``insert_and_map``:
    sequential insert, so voxel v is the v-th distinct coordinate met; models.py:25)."""
    coords = np.asarray(coords)
    if len(coords) == 0:
        return coords.reshape(0, coords.shape[1]), np.zeros(0, np.int64)
    lo, rng = _extent(coords)
    key = _pack(coords, lo, rng)
    _, first, inv_sorted = np.unique(key, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")          # sorted-unique id -> rank by first index
    rank = np.empty_like(order)
    rank[order] = np.arange(len(order))
    return coords[first[order]], rank[inv_sorted].astype(np.int64)


def voxelize(points: np.ndarray, voxel_size: float):
    """models.py:21-25.  Returns (C0 int32 [V0,5], inverse_mapping int64 [N])."""
    return unique_first(quantize(points, voxel_size))


def canonical(coords: np.ndarray) -> np.ndarray:
    """Lexicographically sorted coordinate set -- the comparison form (row order of ME's
    coordinate map is an implementation artefact)."""
    coords = np.asarray(coords)
    if len(coords) == 0:
        return coords
    return coords[np.lexsort(coords.T[::-1])]


# --------------------------------------------------------------------------------------
# a4  strided coordinate maps   (minkunet.py:64-70: stride=[2,2,2,1])
# --------------------------------------------------------------------------------------
def stride_coords(coords: np.ndarray, new_stride: int):
    """ME stride map: spatial coords floored to a multiple of the new tensor stride, batch
    and t untouched, de-duplicated (first occurrence).  Returns (C_coarse, parent[V_fine])."""
    c = np.asarray(coords).copy()
    c[:, 1:4] = np.floor_divide(c[:, 1:4], new_stride) * new_stride
    return unique_first(c)


# --------------------------------------------------------------------------------------
# kernel offsets + kernel maps
# --------------------------------------------------------------------------------------
def kernel_offsets(ksize, tstride):
    """ME HYPER_CUBE region for kernel sizes (k0,k1,k2,k3) on tensor stride ``tstride``:
    k = i0 + k0*(i1 + k1*(i2 + k2*i3)); odd -> (i - k//2)*ts, even -> i*ts.
    Returns int32 [K,4] offsets on (x,y,z,t)."""
    ksize = list(ksize)
    K = int(np.prod(ksize))
    offs = np.zeros((K, 4), dtype=np.int32)
    for k in range(K):
        r = k
        for d in range(4):
            i = r % ksize[d]
            r //= ksize[d]
            offs[k, d] = (i - ksize[d] // 2) * tstride[d] if ksize[d] % 2 else i * tstride[d]
    return offs


def kernel_map(in_coords, out_coords, offsets):
    """nbr[k, o] = row i of ``in_coords`` with in[i] == out[o] + offsets[k] (batch never
    offset), else -1.  ME: ``out[o] += in[o+delta_k] @ W[k]`` (correlation)."""
    in_coords = np.asarray(in_coords)
    out_coords = np.asarray(out_coords)
    K, Vo = len(offsets), len(out_coords)
    nbr = np.full((K, Vo), -1, dtype=np.int32)
    if len(in_coords) == 0 or Vo == 0:
        return nbr
    margin = int(np.abs(offsets).max()) + 1
    lo, rng = _extent(in_coords, out_coords, margin=margin)
    key_in = _pack(in_coords, lo, rng)
    order = np.argsort(key_in, kind="stable")
    key_sorted = key_in[order]
    for k in range(K):
        q = out_coords.astype(np.int64).copy()
        q[:, 1:5] += offsets[k].astype(np.int64)
        kq = _pack(q, lo, rng)
        pos = np.searchsorted(key_sorted, kq)
        pos[pos >= len(key_sorted)] = 0
        hit = key_sorted[pos] == kq
        nbr[k, hit] = order[pos[hit]]
    return nbr


def canonical_kernel_map(nbr, in_coords, out_coords):
    """Kernel map as a sorted array of (k, in_coord[5], out_coord[5]) rows -- the order-free
    comparison form ("bit-exact" = equality of these sets)."""
    k, o = np.nonzero(np.asarray(nbr) >= 0)
    i = np.asarray(nbr)[k, o]
    rows = np.concatenate([k[:, None].astype(np.int64), np.asarray(in_coords)[i].astype(np.int64),
                           np.asarray(out_coords)[o].astype(np.int64)], axis=1)
    if len(rows) == 0:
        return rows
    return rows[np.lexsort(rows.T[::-1])]


# --------------------------------------------------------------------------------------
# a3/a5/a6/a7  convolutions, a11 BN, ReLU
# --------------------------------------------------------------------------------------
def conv(feat_in, nbr, weight, dtype=np.float32):
    """ME CPU conv: for k = 0..K-1 sequentially: gather rows, (n_k x Cin)@(Cin x Cout),
    scatter-add into out (accumulation dtype = ``dtype``; fp64 is the "truth" mode)."""
    feat_in = np.asarray(feat_in, dtype=dtype)
    weight = np.asarray(weight, dtype=dtype)
    K, Vo = nbr.shape
    out = np.zeros((Vo, weight.shape[-1]), dtype=dtype)
    for k in range(K):
        o = np.nonzero(nbr[k] >= 0)[0]
        if len(o):
            out[o] += feat_in[nbr[k, o]] @ weight[k]
    return out


def conv_stride2(feat_in, parent, koff, weight, n_out, dtype=np.float32):
    """minkunet.py:64-70 (kernel [2,2,2,1], stride [2,2,2,1]): out[parent[f]] += in[f] @ W[k(f)]
    with k(f) = ox + 2*oy + 4*oz, o = (f - parent)/s in {0,1}^3."""
    feat_in = np.asarray(feat_in, dtype=dtype)
    weight = np.asarray(weight, dtype=dtype)
    out = np.zeros((n_out, weight.shape[-1]), dtype=dtype)
    for k in range(8):
        f = np.nonzero(koff == k)[0]
        if len(f):
            np.add.at(out, parent[f], feat_in[f] @ weight[k])
    return out


def conv_transpose2(feat_in, parent, koff, weight, dtype=np.float32):
    """minkunet.py:107-113 (MinkowskiConvolutionTranspose k=[2,2,2,1], s=[2,2,2,1]) onto the
    existing finer coordinate map: out[f] = in[parent[f]] @ W[k(f)] (cached down-conv kernel
    map with in/out swapped, same k)."""
    feat_in = np.asarray(feat_in, dtype=dtype)
    weight = np.asarray(weight, dtype=dtype)
    out = np.zeros((len(parent), weight.shape[-1]), dtype=dtype)
    for k in range(8):
        f = np.nonzero(koff == k)[0]
        if len(f):
            out[f] = feat_in[parent[f]] @ weight[k]
    return out


def child_offset_index(fine_coords, coarse_coords, parent, s):
    """k(f) = ox + 2*oy + 4*oz for the 2x2x2x1 kernel between stride s (fine) and 2s."""
    o = (np.asarray(fine_coords)[:, 1:4] - np.asarray(coarse_coords)[parent][:, 1:4]) // s
    assert o.min(initial=0) >= 0 and o.max(initial=0) <= 1
    return (o[:, 0] + 2 * o[:, 1] + 4 * o[:, 2]).astype(np.int32)


def batchnorm(x, sd, prefix, dtype=np.float32):
    """ME.MinkowskiBatchNorm -> nn.BatchNorm1d eval (attribute ``.bn``; resnet.py:92-94)."""
    g = np.asarray(sd[prefix + ".bn.weight"], dtype=dtype)
    b = np.asarray(sd[prefix + ".bn.bias"], dtype=dtype)
    m = np.asarray(sd[prefix + ".bn.running_mean"], dtype=dtype)
    v = np.asarray(sd[prefix + ".bn.running_var"], dtype=dtype)
    return ((x - m) / np.sqrt(v + dtype(BN_EPS)) * g + b).astype(dtype)


def relu(x):
    return np.maximum(x, 0)


# --------------------------------------------------------------------------------------
# network structure, weights
# --------------------------------------------------------------------------------------
def layer_shapes():
    """State-dict tensor shapes of ``CustomMinkUNet(1, 1, D=4)`` (minkunet.py:52-159,
    customminkunet.py:10-12, resnet.py:96-126).  Returns list of
    (name, kind, shape) with kind in conv/convtr/conv1x1/bn/bias."""
    P, I = PLANES, INIT_DIM
    out = [("conv0p1s1.kernel", "conv", (125, 1, I)), ("bn0", "bn", (I,))]
    inpl = I
    enc = [("conv1p1s2", "bn1", "block1", P[0]), ("conv2p2s2", "bn2", "block2", P[1]),
           ("conv3p4s2", "bn3", "block3", P[2]), ("conv4p8s2", "bn4", "block4", P[3])]

    def block(name, cin, cout):
        r = [(f"{name}.0.conv1.kernel", "conv", (81, cin, cout)), (f"{name}.0.norm1", "bn", (cout,)),
             (f"{name}.0.conv2.kernel", "conv", (81, cout, cout)), (f"{name}.0.norm2", "bn", (cout,))]
        if cin != cout:
            r += [(f"{name}.0.downsample.0.kernel", "conv1x1", (cin, cout)),
                  (f"{name}.0.downsample.1", "bn", (cout,))]
        return r

    for cname, bname, blk, planes in enc:
        out += [(cname + ".kernel", "conv", (8, inpl, inpl)), (bname, "bn", (inpl,))]
        out += block(blk, inpl, planes)
        inpl = planes
    dec = [("convtr4p16s2", "bntr4", "block5", P[4], P[2]), ("convtr5p8s2", "bntr5", "block6", P[5], P[1]),
           ("convtr6p4s2", "bntr6", "block7", P[6], P[0]), ("convtr7p2s2", "bntr7", "block8", P[7], I)]
    for cname, bname, blk, planes, skip in dec:
        out += [(cname + ".kernel", "convtr", (8, inpl, planes)), (bname, "bn", (planes,))]
        out += block(blk, planes + skip, planes)
        inpl = planes
    out += [("final.kernel", "conv1x1", (P[7], 1)), ("final.bias", "bias", (1, 1))]
    return out


def make_state_dict(seed=0, randomize_bn=False):
    """Random-init weights with the reference's distributions (resnet.py:87-94 +
    ME defaults; SURVEY.md §8b "Random-init distributions"): MinkowskiConvolution kernels
    kaiming-normal fan_out; MinkowskiConvolutionTranspose keeps ME's default U(-s,s),
    s = 1/sqrt(Cout*K); final.bias U(-s,s), s = 1/sqrt(Cin*K); BN gamma=1, beta=0, mean=0, var=1.
    ``randomize_bn`` perturbs the BN tensors so that parity tests exercise them."""
    rng = np.random.default_rng(seed)
    sd = {}
    for name, kind, shape in layer_shapes():
        if kind == "conv":
            K, cin, cout = shape
            sd[name] = (rng.standard_normal(shape) * np.sqrt(2.0 / (K * cout))).astype(np.float32)
        elif kind == "conv1x1":
            cin, cout = shape
            sd[name] = (rng.standard_normal(shape) * np.sqrt(2.0 / cout)).astype(np.float32)
        elif kind == "convtr":
            K, cin, cout = shape
            s = 1.0 / np.sqrt(cout * K)
            sd[name] = rng.uniform(-s, s, shape).astype(np.float32)
        elif kind == "bias":
            s = 1.0 / np.sqrt(PLANES[7])
            sd[name] = rng.uniform(-s, s, shape).astype(np.float32)
        elif kind == "bn":
            C = shape[0]
            if randomize_bn:
                sd[name + ".bn.weight"] = rng.uniform(0.5, 1.5, C).astype(np.float32)
                sd[name + ".bn.bias"] = rng.uniform(-0.2, 0.2, C).astype(np.float32)
                sd[name + ".bn.running_mean"] = rng.uniform(-0.2, 0.2, C).astype(np.float32)
                sd[name + ".bn.running_var"] = rng.uniform(0.5, 1.5, C).astype(np.float32)
            else:
                sd[name + ".bn.weight"] = np.ones(C, np.float32)
                sd[name + ".bn.bias"] = np.zeros(C, np.float32)
                sd[name + ".bn.running_mean"] = np.zeros(C, np.float32)
                sd[name + ".bn.running_var"] = np.ones(C, np.float32)
            sd[name + ".bn.num_batches_tracked"] = np.zeros((), np.int64)
    return sd


class Levels:
    """Coordinate sets C_0..C_4, parent tables and the kernel maps one forward needs."""

    def __init__(self, c0):
        self.coords = [c0]
        self.parent = []   # parent[L][f] = row in coords[L+1]
        self.koff = []     # koff[L][f]   = 2x2x2x1 offset index of f inside its parent
        for L in range(4):
            s = 2 ** L
            c, par = stride_coords(self.coords[L], 2 * s)
            self.coords.append(c)
            self.parent.append(par)
            self.koff.append(child_offset_index(self.coords[L], c, par, s))
        self.nbr5 = kernel_map(c0, c0, kernel_offsets([5, 5, 5, 1], [1, 1, 1, 1]))
        self.nbr3 = [kernel_map(self.coords[L], self.coords[L],
                                kernel_offsets([3, 3, 3, 3], [2 ** L] * 3 + [1])) for L in range(5)]


def basic_block(x, nbr, sd, name, dtype):
    """ME ``modules.resnet_block.BasicBlock`` (mirrored at c_ws/src/mapmos/scripts/minkunet.py:66-82)."""
    out = conv(x, nbr, sd[f"{name}.0.conv1.kernel"], dtype)
    out = relu(batchnorm(out, sd, f"{name}.0.norm1", dtype))
    out = conv(out, nbr, sd[f"{name}.0.conv2.kernel"], dtype)
    out = batchnorm(out, sd, f"{name}.0.norm2", dtype)
    if f"{name}.0.downsample.0.kernel" in sd:
        w = np.asarray(sd[f"{name}.0.downsample.0.kernel"], dtype=dtype).reshape(x.shape[1], -1)
        res = batchnorm(np.asarray(x, dtype) @ w, sd, f"{name}.0.downsample.1", dtype)
    else:
        res = x
    return relu(out + res)


def unet_forward(levels: Levels, feat0, sd, dtype=np.float32, taps=None):
    """MinkUNetBase.forward (minkunet.py:161-219) on pre-built coordinate/kernel maps.
    ``taps`` (dict) receives intermediate feature matrices for layer-wise parity tests."""
    def tap(name, x):
        if taps is not None:
            taps[name] = x
        return x
    lv = levels
    out = conv(feat0, lv.nbr5, sd["conv0p1s1.kernel"], dtype)
    out_p1 = tap("out_p1", relu(batchnorm(out, sd, "bn0", dtype)))
    skips = [out_p1]
    x = out_p1
    enc = [("conv1p1s2", "bn1", "block1"), ("conv2p2s2", "bn2", "block2"),
           ("conv3p4s2", "bn3", "block3"), ("conv4p8s2", "bn4", "block4")]
    for L, (cname, bname, blk) in enumerate(enc):
        x = conv_stride2(x, lv.parent[L], lv.koff[L], sd[cname + ".kernel"], len(lv.coords[L + 1]), dtype)
        x = tap(cname, relu(batchnorm(x, sd, bname, dtype)))
        x = tap(blk, basic_block(x, lv.nbr3[L + 1], sd, blk, dtype))
        skips.append(x)
    dec = [("convtr4p16s2", "bntr4", "block5"), ("convtr5p8s2", "bntr5", "block6"),
           ("convtr6p4s2", "bntr6", "block7"), ("convtr7p2s2", "bntr7", "block8")]
    for j, (cname, bname, blk) in enumerate(dec):
        L = 3 - j  # output level
        x = conv_transpose2(x, lv.parent[L], lv.koff[L], sd[cname + ".kernel"], dtype)
        x = tap(cname, relu(batchnorm(x, sd, bname, dtype)))
        x = np.concatenate([x, skips[L]], axis=1)          # ME.cat(out, skip)  minkunet.py:192
        x = tap(blk, basic_block(x, lv.nbr3[L], sd, blk, dtype))
    w = np.asarray(sd["final.kernel"], dtype=dtype).reshape(x.shape[1], -1)
    return tap("final", x @ w + np.asarray(sd["final.bias"], dtype=dtype).reshape(1, -1))


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def sps_forward(points, voxel_size, sd, dtype=np.float32, taps=None, return_levels=False):
    """SPSModel.forward (models.py:20-30): quantise -> TensorField.sparse() (features = mean of
    the constant 0.5 = 0.5) -> MinkUNet -> slice (F[inverse_mapping]) -> sigmoid."""
    c0, inv = voxelize(points, voxel_size)
    levels = Levels(c0)
    feat0 = np.full((len(c0), 1), 0.5, dtype=dtype)
    logits = unet_forward(levels, feat0, sd, dtype, taps)
    scores = sigmoid(logits[inv, 0].astype(dtype)).astype(dtype)
    if return_levels:
        return scores, levels, inv
    return scores


# --------------------------------------------------------------------------------------
# a12/a13/a14  ROS-path input assembly   (src/sps/datasets/util.py:67-114,163-184)
# --------------------------------------------------------------------------------------
def to_coords(cloud_xyz, ds):
    """util.py:72-75: ``torch.div(xyz, [ds,ds,ds]).int()`` -- fp32 division then TRUNCATION."""
    q = np.asarray(cloud_xyz, np.float32)[:, :3] / np.float32(ds)
    return np.trunc(q).astype(np.int32)


def prune(map_xyz, scan_xyz, ds):
    """util.py:85-114: voxels present in BOTH the map and the scan (ME Union of one-hot
    features, keep F0*F1 == 1), returned as ``coordinates * ds`` in fp32 (voxel corners),
    plus the number of unique scan voxels.  Row order is free (canonicalise to compare)."""
    cm, _ = unique_first(to_coords(map_xyz, ds))
    cs, _ = unique_first(to_coords(scan_xyz, ds))
    if len(cm) == 0 or len(cs) == 0:
        return np.zeros((0, 3), np.float32), len(cs)
    lo, rng = _extent(cm, cs)
    km, ks = _pack(cm, lo, rng), _pack(cs, lo, rng)
    both = cm[np.isin(km, ks)]
    return (both.astype(np.float32) * np.float32(ds)).astype(np.float32), len(cs)


def radius_crop(map_xyz, center, radius):
    """c_ws/src/mapmos/scripts/mapmos_node.py:63-68: map points with Euclidean distance
    <= radius of ``center``, order kept.  The centre comes from a float64 pose matrix, so numpy
    promotes the (fp32) map to float64 for the whole expression."""
    m = np.asarray(map_xyz)[:, :3].astype(np.float64)
    d = np.sqrt(np.sum((m - np.asarray(center, np.float64)) ** 2, axis=1))
    return np.nonzero(d <= float(radius))[0]


def assemble(scan_xyz, submap_xyz, batch_index=0.0):
    """util.py:156-174: [b, x, y, z, t] rows, scan rows (t=1) first, then submap rows (t=0)."""
    s = np.asarray(scan_xyz, np.float32)[:, :3]
    m = np.asarray(submap_xyz, np.float32)[:, :3].reshape(-1, 3)
    xyz = np.vstack([s, m])
    t = np.concatenate([np.full(len(s), SCAN_TIMESTAMP, np.float32), np.full(len(m), MAP_TIMESTAMP, np.float32)])
    b = np.full(len(xyz), batch_index, np.float32)
    return np.hstack([b[:, None], xyz, t[:, None]]).astype(np.float32)


def infer(scan_xyz, submap_xyz, voxel_size, sd, dtype=np.float32):
    """util.py:163-184: scores of the scan rows only (``scores[:len(scan_points)]``)."""
    pts = assemble(scan_xyz, submap_xyz)
    return sps_forward(pts, voxel_size, sd, dtype)[: len(scan_xyz)]


# --------------------------------------------------------------------------------------
# a15  metrics   (util.py:285-299, models.py:84-104)
# --------------------------------------------------------------------------------------
def calculate_metrics(true_labels, predicted_labels):
    """util.py:285-299 verbatim semantics: class 1 = unstable (score >= eps)."""
    t, p = np.asarray(true_labels), np.asarray(predicted_labels)
    tp = int(np.sum((t == 1) & (p == 1)))
    tn = int(np.sum((t == 0) & (p == 0)))
    fp = int(np.sum((t == 0) & (p == 1)))
    fn = int(np.sum((t == 1) & (p == 0)))
    precision = tp / (tp + fp) if (tp + fp) != 0 else 0
    recall = tp / (tp + fn) if (tp + fn) != 0 else 0
    f1 = 2 * (precision * recall) / (precision + recall) if (precision + recall) != 0 else 0
    accuracy = (tp + tn) / (tp + tn + fp + fn)
    diou = tp / (tp + fn + fp)
    return precision, recall, f1, accuracy, diou


def threshold_labels(scores, eps):
    """models.py:97: ``np.where(scores < eps, 0, 1)``."""
    return np.where(np.asarray(scores) < eps, 0, 1)


def predict_step_metrics(scores, labels, t_col, eps):
    """models.py:84-104: scan rows are ``t == 1``; MSE, R2, thresholded P/R/F1/dIoU."""
    scan = np.nonzero(np.asarray(t_col) == 1)[0]
    s = np.asarray(scores, np.float64)[scan]
    g = np.asarray(labels, np.float64)[scan]
    mse = float(np.mean((s - g) ** 2))
    ss_res = float(np.sum((g - s) ** 2))
    ss_tot = float(np.sum((g - g.mean()) ** 2))
    r2 = 1.0 - ss_res / ss_tot if ss_tot > 0 else 0.0
    # models.py:97-98 compares float32 tensors with the Python float epsilon: torch evaluates that in
    # float32 (the scalar is cast to the tensor dtype), so 0.84f is NOT below 0.84
    s32, g32 = np.asarray(scores, np.float32)[scan], np.asarray(labels, np.float32)[scan]
    precision, recall, f1, acc, diou = calculate_metrics(threshold_labels(g32, np.float32(eps)),
                                                         threshold_labels(s32, np.float32(eps)))
    return {"loss": mse, "r2": r2, "precision": precision, "recall": recall, "f1": f1,
            "accuracy": acc, "dIoU": diou}


# --------------------------------------------------------------------------------------
# "next" rows: offline-loader submap selection, 4DMOS and MapMOS forwards
# --------------------------------------------------------------------------------------
def select_closest_points(map_xyz, scan_xyz, radius):
    """BLTDataset.select_closest_points (src/sps/datasets/blt_dataset.py:258-271): for every scan point, in scan
    order, the indices of all map points within ``radius`` (Euclidean, float64), as one list per scan point.
    Brute force over a uniform grid in float64 -- an independent restatement, checked against the reference's own
    call (scipy ``cKDTree.query_ball_tree``) in tests/test_oracle.py."""
    m = np.asarray(map_xyz, np.float64)[:, :3]
    s = np.asarray(scan_xyz, np.float64)[:, :3]
    r = float(radius)
    cells = {}
    for i, c in enumerate(map(tuple, np.floor(m / r).astype(np.int64))):
        cells.setdefault(c, []).append(i)
    out = []
    for p, c in zip(s, np.floor(s / r).astype(np.int64)):
        hits = []
        for dx in (-1, 0, 1):
            for dy in (-1, 0, 1):
                for dz in (-1, 0, 1):
                    for j in cells.get((c[0] + dx, c[1] + dy, c[2] + dz), ()):
                        d = m[j] - p
                        if d[0] * d[0] + d[1] * d[1] + d[2] * d[2] <= r * r:
                            hits.append(j)
        out.append(sorted(hits))
    return out


def mos4d_forward(points, voxel_size, sd, dtype=np.float32):
    """MOS4DNet.forward (c_ws/src/mos4d/scripts/mos4d.py:17-32): the SPS graph with ``out_channels = 3``, features 0.5,
    returns the raw logit of channel 2 per point.  ``points`` [N,5] = (b, x, y, z, t = scan index)."""
    c0, inv = voxelize(points, voxel_size)
    logits = unet_forward(Levels(c0), np.full((len(c0), 1), 0.5, dtype=dtype), sd, dtype)
    return logits[inv, 2].astype(dtype)


def mapmos_forward(coordinates, indices, voxel_size, sd, dtype=np.float32):
    """MapMOSNet.forward (c_ws/src/mapmos/scripts/mapmos.py:59-83): index-normalised point features
    ``1 + (i_max - i) / (i_max - i_min)`` (ones when all equal), ``TensorField.sparse()`` averages them per voxel
    (UNWEIGHTED_AVERAGE), raw logits out.  ``coordinates`` [N,5] = (b, x, y, z, t)."""
    idx = np.asarray(indices, np.float32).reshape(-1)
    i_max, i_min = idx.max(), idx.min()
    feats = np.ones_like(idx) if i_max == i_min else (1 + (i_max - idx) / (i_max - i_min)).astype(np.float32)
    c0, inv = voxelize(coordinates, voxel_size)
    s = np.zeros(len(c0), np.float64)
    cnt = np.zeros(len(c0), np.float64)
    np.add.at(s, inv, feats)
    np.add.at(cnt, inv, 1.0)
    feat0 = (s / cnt).astype(dtype)[:, None]
    return unet_forward(Levels(c0), feat0, sd, dtype)[inv, 0].astype(dtype)


# --------------------------------------------------------------------------------------
# f3  ROS-path scan I/O   (src/sps/datasets/util.py:117-153,187-194; sps_node.py:89-107,146-149)
# --------------------------------------------------------------------------------------
_PF_NUMPY = {1: "i1", 2: "u1", 3: "i2", 4: "u2", 5: "i4", 6: "u4", 7: "f4", 8: "f8"}   # sensor_msgs/PointField datatypes


def pointcloud2_to_array(data: bytes, width, height, point_step, row_step, fields, is_bigendian=False):
    """util.to_numpy (util.py:146-153) without ros_numpy: ``fields`` = [(name, offset, datatype)]; every field is
    read through a numpy structured dtype (what ros_numpy.numpify builds) and assigned into a float32 column."""
    order = ">" if is_bigendian else "<"
    dt = np.dtype({"names": [f[0] for f in fields], "formats": [order + _PF_NUMPY[f[2]] for f in fields],
                   "offsets": [f[1] for f in fields], "itemsize": point_step})
    buf = np.frombuffer(data, dtype=np.uint8)
    rows = [np.frombuffer(buf[r * row_step: r * row_step + width * point_step].tobytes(), dtype=dt) for r in range(height)]
    pc = np.concatenate(rows) if rows else np.zeros(0, dt)
    scan = np.zeros((height * width, len(fields)), dtype=np.float32)
    for i, f in enumerate(fields):
        scan[:, i] = np.resize(pc[f[0]], height * width)
    return scan


def transform_point_cloud(point_cloud, transformation_matrix):
    """util.py:187-194 verbatim semantics (float64 homogeneous product), followed by the float32 cast of its only
    caller (sps_node.py:106 ``torch.tensor(scan_tr[:, :3], dtype=torch.float32)``)."""
    pc = np.asarray(point_cloud)
    homogeneous = np.hstack((pc, np.ones((pc.shape[0], 1))))
    transformed = np.dot(homogeneous, np.asarray(transformation_matrix).T)
    return (transformed[:, :3] / transformed[:, 3][:, np.newaxis]).astype(np.float32)


def filter_scan(scan, scores, epsilon):
    """sps_node.py:148: rows of the raw scan whose predicted score is <= epsilon."""
    return np.asarray(scan)[np.asarray(scores) <= epsilon]
