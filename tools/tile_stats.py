"""CPU statistics of the 128-row tiles the tensor-core convolution walks (no GPU needed): offsets walked per tile,
fill of the gathered A operand, how many 4/8/32-row groups of a walked offset are entirely empty (copies that could be
skipped), and how many distinct input rows / 128-byte lines a tile touches under different storage orders.

    python tools/tile_stats.py [sensor] [batch]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from oracle import sps_oracle as O  # noqa: E402
from sps_b200 import synth  # noqa: E402


def shape_key(nbr):
    """the 29-bit key of csrc/maps.cu (k_kernel_map_blk3): [has dt=+1][has dt=-1][27-bit spatial presence, OR over t]"""
    pres = (nbr >= 0)
    V = nbr.shape[1]
    m = [np.zeros(V, np.int64) for _ in range(3)]
    for it in range(3):
        for k3 in range(27):
            m[it] |= pres[it * 27 + k3].astype(np.int64) << k3
    pat = m[0] | m[1] | m[2]
    return pat | ((m[0] != 0).astype(np.int64) << 27) | ((m[2] != 0).astype(np.int64) << 28)


def morton_order(coords, s):
    c = coords.astype(np.int64)
    x, y, z = (c[:, 1] // s), (c[:, 2] // s), (c[:, 3] // s)
    x -= x.min(); y -= y.min(); z -= z.min()
    key = np.zeros(len(c), np.int64)
    for b in range(16):
        key |= ((x >> b) & 1) << (3 * b) | ((y >> b) & 1) << (3 * b + 1) | ((z >> b) & 1) << (3 * b + 2)
    key |= c[:, 0].astype(np.int64) << 52 | c[:, 4].astype(np.int64) << 48
    return np.argsort(key, kind="stable")


def block_order(coords, s):
    """storage order = (batch, t, 4x4x4 block in first-occurrence... here: sorted block key, cell x fastest)"""
    c = coords.astype(np.int64)
    x, y, z = (c[:, 1] // s), (c[:, 2] // s), (c[:, 3] // s)
    x -= x.min(); y -= y.min(); z -= z.min()
    bk = (((c[:, 0] * 16 + c[:, 4]) * 65536 + (z >> 2)) * 65536 + (y >> 2)) * 65536 + (x >> 2)
    cell = (x & 3) + 4 * (y & 3) + 16 * (z & 3)
    return np.lexsort((cell, bk))


def stats(nbr, order, name, row_bytes_list=(16, 32, 64, 128), storage=None):
    K, V = nbr.shape
    ntiles = (V + 127) // 128
    pres = nbr >= 0
    walked = pairs = 0
    empty4 = empty8 = empty32 = slots4 = slots8 = slots32 = 0
    distinct_rows = 0
    lines = {b: 0 for b in row_bytes_list}
    store_pos = np.arange(V) if storage is None else np.empty(V, np.int64)
    if storage is not None:
        store_pos[storage] = np.arange(V)
    sample = np.linspace(0, ntiles - 1, min(ntiles, 400)).astype(int)
    for t in sample:
        rows = order[t * 128:(t + 1) * 128]
        p = pres[:, rows]                       # [K, <=128]
        act = p.any(axis=1)
        na = int(act.sum())
        walked += na
        pairs += int(p.sum())
        pa = p[act]
        n = pa.shape[1]
        for g, tag in ((4, "4"), (8, "8"), (32, "32")):
            ng = n // g
            if ng == 0:
                continue
            e = (~pa[:, :ng * g].reshape(na, ng, g).any(axis=2)).sum()
            if g == 4:
                empty4 += e; slots4 += na * ng
            elif g == 8:
                empty8 += e; slots8 += na * ng
            else:
                empty32 += e; slots32 += na * ng
        idx = nbr[:, rows][p]
        pos = store_pos[idx]
        distinct_rows += len(np.unique(pos))
        for b in row_bytes_list:
            lines[b] += len(np.unique(pos * b // 128))
    nt = len(sample)
    print(f"  {name:34s} walked/tile {walked / nt:5.1f}  pairs/tile {pairs / nt:7.1f}  fill {pairs / max(walked * 128, 1):.3f}  "
          f"empty 4-row {empty4 / max(slots4, 1):.3f} 8-row {empty8 / max(slots8, 1):.3f} 32-row {empty32 / max(slots32, 1):.3f}  "
          f"distinct rows/tile {distinct_rows / nt:7.1f}  lines/tile " +
          " ".join(f"{b}B:{lines[b] / nt:.0f}" for b in row_bytes_list))


def main():
    sensor = sys.argv[1] if len(sys.argv) > 1 else "os1-64"
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    world = synth.World(0)
    map_xyz = synth.base_map(world, sensor, n_poses=12, seed=0, voxel=0.1)
    rows = synth.make_batch(sensor=sensor, batch=batch, seed=1, voxel=0.1, submap="radius", world=world, map_xyz=map_xyz)
    c0, inv = O.voxelize(rows[:, :5], 0.1)
    coords = [c0]
    for L in range(4):
        c, _ = O.stride_coords(coords[L], 2 ** (L + 1))
        coords.append(c)
    for L in range(4):
        s = 2 ** L
        nbr = O.kernel_map(coords[L], coords[L], O.kernel_offsets([3, 3, 3, 3], [s] * 3 + [1]))
        V = nbr.shape[1]
        print(f"level {L}: {V} voxels, {(nbr >= 0).sum() / V:.2f} pairs/voxel")
        phys = np.arange(V)
        key = shape_key(nbr)
        srt = np.argsort(key, kind="stable")
        stats(nbr, phys, "physical order, physical storage")
        stats(nbr, srt, "shape-sorted, physical storage")
        bo = block_order(coords[L], s)
        stats(nbr, srt, "shape-sorted, block-order storage", storage=bo)
        stats(nbr, bo, "block-order tiles, block storage", storage=bo)
        # shape sort with the popcount as the leading key
        pc = np.array([bin(int(k) & ((1 << 27) - 1)).count("1") for k in key])
        srt2 = np.lexsort((key, pc))
        stats(nbr, srt2, "popcount-major shape sort", storage=None)


if __name__ == "__main__":
    main()
