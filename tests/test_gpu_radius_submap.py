"""GPU radius submap selection against scipy's cKDTree.query_ball_tree, the call the reference
makes in BLTDataset.select_closest_points (src/sps/datasets/blt_dataset.py:258-271)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def scipy_lists(map_xyz, scan_xyz, r):
    from scipy.spatial import cKDTree
    return cKDTree(scan_xyz).query_ball_tree(cKDTree(map_xyz), r)


def gpu_lists(map_xyz, scan_xyz, r):
    from sps_b200.datasets import RadiusSubmap
    sub = RadiusSubmap(torch.as_tensor(map_xyz).cuda(), r)
    idx, off = sub.select_closest_points(torch.as_tensor(scan_xyz).cuda(), return_offsets=True)
    idx, off = idx.cpu().numpy(), off.cpu().numpy()
    assert off[0] == 0 and off[-1] == len(idx) and np.all(np.diff(off) >= 0)
    return [idx[off[i]:off[i + 1]] for i in range(len(scan_xyz))], idx


@pytest.mark.parametrize("seed,r", [(0, 0.1), (1, 0.25)])
def test_ball_query_matches_scipy_on_lidar_shaped_data(seed, r):
    from sps_b200 import synth
    world = synth.World(seed)
    base = synth.base_map(world, "hdl-32", n_poses=6, seed=seed, voxel=0.1).astype(np.float32)
    scan = synth.scan(world, "hdl-32", pose=(1.0, -2.0, 0.3), seed=seed + 5)[:, :3].astype(np.float32)
    ref = scipy_lists(base, scan, r)
    got, flat = gpu_lists(base, scan, r)
    assert sum(len(l) for l in ref) == len(flat) and len(flat) > len(scan) // 4
    for i, (a, b) in enumerate(zip(ref, got)):
        assert sorted(a) == sorted(b.tolist()), i
    # the reference's merged index vector, as a multiset per scan point and in scan order
    assert np.array_equal(np.sort(np.concatenate([np.asarray(l, int) for l in ref])), np.sort(flat))


def test_ball_query_edge_cases():
    rng = np.random.default_rng(3)
    r = 0.1
    # clustered points incl. exact duplicates, negative coordinates, cell-boundary coordinates
    base = np.concatenate([rng.uniform(-1, 1, (4000, 3)), np.repeat(rng.uniform(-1, 1, (50, 3)), 3, axis=0),
                           np.round(rng.uniform(-1, 1, (500, 3)) / r) * r]).astype(np.float32)
    scan = np.concatenate([rng.uniform(-1.2, 1.2, (3000, 3)), base[:200], np.full((5, 3), 50.0)]).astype(np.float32)
    ref = scipy_lists(base, scan, r)
    got, flat = gpu_lists(base, scan, r)
    for i, (a, b) in enumerate(zip(ref, got)):
        assert sorted(a) == sorted(b.tolist()), i
    assert all(len(g) == 0 for g in got[-5:])                  # scan points far from the map select nothing
    assert all(len(g) >= 1 for g in got[3000:3200])            # a map point is within r of itself
    # empty scan
    from sps_b200.datasets import RadiusSubmap
    sub = RadiusSubmap(torch.as_tensor(base).cuda(), r)
    assert len(sub.select_closest_points(torch.empty((0, 3), device="cuda"))) == 0
    # small capacity guess -> second pass with the exact size (dense cluster: many hits per scan point)
    dense = rng.normal(0, 0.03, (3000, 3)).astype(np.float32)
    ref = scipy_lists(dense, dense[:500], r)
    got, flat = gpu_lists(dense, dense[:500], r)
    assert len(flat) == sum(len(l) for l in ref) > 4 * 500
    assert all(sorted(a) == sorted(b.tolist()) for a, b in zip(ref, got))


def test_item_through_the_model_matches_host_prepared_batch():
    """make_item (GPU selection) -> SPSModel == the scipy-prepared rows of synth.make_batch (same multiset of rows)."""
    from sps_b200 import synth
    from sps_b200.datasets import RadiusSubmap
    from sps_b200.models import SPSModel
    from oracle import sps_oracle as O
    world = synth.World(2)
    base = synth.base_map(world, "hdl-32", n_poses=6, seed=2, voxel=0.1).astype(np.float32)
    xyz = synth.scan(world, "hdl-32", pose=(0.5, 0.5, 0.1), seed=9).astype(np.float32)
    scan = np.hstack([xyz, np.zeros((len(xyz), 1), np.float32)])                              # x, y, z, label
    host_rows = synth.assemble(xyz, synth.submap_radius(base, xyz, 0.1))
    item = RadiusSubmap(torch.as_tensor(base).cuda(), 0.1).make_item(torch.as_tensor(scan).cuda())
    assert item.shape[0] == len(host_rows) and item.shape[1] == 5
    item5 = torch.hstack([torch.zeros(len(item), 1, device="cuda"), item[:, :4]])          # batch index 0, like collate
    model = SPSModel(0.1)
    model.MinkUNet.load_state_dict({k: torch.as_tensor(v) for k, v in O.make_state_dict(seed=0).items()})
    model = model.cuda().eval()
    a = model(item5)[: len(scan)].cpu().numpy()
    b = model(torch.as_tensor(np.ascontiguousarray(host_rows[:, :5])).cuda())[: len(scan)].cpu().numpy()
    model.check()
    assert np.abs(a - b).max() < 3e-4          # same voxel set; fp16 rows: tile grouping may flip a last bit
