"""Full-size configurations of BASELINE.json on the GPU: parity against the C/OpenMP oracle where it
finishes in seconds, plus size-independent properties (determinism, row-permutation invariance,
batch independence, score range)."""
import numpy as np
import pytest
import torch

from oracle import sps_oracle as O
from oracle import me_cpu

pytestmark = [pytest.mark.gpu]
EPS = 0.84


def run(pts, sd, backend=0):
    from sps_b200 import engine
    eng = engine.Engine(len(pts))
    eng.set_conv_backend(backend)
    out = eng.forward(engine.Net(sd), torch.as_tensor(np.ascontiguousarray(pts)).cuda(), 0.1 if pts is None else run.voxel)
    eng.status()
    return out.cpu().numpy(), [eng.count(L) for L in range(5)]


run.voxel = 0.1


def check_against_oracle(pts, sd, voxel, tol, straddle=False):
    run.voxel = voxel
    got, counts = run(pts, sd)
    ref, ref_counts, _ = me_cpu.forward(pts, voxel, me_cpu.pack_weights(sd))
    assert counts == ref_counts.tolist()
    assert np.isfinite(got).all() and got.min() >= 0 and got.max() <= 1
    err = np.abs(got - ref)
    assert err.max() < tol, err.max()
    assert np.mean((got < EPS) == (ref < EPS)) >= 0.999
    if straddle:   # the labels must actually be decided by the scores: both classes well populated
        assert 0.2 < np.mean(ref >= EPS) < 0.8 and ref.min() < 0.3 and ref.max() > 0.95
    return got


def spread(sd, pts, voxel=0.1):
    """Head gain x8 + bias moved so that eps = 0.84 cuts the score distribution in the middle (what a trained
    checkpoint's scores look like; the random-init head alone keeps every score below eps)."""
    from test_gpu_parity import spread_state_dict
    return spread_state_dict(sd, pts, 8.0, voxel)


def test_config1_128k_scan_voxel_submap():
    """configs[0]: one 131 072-point scan + 0.1 m voxel-overlap submap, batch 1."""
    from sps_b200 import synth
    world = synth.World(0)
    scan = synth.scan(world, "os1-128", seed=0)
    base = synth.base_map(world, "os1-128", n_poses=10, seed=0)
    sub = synth.submap_voxel_overlap(base, scan, 0.1)
    pts = synth.assemble(scan, sub)[:, :5]
    assert len(scan) == 131072
    sd = O.make_state_dict(seed=0)
    got = check_against_oracle(pts, sd, 0.1, 2e-3)
    check_against_oracle(pts, spread(sd, pts), 0.1, 2e-3, straddle=True)     # same bar, 1x, on spread-out scores
    # determinism: bit-identical on a second run
    again, _ = run(pts, sd)
    assert np.array_equal(got, again)
    # row-permutation invariance: the voxel set, hence every point's score, does not depend on row order.
    # fp32 rows (backend 2): only the grouping of exact zeros changes.  fp16 rows (default): a last-bit fp32
    # difference may flip the rounding of a stored activation (2^-11 relative) -> 5e-5 on the scores.
    perm = np.random.default_rng(0).permutation(len(pts))
    shuffled, _ = run(pts[perm], sd)
    assert np.abs(shuffled - got[perm]).max() < 3e-4
    got32, _ = run(pts, sd, backend=2)
    shuffled32, _ = run(pts[perm], sd, backend=2)
    assert np.abs(shuffled32 - got32[perm]).max() < 1e-6


def test_config2_batch8_radius_submaps_batch_independence():
    """configs[1] shape (the bench workload): batch items never interact."""
    import bench
    rows = bench.make_batches(0, n_distinct=1, batch=8)[0]
    pts = np.ascontiguousarray(rows[:, :5])
    sd = O.make_state_dict(seed=0)
    run.voxel = 0.1
    first = pts[pts[:, 0] == 0].copy()
    for weights in (sd, spread(sd, first)):       # contract weights, then scores spread over (0,1) around eps
        got, counts = run(pts, weights)
        assert np.isfinite(got).all()
        for b in (0, 5):
            one = pts[pts[:, 0] == b].copy()
            one[:, 0] = 0
            ref, c1, _ = me_cpu.forward(one, 0.1, me_cpu.pack_weights(weights))
            sel = got[pts[:, 0] == b]
            assert np.abs(sel - ref).max() < 2e-3
            assert np.mean((sel < EPS) == (ref < EPS)) >= 0.999
            if weights is not sd:
                assert 0.1 < np.mean(ref >= EPS) < 0.9 and ref.min() < 0.3 and ref.max() > 0.95


def test_config5_stress_dense_scan_small_voxels_radius_crop():
    """configs[4] shape: 524 288-point scan, 0.05 m voxels, 30 m radius crop of the base map."""
    from sps_b200 import synth
    from sps_b200.engine import MapHash
    world = synth.World(3)
    scan = synth.scan(world, "dense-128x4096", pose=(2.0, -1.0, 0.4), seed=3)
    base = synth.base_map(world, "os1-128", n_poses=12, seed=3, voxel=0.05)
    mh = MapHash(torch.as_tensor(base).cuda(), 0.05)
    idx = mh.crop_radius((2.0, -1.0, 1.8), 30.0).cpu().numpy()
    assert np.array_equal(idx, O.radius_crop(base, np.array([2.0, -1.0, 1.8]), 30.0))
    pts = synth.assemble(scan, base[idx])[:, :5]
    assert len(scan) == 524288
    sd = O.make_state_dict(seed=1)
    check_against_oracle(pts, sd, 0.05, 2e-3)
