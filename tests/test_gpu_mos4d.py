"""SURVEY 8f rank 2 (first half): integer-t multi-scan windows and out_channels = 3 -- the 4DMOS baseline the reference
ships (c_ws/src/mos4d/scripts/mos4d.py:10-32, mos4d_node.py:98-117), through the same engine."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def window(n_scans, t0, seed=0):
    """A sliding window as mos4d_node.py builds it: rows (b=0, x, y, z, t = scan index)."""
    from sps_b200 import synth
    world = synth.World(seed)
    rows = []
    for i in range(n_scans):
        xyz = synth.scan(world, "tiny", pose=(0.4 * i, -0.2 * i, 0.05 * i), seed=seed + i).astype(np.float32)
        rows.append(np.hstack([np.zeros((len(xyz), 1), np.float32), xyz, np.full((len(xyz), 1), t0 + i, np.float32)]))
    return np.vstack(rows)


def state_dict3(seed=0):
    from oracle import sps_oracle as O
    sd = O.make_state_dict(seed=seed, randomize_bn=True)
    rng = np.random.default_rng(seed + 77)
    sd["final.kernel"] = (rng.standard_normal((8, 3)) * np.sqrt(2.0 / 3)).astype(np.float32)
    sd["final.bias"] = rng.uniform(-0.3, 0.3, (1, 3)).astype(np.float32)
    return sd


@pytest.mark.parametrize("n_scans,t0", [(4, 0), (10, 1000)])
def test_mos4dnet_matches_oracle(n_scans, t0):
    from oracle import sps_oracle as O
    from sps_b200.models import MOS4DNet
    from sps_b200 import _cabi
    pts = window(n_scans, t0)
    sd = state_dict3()
    model = MOS4DNet(0.1)
    model.MinkUNet.load_state_dict({k: torch.as_tensor(v) for k, v in sd.items()})
    model = model.cuda().eval()
    # oracle: the reference's graph on the window shifted to t = 0 (translation in t leaves ME's maps unchanged)
    shifted = pts.copy()
    shifted[:, 4] -= t0
    c0, _ = O.voxelize(shifted, 0.1)
    assert len(np.unique(c0[:, 4])) == n_scans                       # a true 4-D input: one time plane per scan
    ref = O.mos4d_forward(shifted, 0.1, sd)                          # mos4d.py:32  out.features[:, 2]
    lib = _cabi.load()
    for backend, tol in ((1, 2e-4), (0, 2e-2)):
        model.set_conv_backend(backend)
        got = model(torch.as_tensor(pts).cuda()).cpu().numpy()
        model.check()
        assert got.shape == ref.shape
        scale = max(1.0, np.abs(ref).max())
        assert np.abs(got - ref).max() < tol * scale, (backend, np.abs(got - ref).max(), scale)
        assert np.mean((got > 0) == (ref > 0)) >= 0.995              # mos4d_node.py:114: moving iff logit > 0


def test_window_longer_than_16_scans_is_rejected():
    from sps_b200.models import MOS4DNet
    from sps_b200._cabi import SpsError
    pts = window(17, 5)
    model = MOS4DNet(0.1).cuda().eval()
    with pytest.raises(SpsError):
        model(torch.as_tensor(pts).cuda())
        model.check()


def test_mapmosnet_matches_oracle():
    """MapMOSNet (c_ws/src/mapmos/scripts/mapmos.py:32-90): per-point index features averaged per voxel, t = 0 / -1,
    raw logits -- through sps_forward_features."""
    from oracle import sps_oracle as O
    from sps_b200 import synth, _cabi
    from sps_b200.models import MapMOSNet
    world = synth.World(5)
    scan = synth.scan(world, "tiny", pose=(0.2, 0.1, 0.0), seed=3).astype(np.float32)
    mp = synth.base_map(world, "tiny", n_poses=5, seed=5).astype(np.float32)[:6000]
    rng = np.random.default_rng(1)
    scan_idx = np.full(len(scan), 7.0, np.float32)
    map_idx = rng.integers(0, 7, len(mp)).astype(np.float32)
    sd = O.make_state_dict(seed=3, randomize_bn=True)
    model = MapMOSNet(0.1)
    model.MinkUNet.load_state_dict({k: torch.as_tensor(v) for k, v in sd.items()})
    model = model.cuda().eval()
    # oracle (mapmos.py:59-83 on the shifted time axis)
    coords = np.vstack([np.hstack([np.zeros((len(scan), 1), np.float32), scan, np.ones((len(scan), 1), np.float32)]),
                        np.hstack([np.zeros((len(mp), 1), np.float32), mp, np.zeros((len(mp), 1), np.float32)])])
    ref = O.mapmos_forward(coords, np.concatenate([scan_idx, map_idx]), 0.1, sd)
    lib = _cabi.load()
    t = lambda a: torch.as_tensor(a).cuda()
    for backend, tol in ((1, 2e-4), (0, 2e-2)):
        model.set_conv_backend(backend)
        ls, lm = model.predict(t(scan), t(mp), t(scan_idx), t(map_idx))
        model.check()
        got = np.concatenate([ls.cpu().numpy(), lm.cpu().numpy()])
        scale = max(1.0, np.abs(ref).max())
        assert np.abs(got - ref).max() < tol * scale, (backend, np.abs(got - ref).max(), scale)
        assert np.mean((got > 0) == (ref > 0)) >= 0.995
