"""CUDA path against the committed golden fixtures (tests/golden/*.npz): canonical coordinate sets
and kernel maps bit-exact (digests), scores within tolerance; plus the device metric partials."""
import json
import os

import numpy as np
import pytest
import torch

from golden.make_golden import CASES, case_points, digest
from oracle import sps_oracle as O

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_path_reproduces_golden(name):
    from sps_b200 import engine, _cabi
    lib = _cabi.load()
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    rows = case_points(name)
    pts = rows[:, :5]
    assert np.array_equal(digest(pts), g["points_digest"]), "seeded generator drifted"
    sd = O.make_state_dict(seed=0, randomize_bn=True)
    net = engine.Net(sd)
    eng = engine.Engine(len(pts))
    d = torch.as_tensor(pts).cuda()
    eng.voxelize(d, 0.1)
    eng.build_maps()
    eng.status()
    assert np.array_equal(digest(eng.inverse_map().astype(np.int64)), g["inverse_digest"])
    for L in range(5):
        c = eng.coords(L)
        canon = O.canonical(c)
        assert len(canon) == int(g[f"coords{L}_count"][0])
        assert np.array_equal(digest(canon.astype(np.int32)), g[f"coords{L}_digest"])
        assert np.array_equal(canon[:8], g[f"coords{L}_head"])
        km = O.canonical_kernel_map(eng.kernel_map(L, "3"), c, c)
        assert len(km) == int(g[f"kmap3_{L}_pairs"][0])
        assert np.array_equal(digest(km.astype(np.int64)), g[f"kmap3_{L}_digest"])
    c0 = eng.coords(0)
    km5 = O.canonical_kernel_map(eng.kernel_map(0, "5"), c0, c0)
    assert len(km5) == int(g["kmap5_pairs"][0]) and np.array_equal(digest(km5.astype(np.int64)), g["kmap5_digest"])
    for backend, tol in ((1, 2e-5), (0, 2e-3)):
        eng.set_conv_backend(backend)
        got = eng.forward(net, d, 0.1).cpu().numpy()
        eng.status()
        eng.set_conv_backend(0)
        assert np.abs(got - g["scores"]).max() < tol, (backend, np.abs(got - g["scores"]).max())


def test_device_metric_partials_match_reference_formulas():
    from sps_b200.parallel import device_partials, metrics_from_partials
    rng = np.random.default_rng(0)
    n = 5000
    rows = np.zeros((n, 6), np.float32)
    rows[:, 4] = rng.integers(0, 2, n)
    rows[:, 5] = rng.uniform(0, 1, n)
    scores = rng.uniform(0, 1, n).astype(np.float32)
    scores[:10] = 0.84
    c, s = device_partials(torch.as_tensor(scores).cuda(), torch.as_tensor(rows).cuda(), 0.84)
    m = O.predict_step_metrics(scores, rows[:, 5], rows[:, 4], 0.84)
    got = metrics_from_partials(c.cpu()[None], s.cpu()[None])
    assert abs(got["Loss"] - m["loss"]) < 1e-9 and abs(got["R2"] - m["r2"]) < 1e-9
    assert abs(got["Precision"] - m["precision"]) < 1e-12 and abs(got["Recall"] - m["recall"]) < 1e-12
    assert abs(got["F1"] - m["f1"]) < 1e-12 and abs(got["dIoU"] - m["dIoU"]) < 1e-12
    assert int(c.sum()) == int((rows[:, 4] == 1).sum())


def test_next_rows_reproduce_golden():
    """Radius submap selection against lists produced by the reference's own call (scipy), MOS4DNet / MapMOSNet against
    the oracle's logits (tests/golden/next_rows_s11.npz)."""
    from golden.make_golden import next_rows_inputs, next_rows_state_dicts
    from sps_b200 import _cabi
    from sps_b200.datasets import RadiusSubmap
    from sps_b200.models import MOS4DNet, MapMOSNet
    g = np.load(os.path.join(GOLDEN, "next_rows_s11.npz"))
    base, scans, window, map_idx = next_rows_inputs()
    assert np.array_equal(digest(np.concatenate([base.ravel(), window.ravel(), map_idx])), g["inputs_digest"]), "generator drifted"
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a)).cuda()
    idx, off = RadiusSubmap(t(base), 0.1).select_closest_points(t(scans[0]), return_offsets=True)
    idx, off = idx.cpu().numpy(), off.cpu().numpy()
    assert np.array_equal(np.diff(off).astype(np.int32), g["ball_counts"]) and len(idx) == int(g["ball_total"][0])
    flat = np.concatenate([np.sort(idx[off[i]:off[i + 1]]) for i in range(len(scans[0]))]).astype(np.int64)
    assert np.array_equal(digest(flat), g["ball_sorted_digest"])
    sd, sd3 = next_rows_state_dicts()
    lib = _cabi.load()
    mos = MOS4DNet(0.1)
    mos.MinkUNet.load_state_dict({k: torch.as_tensor(v) for k, v in sd3.items()})
    mos = mos.cuda().eval()
    mm = MapMOSNet(0.1)
    mm.MinkUNet.load_state_dict({k: torch.as_tensor(v) for k, v in sd.items()})
    mm = mm.cuda().eval()
    for backend, tol in ((1, 2e-4), (0, 2e-2)):
        mos.set_conv_backend(backend)
        mm.set_conv_backend(backend)
        a = mos(t(window)).cpu().numpy()
        ls, lm = mm.predict(t(scans[0]), t(base), t(np.full(len(scans[0]), 6.0, np.float32)), t(map_idx))
        b = np.concatenate([ls.cpu().numpy(), lm.cpu().numpy()])
        mos.check(); mm.check()
        for got, ref in ((a, g["mos4d_logits"]), (b, g["mapmos_logits"])):
            assert got.shape == ref.shape
            assert np.abs(got - ref).max() < tol * max(1.0, np.abs(ref).max()), (backend, np.abs(got - ref).max())
