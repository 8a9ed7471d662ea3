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
        lib.sps_set_conv_backend(backend)
        got = eng.forward(net, d, 0.1).cpu().numpy()
        eng.status()
        lib.sps_set_conv_backend(0)
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
