import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "tensor_path: run with the default conv dispatch (tcgen05 TF32 kernels)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def make_case(sensor="tiny", seed=0, voxel=0.1, n_map_poses=4, submap="voxel", batch=1):
    """Seeded synthetic input rows [N,6] = (b,x,y,z,t,label)."""
    from sps_b200 import synth
    return synth.make_batch(sensor=sensor, batch=batch, seed=seed, voxel=voxel, submap=submap,
                            n_map_poses=n_map_poses)


@pytest.fixture(scope="session")
def tiny_case():
    return make_case("tiny", seed=3)


@pytest.fixture(scope="session")
def state_dict():
    from oracle import sps_oracle as O
    return O.make_state_dict(seed=0, randomize_bn=True)
