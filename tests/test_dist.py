"""World-size-2 gloo test (CPU) of the scan-sharding / gather / metric-aggregation host logic
(BASELINE.json configs[3]).  The GPU forward is replaced by a deterministic stand-in
score function: this test is about the collective plumbing, not the kernels."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import sps_oracle as O

EPS = 0.84
N_SCANS = 7


def make_scan(i):
    rng = np.random.default_rng(i)
    ns, nm = 50 + 7 * i, 20 + i
    rows = np.zeros((ns + nm, 6), np.float32)
    rows[:, 1:4] = rng.uniform(-5, 5, (ns + nm, 3))
    rows[:ns, 4] = 1
    rows[:ns, 5] = rng.uniform(0, 1, ns)
    rows[ns:, 5] = 1
    return torch.as_tensor(rows)


def fake_scores(rows):
    return torch.sigmoid(3 * torch.sin(rows[:, 1] * 1.7 + rows[:, 2]) + 1.5)


def cpu_partials(scores, rows, eps):
    scan = rows[:, 4] == 1
    s, g = scores[scan].double(), rows[scan, 5].double()
    p, t = (s >= eps).long(), (g >= eps).long()
    counts = torch.tensor([int(((t == 1) & (p == 1)).sum()), int(((t == 0) & (p == 0)).sum()),
                           int(((t == 0) & (p == 1)).sum()), int(((t == 1) & (p == 0)).sum())], dtype=torch.int64)
    sums = torch.tensor([float(len(s)), float(((s - g) ** 2).sum()), float(g.sum()), float((g * g).sum()), float(s.sum())],
                        dtype=torch.float64)
    return counts, sums


def worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sps_b200.parallel import ShardedPredictor, shard_ids
    assert shard_ids(N_SCANS, rank, world) == list(range(rank, N_SCANS, world))
    scans = [make_scan(i) for i in range(N_SCANS)]
    res = ShardedPredictor(fake_scores, EPS, partials_fn=cpu_partials).predict(scans)
    if rank == 0:
        out["metrics"] = res["metrics"]
        out["scores"] = [s.numpy().copy() for s in res["scores"]]
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_predict_world2_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(worker, args=(2, port, out), nprocs=2, join=True)
    # single-process expectation with the oracle's metric formulas (mean of per-scan values)
    exp = {k: [] for k in ("loss", "r2", "precision", "recall", "f1", "dIoU")}
    for i in range(N_SCANS):
        rows = make_scan(i)
        s = fake_scores(rows).numpy()
        m = O.predict_step_metrics(s, rows[:, 5].numpy(), rows[:, 4].numpy(), EPS)
        for k in exp:
            exp[k].append(m[k])
        scan = rows[:, 4].numpy() == 1
        assert np.allclose(out["scores"][i], s[scan])          # predictions gathered in scan-id order
    got = out["metrics"]
    assert abs(got["Loss"] - np.mean(exp["loss"])) < 1e-9 and abs(got["R2"] - np.mean(exp["r2"])) < 1e-9
    assert abs(got["Precision"] - np.mean(exp["precision"])) < 1e-12 and abs(got["Recall"] - np.mean(exp["recall"])) < 1e-12
    assert abs(got["F1"] - np.mean(exp["f1"])) < 1e-12 and abs(got["dIoU"] - np.mean(exp["dIoU"])) < 1e-12


def test_single_process_path_matches():
    from sps_b200.parallel import ShardedPredictor
    scans = [make_scan(i) for i in range(3)]
    res = ShardedPredictor(fake_scores, EPS, partials_fn=cpu_partials).predict(scans)
    assert len(res["scores"]) == 3 and res["counts"].shape == (3, 4)
    assert int(res["counts"].sum()) == sum(int((s[:, 4] == 1).sum()) for s in scans)


# ---------------------------------------------------------------------------------------------------------------
# The same plumbing under NCCL with the real network and the library's metric kernel (BASELINE.json configs[3]).
# Needs two GPUs (NCCL refuses two ranks on one device): skipped on a single-GPU box, run with `gpurun --gpus 2`.
# ---------------------------------------------------------------------------------------------------------------
import pytest


def _nccl_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from conftest import make_case
    from sps_b200.models import SPSNet
    from sps_b200.parallel import ShardedPredictor
    sd = O.make_state_dict(seed=0, randomize_bn=True)
    sd["final.kernel"] = sd["final.kernel"] * np.float32(8.0)      # scores on both sides of eps
    cfg = {"MODEL": {"VOXEL_SIZE": 0.1}, "FILTER": {"THRESHOLD": EPS}}
    net = SPSNet(cfg)
    net.model.MinkUNet.load_state_dict({k: torch.as_tensor(v) for k, v in sd.items()})
    net = net.cuda()
    net.freeze()
    scans = [torch.as_tensor(make_case("tiny", seed=40 + i)).cuda() for i in range(5)]
    res = ShardedPredictor(lambda rows: net(rows), EPS).predict(scans)
    net.model.check()
    if rank == 0:
        out["metrics"] = res["metrics"]
        out["scores"] = [s.cpu().numpy().copy() for s in res["scores"]]
        out["counts"] = res["counts"].numpy().copy()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_sharded_predict_world2_nccl_real_network():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (NCCL: one rank per device)")
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from conftest import make_case
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_nccl_worker, args=(2, port, out), nprocs=2, join=True)
    sd = O.make_state_dict(seed=0, randomize_bn=True)
    sd["final.kernel"] = sd["final.kernel"] * np.float32(8.0)
    exp = {k: [] for k in ("loss", "r2", "precision", "recall", "f1", "dIoU")}
    for i in range(5):
        rows = make_case("tiny", seed=40 + i)
        ref = O.sps_forward(rows[:, :5], 0.1, sd)
        scan = rows[:, 4] == 1
        assert np.abs(out["scores"][i] - ref[scan]).max() < 2e-3       # gathered in scan-id order from both ranks
        m = O.predict_step_metrics(_scores_full(out["scores"][i], scan, ref), rows[:, 5], rows[:, 4], EPS)
        for k in exp:
            exp[k].append(m[k])
    got = out["metrics"]
    for a, b in (("Loss", "loss"), ("R2", "r2"), ("Precision", "precision"), ("Recall", "recall"), ("F1", "f1"), ("dIoU", "dIoU")):
        assert abs(got[a] - np.mean(exp[b])) < 1e-5, (a, got[a], np.mean(exp[b]))
    assert out["counts"].sum() == sum(int((make_case("tiny", seed=40 + i)[:, 4] == 1).sum()) for i in range(5))


def _scores_full(scan_scores, scan_mask, ref):
    """Per-row score vector with the GPU's scan-row scores in place (the metrics only read the scan rows)."""
    full = ref.copy()
    full[scan_mask] = scan_scores
    return full
