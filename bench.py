#!/usr/bin/env python
"""Benchmark of the SPS inference hot path on B200 (contract: see the task's bench.py section).

A *step* = one collated batch (BASELINE.json configs[1]: OS1-64-like 65 536-point scans, batch 8,
kd-tree-style radius-0.1 m submaps with duplicates, 0.1 m voxels, random-init weights) through
``SPSModel.forward``: voxelise -> kernel maps -> 4-D MinkUNet -> devoxelise + sigmoid.

  value  scans/s, whole job, inputs resident in HBM when the timed region starts
  e2e    the same through the reference-facing call with HOST (pinned) buffers: H2D of the rows,
         forward, D2H of the scores, inside the timed region
  --impl reference   the CPU restatement of MinkowskiEngine's algorithm (oracle/me_cpu.c; the
         reference itself cannot run: ME is not vendored/installable) on all host cores
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SENSOR, BATCH, VOXEL, N_DISTINCT = "os1-64", 8, 0.1, 2
WORKLOAD = "config2: os1-64 (65536 pts/scan) x batch 8 + radius-0.1m submap (duplicates kept), 0.1 m voxels"


def make_batches(rank: int, n_distinct=N_DISTINCT, batch=BATCH):
    """Seeded synthetic batches, rows [N,6] = (b,x,y,z,t,label); different scans per rank."""
    from sps_b200 import synth
    world = synth.World(0)
    map_xyz = synth.base_map(world, SENSOR, n_poses=12, seed=0, voxel=VOXEL)
    return [synth.make_batch(sensor=SENSOR, batch=batch, seed=100 * rank + i + 1, voxel=VOXEL, submap="radius",
                             world=world, map_xyz=map_xyz) for i in range(n_distinct)]


def random_state_dict():
    from oracle import sps_oracle as O  # weight factory only (init distributions of SURVEY §8b)
    return O.make_state_dict(seed=0)


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 200 ms during the timed regions."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm": p["hbm_gbs"], "tensor": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "src": "measured"}
    return {"hbm": 6650.0, "tensor": 1400.0, "src": "fallback"}


def stage_accounting(engine, points, act_bytes=4):
    """Algorithmic bytes / FLOPs per stage of ONE forward (SURVEY.md §8d formulas; `act_bytes` per stored
    activation / weight element: 2 with fp16 rows, 4 with fp32 rows)."""
    lib, h = engine.lib, engine.handle
    import ctypes as C
    from sps_b200.engine import _stream
    n_points = len(points)
    engine.voxelize(points, VOXEL)      # explicit (un-fused) map build so that every table can be counted
    engine.build_maps()

    def pairs(level, kind):
        out = C.c_int64()
        lib.sps_ctx_pair_count(h, level, kind, C.byref(out), _stream())
        return out.value
    V = [engine.count(L) for L in range(5)]
    P3 = [pairs(L, 3) for L in range(5)]
    P5 = pairs(0, 5)
    acc = {}
    cap = 1024
    while cap < 2 * n_points:
        cap *= 2
    acc["vox.clear"] = {"bytes": cap * 16}                       # the open-addressing table, 16-byte slots
    acc["vox.insert"] = {"bytes": n_points * 20 + n_points * 4}   # rows in, slot index out
    acc["vox.rank"] = {"bytes": n_points * 4 + n_points * 4}      # slot index in, block-local rank out
    acc["vox.assign"] = {"bytes": n_points * 4 + n_points * 4 + V[0] * 8}   # slot index in, inverse map + voxel keys out
    for L in range(5):
        acc[f"blocks.L{L}"] = {"bytes": V[L] * 8 + V[L] * 4}
    acc["kmap5.L0"] = {"bytes": V[0] * 20 + 125 * V[0] * 4}
    for L in range(5):
        acc[f"kmap3.L{L}"] = {"bytes": V[L] * 20 + 81 * V[L] * 4}
    for L in range(1, 5):
        acc[f"stride.L{L}"] = {"bytes": V[L - 1] * 20 + V[L] * 20 + V[L - 1] * 4}
    acc["devox_sigmoid"] = {"bytes": n_points * 4 + V[0] * 4 + n_points * 4}

    def conv(name, vin, vout, K, npairs, cin, cout, extra_flops=0, extra_bytes=0):
        acc[name] = {"bytes": act_bytes * (vin * cin + vout * cout + K * cin * cout) + 4 * K * vout + extra_bytes,
                     "flops": 2 * npairs * cin * cout + extra_flops}
    P = (8, 16, 32, 64, 64, 32, 16, 8)
    conv("conv0", V[0], V[0], 125, P5, 1, 8)
    acc["conv0+kmap5"] = {"bytes": V[0] * 20 + 4 * V[0] + 8 * act_bytes * V[0], "flops": 2 * P5 * 8}
    for L in range(4):   # shape sort (key + row index in, row index out) and per-tile slices (present entries x 4 B, twice)
        acc[f"sort.L{L}"] = {"bytes": V[L] * 12 + V[L] * 4}
        acc[f"slices.L{L}"] = {"bytes": 2 * 4 * P3[L] + V[L] * 16}
    c = 8
    for i in range(4):
        L = i + 1
        conv(["conv1p1s2", "conv2p2s2", "conv3p4s2", "conv4p8s2"][i], V[L - 1], V[L], 8, V[L - 1], c, c)
        conv(f"block{L}.conv1", V[L], V[L], 81, P3[L], c, P[i])
        ds = 2 * V[L] * c * P[i] if c != P[i] else 0
        conv(f"block{L}.conv2", V[L], V[L], 81, P3[L], P[i], P[i], ds, act_bytes * V[L] * c)
        c = P[i]
    skip = (32, 16, 8, 8)
    for i in range(4):
        L = 3 - i
        co = P[4 + i]
        conv(["convtr4p16s2", "convtr5p8s2", "convtr6p4s2", "convtr7p2s2"][i], V[L + 1], V[L], 8, V[L], c, co)
        cin = co + skip[i]
        conv(f"block{5 + i}.conv1", V[L], V[L], 81, P3[L], cin, co)
        name = f"block{5 + i}.conv2" + ("+final" if i == 3 else "")
        conv(name, V[L], V[L], 81, P3[L], co, co, 2 * V[L] * cin * co, act_bytes * V[L] * cin)
        c = co
    return acc, V, P3, P5


def profile_pass(engine, net, d_batches, steps):
    """Per-stage CUDA-event durations (events recorded by the library on its launch stream)."""
    import ctypes as C
    import torch
    lib = engine.lib
    lib.sps_profile_enable(1)
    sums, n = {}, 0
    names = C.create_string_buffer(32 * 128)
    ms = (C.c_float * 128)()
    cnt = C.c_int()
    for k in range(steps):
        engine.forward(net, d_batches[k % len(d_batches)], VOXEL)
        torch.cuda.synchronize()
        lib.sps_profile_read(names, ms, 128, C.byref(cnt))
        for i in range(cnt.value):
            nm = names.raw[32 * i:32 * i + 32].split(b"\0")[0].decode()
            sums[nm] = sums.get(nm, 0.0) + ms[i]
        n += 1
    lib.sps_profile_enable(0)
    return {k: v / n for k, v in sums.items()}


def roofline_from(stage_ms, acc, peaks):
    """Roofline of the dominant stage: achieved = algorithmic work / live CUDA-event duration."""
    rows = {}
    for name, t in stage_ms.items():
        a = acc.get(name)
        if not a or t <= 0:
            continue
        gbs = a["bytes"] / (t * 1e-3) / 1e9
        row = {"ms": round(t, 4), "GB/s": round(gbs, 1), "hbm_frac": round(gbs / peaks["hbm"], 4)}
        if "flops" in a:
            tf = a["flops"] / (t * 1e-3) / 1e12
            row.update({"TFLOP/s": round(tf, 3), "tensor_frac": round(tf / peaks["tensor"], 5)})
        rows[name] = row
    # group the repeated kernels: all 3x3x3x3 block convs are one kernel, etc.
    dom = max(rows, key=lambda k: rows[k]["ms"])
    r, a = rows[dom], acc[dom]
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(dom)
    if "tensor_frac" in r and r["tensor_frac"] > r["hbm_frac"]:
        roof = {"bound": "tensor", "achieved": r["TFLOP/s"], "peak": peaks["tensor"], "unit": "TFLOP/s",
                "frac": r["tensor_frac"]}
    else:
        roof = {"bound": "hbm", "achieved": r["GB/s"], "peak": peaks["hbm"], "unit": "GB/s", "frac": r["hbm_frac"]}
    roof.update({"kernel": dom, "kernel_ms": r["ms"], "traffic": traffic, "peak_source": peaks["src"],
                 "algorithmic_bytes": a["bytes"], "algorithmic_flops": a.get("flops")})
    return roof, rows


def cpu_baseline(rows, steps=3, warmup=1, nthreads=0):
    """oracle/me_cpu.c on ONE scan of the batch (bounded sample), all host threads."""
    from oracle import me_cpu
    blob = me_cpu.pack_weights(random_state_dict())
    one = np.ascontiguousarray(rows[rows[:, 0] == 0][:, :5])
    for _ in range(warmup):
        me_cpu.forward(one, VOXEL, blob, nthreads)
    t0 = time.perf_counter()
    for _ in range(steps):
        me_cpu.forward(one, VOXEL, blob, nthreads)
    dt = (time.perf_counter() - t0) / steps
    return {"value": 1.0 / dt, "unit": "scans/s", "cores": nthreads or me_cpu.max_threads(), "kind": "port",
            "sample": f"1 scan of the batch ({len(one)} rows: 65536 scan pts + submap), {steps} timed forwards, "
                      f"{dt * 1e3:.0f} ms each; C/OpenMP restatement of ME's CPU algorithm (ME itself not installable)",
            "host_cpus": os.cpu_count()}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rows = make_batches(0, n_distinct=1, batch=1)[0]
    from oracle import me_cpu
    blob = me_cpu.pack_weights(random_state_dict())
    one = np.ascontiguousarray(rows[:, :5])
    for _ in range(args.warmup):
        me_cpu.forward(one, VOXEL, blob)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        me_cpu.forward(one, VOXEL, blob)
    dt = time.perf_counter() - t0
    value = args.steps / dt
    cores = me_cpu.max_threads()
    sample = (f"each step = 1 scan of the batch-8 workload ({len(one)} rows), all {cores} host threads; "
              "C/OpenMP restatement of MinkowskiEngine's CPU algorithm (reference not installable offline)")
    print(json.dumps({
        "impl": "reference", "metric": "scans/s", "value": value, "unit": "scans/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": value, "unit": "scans/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--backend", type=int, default=0,
                    help="0 auto (tcgen05, fp16 rows), 1 fp32 CUDA-core, 2 tcgen05 TF32 on fp32 rows, 3 = 0")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--lanes", type=int, default=3, help="engine contexts/streams forward_async alternates between")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from sps_b200 import _cabi
    from sps_b200.engine import Engine, Net
    from sps_b200.models import SPSModel

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = _cabi.load()
    lib.sps_set_conv_backend(args.backend)

    batches = make_batches(rank)
    host = [torch.as_tensor(np.ascontiguousarray(b[:, :5])).pin_memory() for b in batches]
    dev = [h.cuda() for h in host]
    n_max = max(len(h) for h in host)
    sd = random_state_dict()
    model = SPSModel(VOXEL, max_points=n_max)
    model.MinkUNet.load_state_dict({k: torch.as_tensor(v) for k, v in sd.items()})
    model = model.cuda().eval()
    engine, net = model._prepare(n_max, torch.device("cuda", local))
    out_dev = torch.empty(n_max, dtype=torch.float32, device="cuda")
    gather = [torch.empty_like(out_dev) for _ in range(world)] if world > 1 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    pending_dev = []

    def step_device(k):
        # inputs resident in HBM; consecutive steps alternate between the model's lanes (engine context +
        # stream each), every step's scores are waited for inside the timed region
        pending_dev.append(model.forward_async(dev[k % len(dev)]))
        if len(pending_dev) >= model.lanes:
            out = pending_dev.pop(0).result()
            if world > 1:  # NCCL only gathers predictions (north star): padded scores of every rank
                out_dev[: len(out)].copy_(out)
                dist.all_gather(gather, out_dev)

    def drain_dev():
        while pending_dev:
            pending_dev.pop(0).result()

    pending = []

    def step_host(k):
        # public host-side API, two calls in flight: H2D (pinned) of step k+1 overlaps the kernels of step k;
        # every step's scores are brought back to pinned host memory and waited for inside the timed region
        pending.append(model.forward_async(host[k % len(host)]))
        if len(pending) >= model.lanes:
            pending.pop(0).result()
        if world > 1:
            dist.all_gather(gather, out_dev)

    def drain():
        while pending:
            pending.pop(0).result()

    def timed(fn, steps, finish=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(steps):
            fn(k)
        if finish:
            finish()
        e1.record()
        torch.cuda.synchronize()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local)
    model.lanes = args.lanes
    for k in range(max(args.warmup, 3)):
        step_device(k)
    drain_dev()
    engine.status()
    sampler.start()
    ms_dev = timed(step_device, args.steps, finish=drain_dev)
    for k in range(max(args.warmup, 3)):
        step_host(k)
    drain()
    ms_host = timed(step_host, args.steps, finish=drain)
    clocks = sampler.stop()
    launches = engine.launch_count() * args.steps

    scans = world * BATCH * args.steps
    value = scans / (ms_dev * 1e-3)
    e2e = scans / (ms_host * 1e-3)
    pts_per_scan = 65536
    result = {
        "metric": "scans/s", "value": value, "unit": "scans/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": {0: "f16", 1: "f32", 2: "tf32", 3: "f16"}[args.backend],
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "arithmetic": {0: "fp16 operands and stored activations, fp32 accumulate and epilogue",
                                                         1: "fp32 CUDA cores", 2: "TF32 operands on fp32 rows, fp32 accumulate",
                                                         3: "fp16 operands and stored activations, fp32 accumulate and epilogue"}[args.backend],
                   "rows_per_step": int(np.mean([len(h) for h in host])),
                   "scans_per_step_per_gpu": BATCH, "weights": "random-init (seed 0), BN eval fresh stats",
                   "l2": f"per-step working set (kernel maps + features, several GB) exceeds the 126 MB L2; "
                         f"{len(host)} distinct batches rotate",
                   "conv_backend": args.backend, "lanes": args.lanes, "sharding": "scan-sharded, replicated weights, NCCL all_gather of scores"},
        "mpoints_per_s": value * pts_per_scan / 1e6,
        "e2e": {"value": e2e, "unit": "scans/s", "ms_per_step": ms_host / args.steps,
                "h2d_bytes_per_step": int(np.mean([h.numel() * 4 for h in host])),
                "d2h_bytes_per_step": int(np.mean([len(h) * 4 for h in host]))},
        "gpu_launches": launches, "clocks": clocks,
    }
    if rank == 0:
        peaks = load_peaks()
        stage_ms = profile_pass(engine, net, dev, min(args.steps, 5))
        acc, V, P3, P5 = stage_accounting(engine, dev[(min(args.steps, 5) - 1) % len(dev)],
                                          act_bytes=2 if args.backend in (0, 3) else 4)
        roof, rows = roofline_from(stage_ms, acc, peaks)
        result["roofline"] = roof
        result["stages"] = rows
        result["sizes"] = {"voxels_per_level": V, "pairs3": P3, "pairs5": P5,
                           "flops_per_step": int(sum(a.get("flops", 0) for a in acc.values()))}
        if world == 1 and not args.no_cpu_baseline:
            result["cpu_baseline"] = cpu_baseline(batches[0])
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(result))


if __name__ == "__main__":
    main()
