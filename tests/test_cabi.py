"""CPU-side checks of the boundary: the C-ABI library builds for sm_100a, loads, and exports
every symbol include/sps_b200.h declares; host logic that needs no GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "sps_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sps_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from sps_b200 import _cabi
    lib = _cabi.load()
    names = header_symbols()
    assert len(names) >= 30
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/sps_b200.h but not exported"
    assert sorted(_cabi.SYMBOLS) == names, "ctypes binding and header disagree"
    assert b"sm_100a" in lib.sps_version()


def test_struct_layouts_match_header():
    from sps_b200 import _cabi
    # sizes follow from the C declaration order (LP64): 10*8 and the conv argument block
    assert C.sizeof(_cabi.LevelView) == 80
    a = _cabi.ConvArgs
    assert a.mode.offset == 0 and a.map.offset == 16 and a.n_out.offset == 32
    assert a.weight_kmajor.offset == a.head_out.offset + 8 and a.tile_slices.offset == a.perm.offset + 8
    assert a.io_dtype.offset == a.tile_slices.offset + 8 and a.backend.offset == a.io_dtype.offset + 4
    assert a.flags.offset == a.backend.offset + 4 and a.flags.offset + 8 == C.sizeof(a)   # 3 ints + tail padding


def test_argument_validation_without_gpu():
    """Bad arguments are rejected before any CUDA call is made."""
    from sps_b200 import _cabi
    lib = _cabi.load()
    assert lib.sps_workspace_bytes(1000) > 0
    assert lib.sps_workspace_bytes(2000) > lib.sps_workspace_bytes(1000)
    assert lib.sps_map_bytes(1000) > 0
    h = C.c_void_p()
    assert lib.sps_ctx_create(C.byref(h), None, 0, 100) == _cabi.SPS_ERR_BAD_ARG
    assert lib.sps_voxelize(None, None, 0, 5, 0.1, None) == _cabi.SPS_ERR_BAD_ARG
    assert lib.sps_conv_fwd(None, None) == _cabi.SPS_ERR_BAD_ARG
    assert lib.sps_ctx_set_conv_backend(None, 0) == _cabi.SPS_ERR_BAD_ARG      # settings live in a context: no process-wide switch
    assert lib.sps_ctx_set_pattern_sort(None, 1) == _cabi.SPS_ERR_BAD_ARG
    net = C.c_void_p()
    assert lib.sps_net_create(C.byref(net)) == 0
    x = np.zeros(8, np.float32)
    assert lib.sps_net_set_tensor(net, b"final.kernel", x.ctypes.data_as(C.c_void_p), 8) == 0
    # finalize without the other tensors must fail loudly (needs a device pointer argument first)
    assert lib.sps_net_finalize(net, None, 0, None) == _cabi.SPS_ERR_BAD_ARG
    assert lib.sps_net_destroy(net) == 0


def test_ballmap_argument_validation_without_gpu():
    from sps_b200 import _cabi
    lib = _cabi.load()
    assert lib.sps_ballmap_bytes(1000) > 0 and lib.sps_ballmap_bytes(100000) > lib.sps_ballmap_bytes(1000)
    assert lib.sps_ball_query_scratch_bytes(1000) >= 2 * 4 * 1000
    h = C.c_void_p()
    assert lib.sps_ballmap_build(C.byref(h), None, 0, None, 10, 0.1, None) == _cabi.SPS_ERR_BAD_ARG
    assert lib.sps_submap_ball_query(None, None, 0, None, None, 0, None, None, 0, None) == _cabi.SPS_ERR_BAD_ARG


def test_key_packing_contract():
    """The documented coordinate range of the 64-bit voxel key."""
    text = open(os.path.join(ROOT, "include", "sps_b200.h")).read()
    assert "#define SPS_X_BIAS 131072" in text and "#define SPS_Z_BIAS 32768" in text


def test_state_dict_layout_matches_reference_checkpoint_contract():
    """key names and shapes of CustomMinkUNet(1,1,D=4); Lightning prefix stripped
    as in src/sps/datasets/util.py:33-39."""
    from sps_b200.models import CustomMinkUNet, SPSNet
    from oracle import sps_oracle as O
    net = CustomMinkUNet()
    sd = net.state_dict()
    exp = O.make_state_dict(0)
    assert set(sd) == set(exp)
    for k in sd:
        assert tuple(sd[k].shape) == tuple(np.shape(exp[k])), k
    assert tuple(sd["conv0p1s1.kernel"].shape) == (125, 1, 8)
    assert tuple(sd["block5.0.conv1.kernel"].shape) == (81, 96, 64)
    assert tuple(sd["block5.0.downsample.0.kernel"].shape) == (96, 64)
    assert tuple(sd["final.kernel"].shape) == (8, 1) and tuple(sd["final.bias"].shape) == (1, 1)
    n_conv = sum(v.numel() for k, v in sd.items() if k.endswith("kernel"))
    assert n_conv == 1845168
    # a Lightning checkpoint round trip through the reference's prefix stripping
    cfg = {"MODEL": {"VOXEL_SIZE": 0.1}, "FILTER": {"THRESHOLD": 0.84}}
    model = SPSNet(cfg)
    ckpt = {"model.MinkUNet." + k: torch.as_tensor(v) for k, v in exp.items()}
    ckpt["model.MOSLoss.weight"] = torch.zeros(1)
    stripped = {k.replace("model.MinkUNet.", ""): v for k, v in ckpt.items()}
    stripped = {k: v for k, v in stripped.items() if "MOSLoss" not in k}
    v0 = model.model.MinkUNet.weights_version
    model.model.MinkUNet.load_state_dict(stripped)
    assert model.model.MinkUNet.weights_version > v0        # weights get re-folded on next forward
    assert torch.equal(model.model.MinkUNet.state_dict()["final.bias"], torch.as_tensor(exp["final.bias"]))


def test_random_init_distributions():
    """resnet.py:87-94 + ME defaults."""
    from sps_b200.models import CustomMinkUNet
    torch.manual_seed(0)
    net = CustomMinkUNet()
    k = net.block5[0].conv1.kernel
    assert abs(k.std().item() - np.sqrt(2.0 / (81 * 64))) < 2e-3
    t = net.convtr4p16s2.kernel
    s = 1.0 / np.sqrt(64 * 8)
    assert t.abs().max().item() <= s + 1e-7 and t.abs().max().item() > 0.9 * s
    assert torch.all(net.bn0.bn.weight == 1) and torch.all(net.bn0.bn.bias == 0)


def test_no_cpu_fallback():
    from sps_b200.models import SPSModel
    m = SPSModel(0.1).eval()
    with pytest.raises(RuntimeError):
        m(torch.zeros(4, 5))         # model on CPU: refused, not silently computed
    with pytest.raises(RuntimeError):
        SPSModel(0.1).train()(torch.zeros(4, 5))
    from sps_b200.engine import Engine
    with pytest.raises(RuntimeError):
        Engine(100, device="cpu")


def test_next_row_mirrors_refuse_cpu_and_keep_the_reference_layout():
    """RadiusSubmap / MOS4DNet / MapMOSNet: same rules as the hot path -- no CPU fallback; state_dict keys and shapes of
    the baselines the reference ships (mos4d.py:15 CustomMinkUNet(1, 3, D=4); mapmos.py:37 CustomMinkUNet14(1, 1, D=4))."""
    from sps_b200.datasets import RadiusSubmap
    from sps_b200.models import MOS4DNet, MapMOSNet, SPSModel
    with pytest.raises(RuntimeError):
        RadiusSubmap(torch.zeros(10, 3), 0.1)
    mos, mapmos, sps = MOS4DNet(0.1).eval(), MapMOSNet(0.1).eval(), SPSModel(0.1)
    assert tuple(mos.MinkUNet.state_dict()["final.kernel"].shape) == (8, 3)
    assert tuple(mos.MinkUNet.state_dict()["final.bias"].shape) == (1, 3)
    assert (mos.output_channel, mos.apply_sigmoid) == (2, False)
    assert (mapmos.output_channel, mapmos.apply_sigmoid) == (0, False)
    assert (sps.output_channel, sps.apply_sigmoid) == (0, True)
    keys = lambda m: {k: tuple(v.shape) for k, v in m.MinkUNet.state_dict().items() if not k.startswith("final")}
    assert keys(mos) == keys(sps) == keys(mapmos)
    with pytest.raises(RuntimeError):
        mos(torch.zeros(4, 5))
    with pytest.raises(RuntimeError):
        mapmos(torch.zeros(4, 5), torch.zeros(4))


def test_kmajor_weight_packers_host_side():
    """sps_conv_pack_kmajor / _f16 are host functions: K-major rows [cout][ld], per kernel offset a group-padded channel
    run (4 fp32 / 8 fp16 channels per 16-byte group; 1, 2, 4 or 8k groups), then the fused 1x1 term padded to a stage."""
    from sps_b200 import _cabi
    lib = _cabi.load()
    rng = np.random.default_rng(0)
    for K, cin, cout, cin2 in [(81, 8, 8, 0), (81, 16, 8, 16), (81, 24, 16, 24), (81, 48, 32, 48), (81, 96, 64, 96), (8, 8, 8, 0),
                               (8, 64, 32, 0)]:
        w = rng.standard_normal((K, cin, cout)).astype(np.float32)
        w2 = rng.standard_normal((cin2, cout)).astype(np.float32) if cin2 else None
        p2 = w2.ctypes.data_as(C.c_void_p) if cin2 else None
        # fp16
        g = (cin + 7) // 8
        gp = 1 if g <= 1 else 2 if g <= 2 else 4 if g <= 4 else (g + 7) // 8 * 8
        ld = lib.sps_conv_kmajor_ld_f16(K, cin, cin2)
        assert ld == K * gp * 8 + (cin2 + 63) // 64 * 64 and ld % 8 == 0
        out = np.full((cout, ld), 7, np.float16)
        assert lib.sps_conv_pack_kmajor_f16(w.ctypes.data_as(C.c_void_p), K, cin, cout, p2, cin2, out.ctypes.data_as(C.c_void_p)) == 0
        blk = out[:, : K * gp * 8].reshape(cout, K, gp * 8)
        assert np.array_equal(blk[:, :, :cin], w.astype(np.float16).transpose(2, 0, 1))
        assert not blk[:, :, cin:].any()
        tail = out[:, K * gp * 8:]
        if cin2:
            assert np.array_equal(tail[:, :cin2], w2.astype(np.float16).T) and not tail[:, cin2:].any()
        # fp32 / TF32 (round to nearest even on 13 dropped bits)
        g4 = (cin + 3) // 4
        gp4 = 2 if g4 <= 2 else 4 if g4 <= 4 else (g4 + 7) // 8 * 8
        ld4 = lib.sps_conv_kmajor_ld(K, cin, cin2)
        assert ld4 == K * gp4 * 4 + (cin2 + 31) // 32 * 32
        out4 = np.full((cout, ld4), 7, np.float32)
        assert lib.sps_conv_pack_kmajor(w.ctypes.data_as(C.c_void_p), K, cin, cout, p2, cin2, out4.ctypes.data_as(C.c_void_p)) == 0
        blk4 = out4[:, : K * gp4 * 4].reshape(cout, K, gp4 * 4)
        ref = w.transpose(2, 0, 1)
        assert (blk4[:, :, :cin].view(np.uint32) & 0x1FFF == 0).all()                  # TF32: low 13 mantissa bits cleared
        assert np.abs(blk4[:, :, :cin] - ref).max() <= np.abs(ref).max() * 2.0 ** -11
        assert not blk4[:, :, cin:].any()
    assert lib.sps_conv_pack_kmajor_f16(None, 81, 8, 8, None, 0, None) == _cabi.SPS_ERR_BAD_ARG


def test_split_precision_weight_packer_host_side():
    """sps_conv_pack_kmajor_f16x: hi|lo input rows duplicate every 8-channel group of the weights along K; folded low
    parts appear as rows 8..15, and hi + lo reproduces the fp32 weight to 2^-22 relative."""
    from sps_b200 import _cabi
    lib = _cabi.load()
    rng = np.random.default_rng(1)
    K, cin, cin2 = 81, 16, 16
    w = rng.standard_normal((K, cin, 8)).astype(np.float32)
    w2 = rng.standard_normal((cin2, 8)).astype(np.float32)
    flags = _cabi.SPS_PACK_IN_SPLIT | _cabi.SPS_PACK_IN2_SPLIT | _cabi.SPS_PACK_FOLD_LO
    ld = lib.sps_conv_kmajor_ld_f16x(K, cin, cin2, flags)
    assert ld == K * 32 + 64                      # 2 x 16 channels per offset (4 groups), the 1x1 term 2 x 16 padded to 64
    out = np.full((16, ld), np.nan, np.float16)
    assert lib.sps_conv_pack_kmajor_f16x(w.ctypes.data_as(C.c_void_p), K, cin, 8, w2.ctypes.data_as(C.c_void_p), cin2, flags,
                                         out.ctypes.data_as(C.c_void_p)) == 0
    assert np.isfinite(out).all()
    blk = out[:, : K * 32].reshape(16, K, 2, 2, 8)            # [row][offset][8-channel group][hi|lo slot of the ROW format][8]
    assert np.array_equal(blk[:, :, :, 0], blk[:, :, :, 1])   # the same weight meets the hi and the lo half of an activation
    hi, lo = blk[:8, :, :, 0].astype(np.float64), blk[8:, :, :, 0].astype(np.float64)
    ref = w.transpose(2, 0, 1).reshape(8, K, 2, 8).astype(np.float64)
    assert np.array_equal(blk[:8, :, :, 0], ref.astype(np.float16))
    assert np.abs(hi + lo - ref).max() <= 2.0 ** -21 * np.abs(ref).max()
    t2 = out[:, K * 32: K * 32 + 32].reshape(16, 2, 2, 8)
    assert np.abs(t2[:8, :, 0].astype(np.float64) + t2[8:, :, 0].astype(np.float64) - w2.T.reshape(8, 2, 8)).max() <= 2.0 ** -21 * np.abs(w2).max()
    assert not out[:, K * 32 + 32:].any()
    # plain packing is the flags == 0 case
    assert lib.sps_conv_kmajor_ld_f16x(K, cin, cin2, 0) == lib.sps_conv_kmajor_ld_f16(K, cin, cin2)
    assert lib.sps_conv_pack_kmajor_f16x(w.ctypes.data_as(C.c_void_p), K, cin, 16, None, 0, _cabi.SPS_PACK_FOLD_LO, out.ctypes.data_as(C.c_void_p)) == _cabi.SPS_ERR_BAD_ARG


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "sps_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("sps_oracle.layer_shapes", ""), f"{f} mentions the oracle"


def test_calculate_metrics_matches_reference_formulas():
    from sps_b200 import util
    from oracle import sps_oracle as O
    rng = np.random.default_rng(0)
    gt, pred = rng.integers(0, 2, 1000), rng.integers(0, 2, 1000)
    assert np.allclose(util.calculate_metrics(gt, pred), O.calculate_metrics(gt, pred))
    tp = np.sum((gt == 1) & (pred == 1)); fp = np.sum((gt == 0) & (pred == 1)); fn = np.sum((gt == 1) & (pred == 0))
    assert util.calculate_metrics(gt, pred)[4] == tp / (tp + fn + fp)
    assert util.calculate_metrics(np.zeros(4), np.zeros(4))[:3] == (0, 0, 0)


def test_two_segment_weight_packer_host_side():
    """sps_conv_pack_kmajor_f16s: with cin_split the K axis holds segment A of every offset (group-padded), then segment B of
    every offset (unpadded), then the 1x1 term; cin_split = 0 reproduces sps_conv_pack_kmajor_f16x."""
    from sps_b200 import _cabi
    lib = _cabi.load()
    rng = np.random.default_rng(2)
    for K, cin, cs, cout in [(81, 96, 64, 64), (81, 48, 32, 32), (81, 24, 16, 16), (8, 24, 16, 8)]:
        w = rng.standard_normal((K, cin, cout)).astype(np.float32)
        w2 = rng.standard_normal((cin, cout)).astype(np.float32)
        ga = cs // 8
        gpa = 1 if ga <= 1 else 2 if ga <= 2 else 4 if ga <= 4 else (ga + 7) // 8 * 8
        cb = cin - cs
        ld = lib.sps_conv_kmajor_ld_f16s(K, cin, cin, 0, cs)
        assert ld == K * (gpa * 8 + cb) + (cin + 63) // 64 * 64
        out = np.full((cout, ld), 7, np.float16)
        assert lib.sps_conv_pack_kmajor_f16s(w.ctypes.data_as(C.c_void_p), K, cin, cout, w2.ctypes.data_as(C.c_void_p), cin, 0, cs,
                                             out.ctypes.data_as(C.c_void_p)) == 0
        wt = w.astype(np.float16).transpose(2, 0, 1)                       # [cout][K][cin]
        seg_a = out[:, : K * gpa * 8].reshape(cout, K, gpa * 8)
        assert np.array_equal(seg_a[:, :, :cs], wt[:, :, :cs]) and not seg_a[:, :, cs:].any()
        seg_b = out[:, K * gpa * 8: K * (gpa * 8 + cb)].reshape(cout, K, cb)
        assert np.array_equal(seg_b, wt[:, :, cs:])
        tail = out[:, K * (gpa * 8 + cb):]
        assert np.array_equal(tail[:, :cin], w2.astype(np.float16).T) and not tail[:, cin:].any()
        # one segment: the _f16x layout
        ld0 = lib.sps_conv_kmajor_ld_f16s(K, cin, cin, 0, 0)
        assert ld0 == lib.sps_conv_kmajor_ld_f16x(K, cin, cin, 0)
        a = np.zeros((cout, ld0), np.float16)
        b = np.zeros((cout, ld0), np.float16)
        assert lib.sps_conv_pack_kmajor_f16s(w.ctypes.data_as(C.c_void_p), K, cin, cout, w2.ctypes.data_as(C.c_void_p), cin, 0, 0,
                                             a.ctypes.data_as(C.c_void_p)) == 0
        assert lib.sps_conv_pack_kmajor_f16x(w.ctypes.data_as(C.c_void_p), K, cin, cout, w2.ctypes.data_as(C.c_void_p), cin, 0,
                                             b.ctypes.data_as(C.c_void_p)) == 0
        assert np.array_equal(a, b)
    # bad arguments: split together with hi|lo rows, split not a multiple of 8, split >= cin
    w = np.zeros((8, 24, 8), np.float32)
    o = np.zeros((16, 4096), np.float16)
    for pf, cs in [(1, 16), (0, 12), (0, 24)]:
        assert lib.sps_conv_pack_kmajor_f16s(w.ctypes.data_as(C.c_void_p), 8, 24, 8, None, 0, pf, cs, o.ctypes.data_as(C.c_void_p)) != 0
