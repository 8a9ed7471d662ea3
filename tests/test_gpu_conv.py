"""Layer-level parity of sps_conv_fwd: the tcgen05 (TF32) kernel and the fp32 CUDA-core kernel
against a float64 numpy statement of the same contraction on random sparse kernel maps."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def tf32(x):
    a = np.ascontiguousarray(x, np.float32).view(np.uint32).astype(np.uint64)
    a = (a + 0xFFF + ((a >> 13) & 1)) & 0xFFFFE000
    return a.astype(np.uint32).view(np.float32)


def ref_conv(x, nbr, w, shift=None, x2=None, w2=None, res=None, relu=False, quant=None):
    q = quant or (lambda v: v)
    K, V = nbr.shape
    out = np.zeros((V, w.shape[-1]), np.float64)
    xq, wq = q(x).astype(np.float64), q(w).astype(np.float64)
    for k in range(K):
        o = np.nonzero(nbr[k] >= 0)[0]
        out[o] += xq[nbr[k, o]] @ wq[k]
    if x2 is not None:
        out += q(x2).astype(np.float64) @ q(w2).astype(np.float64)
    if shift is not None:
        out += shift
    if res is not None:
        out += res
    return np.maximum(out, 0) if relu else out


def random_map(rng, K, v_in, v_out, density):
    nbr = rng.integers(0, v_in, (K, v_out)).astype(np.int32)
    nbr[rng.random((K, v_out)) > density] = -1
    return nbr


def run(backend, x, nbr, w, ld=None, **kw):
    from sps_b200 import convops
    dev = lambda a: None if a is None else torch.as_tensor(np.ascontiguousarray(a)).cuda()
    K, V = nbr.shape
    ld = ld or ((V + 31) // 32 * 32)
    m = np.full((K, ld), -1, np.int32)
    m[:, :V] = nbr
    wd, w2d = dev(w), dev(kw.get("w2"))
    wt = convops.pack_kmajor(wd, w2d) if backend != 1 else None
    n_out = torch.tensor([V], dtype=torch.int32, device="cuda")
    out = convops.conv_fwd(dev(x), wd, n_out, map=dev(m), map_ld=ld, shift=dev(kw.get("shift")), in2=dev(kw.get("x2")),
                           weight2=w2d, res=dev(kw.get("res")), relu=kw.get("relu", False), weight_kmajor=wt,
                           backend=backend)
    torch.cuda.synchronize()
    return out[:V].cpu().numpy()


def test_umma_dense_gemm_identity_map():
    """K=1 identity map = plain [M,K]@[K,N]: isolates descriptors / swizzle / TMEM read-back."""
    rng = np.random.default_rng(0)
    for cin, cout, V in [(32, 64, 128), (64, 64, 1000), (8, 16, 300), (96, 32, 257), (16, 8, 129)]:
        x = rng.standard_normal((V, cin)).astype(np.float32)
        w = rng.standard_normal((1, cin, cout)).astype(np.float32)
        nbr = np.arange(V, dtype=np.int32)[None, :]
        got = run(2, x, nbr, w)
        ref = ref_conv(x, nbr, w, quant=tf32)
        assert np.abs(got - ref).max() < 6e-3 * np.sqrt(cin), (cin, cout, V, np.abs(got - ref).max())
        # input rounding is truncation in hardware for non-pre-rounded activations: pre-round and be tight
        got = run(2, tf32(x), nbr, w)
        assert np.abs(got - ref).max() < 1e-4 * np.sqrt(cin), (cin, cout, V, np.abs(got - ref).max())


@pytest.mark.parametrize("cin,cout", [(8, 8), (8, 16), (16, 16), (24, 16), (16, 32), (32, 32), (48, 32), (32, 64),
                                      (64, 64), (96, 64), (16, 8)])
@pytest.mark.parametrize("backend", [1, 2])
def test_sparse_conv_81_offsets(cin, cout, backend):
    rng = np.random.default_rng(cin * 100 + cout)
    V = 1000 + cin
    x = tf32(rng.standard_normal((V, cin)).astype(np.float32))
    w = (rng.standard_normal((81, cin, cout)) / np.sqrt(20 * cin)).astype(np.float32)
    nbr = random_map(rng, 81, V, V, 0.25)
    nbr[:, 100:228] = -1            # a whole tile without neighbours
    nbr[5:40, :] = -1               # offsets absent everywhere (skipped by the tile prologue)
    shift = rng.standard_normal(cout).astype(np.float32)
    x2 = tf32(rng.standard_normal((V, 24)).astype(np.float32))
    w2 = (rng.standard_normal((24, cout)) / 5).astype(np.float32)
    res = rng.standard_normal((V, cout)).astype(np.float32)
    quant = tf32 if backend == 2 else None
    tol = 2e-4 if backend == 2 else 2e-5
    got = run(backend, x, nbr, w, shift=shift, relu=True)
    assert np.abs(got - ref_conv(x, nbr, w, shift=shift, relu=True, quant=quant)).max() < tol
    got = run(backend, x, nbr, w, shift=shift, x2=x2, w2=w2, relu=True)
    assert np.abs(got - ref_conv(x, nbr, w, shift=shift, x2=x2, w2=w2, relu=True, quant=quant)).max() < tol
    got = run(backend, x, nbr, w, res=res)
    assert np.abs(got - ref_conv(x, nbr, w, res=res, quant=quant)).max() < tol


def f16(x):
    return np.ascontiguousarray(x, np.float32).astype(np.float16).astype(np.float32)


@pytest.mark.parametrize("cin,cout", [(8, 8), (8, 16), (16, 16), (24, 16), (16, 32), (32, 32), (48, 32), (64, 64), (96, 64),
                                      (16, 8)])
def test_sparse_conv_fp16_rows(cin, cout):
    """SPS_IO_F16 through sps_conv_fwd: fp16 rows and weights (kind::f16), fp32 accumulate and epilogue, fp16 out;
    81 offsets, the fused 1x1 term, residual, ReLU -- against float64 numpy on the fp16-rounded operands."""
    from sps_b200 import convops
    rng = np.random.default_rng(cin * 100 + cout)
    V, K = 1500, 81
    nbr = random_map(rng, K, V, V, 0.3)
    nbr[:, 700:900] = -1                      # rows without any neighbour
    nbr[40:, 1000:1200] = -1                  # tiles that use only some offsets
    x = rng.standard_normal((V, cin)).astype(np.float32)
    w = (rng.standard_normal((K, cin, cout)) / np.sqrt(cin * 8)).astype(np.float32)
    shift = rng.standard_normal(cout).astype(np.float32)
    cin2 = 8 if cin == 8 else cin // 2 if (cin // 2) % 8 == 0 else cin
    x2 = rng.standard_normal((V, cin2)).astype(np.float32)
    w2 = (rng.standard_normal((cin2, cout)) / np.sqrt(cin2)).astype(np.float32)
    res = rng.standard_normal((V, cout)).astype(np.float32)
    ld = (V + 31) // 32 * 32
    m = np.full((K, ld), -1, np.int32)
    m[:, :V] = nbr
    dev = lambda a, dt=None: torch.as_tensor(np.ascontiguousarray(a)).cuda().to(dt) if dt else torch.as_tensor(np.ascontiguousarray(a)).cuda()
    n_out = torch.tensor([V], dtype=torch.int32, device="cuda")
    half = torch.float16
    for kw in (dict(shift=True, relu=True), dict(shift=True, relu=True, fused=True), dict(res=True)):
        wt = convops.pack_kmajor_f16(dev(w), dev(w2) if kw.get("fused") else None)
        out = convops.conv_fwd(dev(x, half), dev(w), n_out, map=dev(m), map_ld=ld,
                               shift=dev(shift) if kw.get("shift") else None,
                               in2=dev(x2, half) if kw.get("fused") else None, weight2=dev(w2) if kw.get("fused") else None,
                               res=dev(res, half) if kw.get("res") else None, relu=kw.get("relu", False),
                               weight_kmajor=wt, io_f16=True, backend=3)   # 3: the tensor-core kernel also for 8 output channels
        torch.cuda.synchronize()
        assert out.dtype == half
        got = out[:V].float().cpu().numpy()
        ref = ref_conv(x, nbr, w, shift=shift if kw.get("shift") else None, x2=x2 if kw.get("fused") else None,
                       w2=w2 if kw.get("fused") else None, res=f16(res) if kw.get("res") else None,
                       relu=kw.get("relu", False), quant=f16)
        # fp32 accumulation of exact fp16 products, then ONE rounding to fp16 on the way out (2^-11 relative)
        assert np.abs(got - ref).max() < 2e-3 * max(1.0, np.abs(ref).max()), (kw, np.abs(got - ref).max())


def test_unet_backends_agree():
    """Whole forward: tcgen05 path vs fp32 CUDA-core path vs CPU oracle (2e-3 bar, north star)."""
    from conftest import make_case
    from oracle import sps_oracle as O, me_cpu
    from sps_b200 import engine, _cabi
    lib = _cabi.load()
    rows = make_case("hdl-32", seed=4, n_map_poses=6)
    pts = rows[:, :5]
    sd = O.make_state_dict(seed=0, randomize_bn=True)
    ref, _, _ = me_cpu.forward(pts, 0.1, me_cpu.pack_weights(sd))
    net = engine.Net(sd)
    eng = engine.Engine(len(pts))
    d = torch.as_tensor(pts).cuda()
    res = {}
    for backend in (1, 2, 0):   # exact fp32 / TF32 operands on fp32 rows / fp16 rows (default)
        eng.set_conv_backend(backend)
        res[backend] = eng.forward(net, d, 0.1).cpu().numpy()
        eng.status()
    assert np.abs(res[1] - ref).max() < 1e-5
    for backend in (2, 0):
        assert np.abs(res[backend] - ref).max() < 5e-4, (backend, np.abs(res[backend] - ref).max())
        assert np.mean((res[backend] < 0.84) == (ref < 0.84)) >= 0.999


def test_pattern_sorted_processing_order_does_not_change_results():
    """sps_set_pattern_sort: rows visited in neighbourhood-shape order (radix sort + permuted tile masks)."""
    from conftest import make_case
    from oracle import sps_oracle as O, me_cpu
    from sps_b200 import engine, _cabi
    lib = _cabi.load()
    rows = make_case("hdl-32", seed=9, n_map_poses=6)
    pts = rows[:, :5]
    sd = O.make_state_dict(seed=0, randomize_bn=True)
    net = engine.Net(sd)
    eng = engine.Engine(len(pts))
    d = torch.as_tensor(pts).cuda()
    ref, _, _ = me_cpu.forward(pts, 0.1, me_cpu.pack_weights(sd))
    # fp32 rows: only the grouping of zero contributions differs.  fp16 rows: a last-bit fp32 difference can
    # flip the rounding of a stored activation (2^-11 relative), so the two orders agree to ~3e-4.
    for backend, tol in ((2, 2e-6), (0, 1e-3)):
        out = {}
        eng.set_conv_backend(backend)
        for mode in (0, 2):
            eng.set_pattern_sort(mode)
            out[mode] = eng.forward(net, d, 0.1).cpu().numpy()
            eng.status()
        assert np.abs(out[0] - out[2]).max() < tol, (backend, np.abs(out[0] - out[2]).max())
        assert np.abs(out[2] - ref).max() < 5e-4


@pytest.mark.parametrize("cin,in_split", [(8, False), (8, True), (16, False), (16, True)])
@pytest.mark.parametrize("K", [81, 8])
def test_split_precision_options_of_the_fp16_path(cin, in_split, K):
    """SPS_CONV_FOLD_LO (low weight parts in accumulator columns 8..15) and hi|lo rows (SPS_CONV_OUT_SPLIT on the way
    out, doubled channels + K-duplicated weights on the way in): with all of them an 8-output-channel fp16-row layer
    reproduces float64 numpy on UNROUNDED operands to ~1e-6 relative, i.e. ~1000x tighter than plain fp16 operands."""
    from sps_b200 import convops, _cabi
    rng = np.random.default_rng(cin * 10 + K + in_split)
    V = 1300
    nbr = random_map(rng, K, V, V, 0.3 if K == 81 else 0.6)
    nbr[:, 300:500] = -1
    x = rng.standard_normal((V, cin)).astype(np.float32)
    w = (rng.standard_normal((K, cin, 8)) / np.sqrt(cin * 8)).astype(np.float32)
    shift = rng.standard_normal(8).astype(np.float32)
    x2 = rng.standard_normal((V, 16)).astype(np.float32)
    w2 = (rng.standard_normal((16, 8)) / 4).astype(np.float32)
    head_w = rng.standard_normal(8).astype(np.float32)
    ld = (V + 31) // 32 * 32
    m = np.full((K, ld), -1, np.int32)
    m[:, :V] = nbr
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a)).cuda()
    n_out = torch.tensor([V], dtype=torch.int32, device="cuda")
    xin = convops.split_rows(t(x)) if in_split else t(x).half()
    xq = x if in_split else f16(x)                       # plain fp16 rows round the input once
    # (a) folded low weights, hi|lo output
    wt = convops.pack_kmajor_f16x(t(w), in_split=in_split, fold_lo=True)
    out = convops.conv_fwd(xin, t(w), n_out, map=t(m), map_ld=ld, shift=t(shift), relu=True, weight_kmajor=wt, io_f16=True,
                           flags=_cabi.SPS_CONV_FOLD_LO | _cabi.SPS_CONV_OUT_SPLIT, cin_rows=2 * cin if in_split else None)
    ref = ref_conv(xq, nbr, w, shift=shift, relu=True)
    assert out.shape[1] == 16 and out.dtype == torch.float16
    got = convops.merge_rows(out[:V]).cpu().numpy()
    assert np.abs(got - ref).max() < 4e-6 * max(1.0, np.abs(ref).max()), np.abs(got - ref).max()
    # (b) the same with plain fp16 output: one rounding on the way out
    out = convops.conv_fwd(xin, t(w), n_out, map=t(m), map_ld=ld, shift=t(shift), relu=True, weight_kmajor=wt, io_f16=True,
                           flags=_cabi.SPS_CONV_FOLD_LO, cin_rows=2 * cin if in_split else None)
    assert out.shape[1] == 8
    assert np.abs(out[:V].float().cpu().numpy() - ref).max() < 1e-3 * max(1.0, np.abs(ref).max())
    # (c) fused 1x1 term on hi|lo rows + fused head (block8.conv2 + final)
    wt = convops.pack_kmajor_f16x(t(w), t(w2), in_split=in_split, in2_split=True, fold_lo=True)
    head_out = torch.zeros(V, dtype=torch.float32, device="cuda")
    convops.conv_fwd(xin, t(w), n_out, map=t(m), map_ld=ld, shift=t(shift), relu=True, in2=convops.split_rows(t(x2)),
                     weight2=t(w2), head_w=t(head_w), head_b=0.25, head_out=head_out, weight_kmajor=wt, io_f16=True,
                     flags=_cabi.SPS_CONV_FOLD_LO, cin_rows=2 * cin if in_split else None, cin2_rows=32)
    ref = ref_conv(xq, nbr, w, shift=shift, x2=x2, w2=w2, relu=True) @ head_w.astype(np.float64) + 0.25
    assert np.abs(head_out.cpu().numpy() - ref).max() < 1e-5 * max(1.0, np.abs(ref).max())


def test_split_rows_into_concat_buffer_slices():
    """hi|lo output into a channel slice of a wider hi|lo buffer (convtr7p2s2 -> the level-0 concat buffer)."""
    from sps_b200 import convops, _cabi
    rng = np.random.default_rng(5)
    V, K = 700, 8
    nbr = random_map(rng, K, V, V, 0.5)
    ld = (V + 31) // 32 * 32
    m = np.full((K, ld), -1, np.int32)
    m[:, :V] = nbr
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a)).cuda()
    x = rng.standard_normal((V, 16)).astype(np.float32)
    w = (rng.standard_normal((K, 16, 8)) / 8).astype(np.float32)
    cat8 = torch.zeros((V, 32), dtype=torch.float16, device="cuda")
    n_out = torch.tensor([V], dtype=torch.int32, device="cuda")
    wt = convops.pack_kmajor_f16x(t(w), fold_lo=True)
    convops.conv_fwd(t(x).half(), t(w), n_out, map=t(m), map_ld=ld, relu=True, out=cat8[:, :16], weight_kmajor=wt, io_f16=True,
                     flags=_cabi.SPS_CONV_FOLD_LO | _cabi.SPS_CONV_OUT_SPLIT)
    ref = ref_conv(f16(x), nbr, w, relu=True)
    got = convops.merge_rows(cat8[:, :16].contiguous()).cpu().numpy()
    assert np.abs(got - ref).max() < 4e-6 * max(1.0, np.abs(ref).max())
    assert (cat8[:, 16:] == 0).all()           # the other half of the concat buffer is untouched


@pytest.mark.parametrize("cin,cout", [(64, 128), (128, 128), (192, 128), (96, 64), (128, 256), (384, 256), (256, 512), (768, 512)])
def test_wide_layers_of_the_width_sweep(cin, cout):
    """BASELINE configs[4]: PLANES x2 ... x8 (Cin / Cout up to 768 / 512).  N = 128 / 256 accumulators (512 output channels
    as two passes), weight slabs through TMA, chunked epilogue -- against float64 numpy on the fp16-rounded operands,
    with the fused 1x1 term and the identity residual."""
    from sps_b200 import convops
    rng = np.random.default_rng(cin + cout)
    V, K = 700, 81
    nbr = random_map(rng, K, V, V, 0.3)
    nbr[:, 300:450] = -1
    nbr[50:, 500:] = -1
    x = rng.standard_normal((V, cin)).astype(np.float32)
    w = (rng.standard_normal((K, cin, cout)) / np.sqrt(cin * 8)).astype(np.float32)
    shift = rng.standard_normal(cout).astype(np.float32)
    cin2 = cin // 2 if (cin // 2) % 64 == 0 else 64
    x2 = rng.standard_normal((V, cin2)).astype(np.float32)
    w2 = (rng.standard_normal((cin2, cout)) / np.sqrt(cin2)).astype(np.float32)
    res = rng.standard_normal((V, cout)).astype(np.float32)
    ld = (V + 31) // 32 * 32
    m = np.full((K, ld), -1, np.int32)
    m[:, :V] = nbr
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a)).cuda()
    n_out = torch.tensor([V], dtype=torch.int32, device="cuda")
    for kw in (dict(shift=True, relu=True, fused=True), dict(res=True)):
        wt = convops.pack_kmajor_f16(t(w), t(w2) if kw.get("fused") else None)
        out = convops.conv_fwd(t(x).half(), t(w), n_out, map=t(m), map_ld=ld, shift=t(shift) if kw.get("shift") else None,
                               in2=t(x2).half() if kw.get("fused") else None, weight2=t(w2) if kw.get("fused") else None,
                               res=t(res).half() if kw.get("res") else None, relu=kw.get("relu", False), weight_kmajor=wt,
                               io_f16=True, backend=3)
        torch.cuda.synchronize()
        got = out[:V].float().cpu().numpy()
        ref = ref_conv(x, nbr, w, shift=shift if kw.get("shift") else None, x2=x2 if kw.get("fused") else None,
                       w2=w2 if kw.get("fused") else None, res=f16(res) if kw.get("res") else None, relu=kw.get("relu", False),
                       quant=f16)
        assert got.shape == (V, cout)
        assert np.abs(got - ref).max() < 2e-3 * max(1.0, np.abs(ref).max()), (kw, np.abs(got - ref).max())


@pytest.mark.parametrize("cin,cin_split,cout", [(96, 64, 64), (48, 32, 32), (24, 16, 16), (96, 64, 32), (24, 16, 8)])
def test_two_segment_input_rows(cin, cin_split, cout):
    """sps_conv_args.cin_split: the rows of a concat buffer walked segment by segment along K (no padded columns).  Same
    products, same fp32 accumulator: equal to the float64 reference at the fp16 bound, and to the one-segment kernel
    within fp32 summation-order noise.  With the fused 1x1 term on the same rows (the decoder blocks' downsample)."""
    from sps_b200 import convops
    rng = np.random.default_rng(cin * 7 + cout)
    V, K = 2100, 81
    nbr = random_map(rng, K, V, V, 0.3)
    nbr[:, 700:900] = -1
    nbr[33:, 1000:1400] = -1                  # tiles with an odd number of present offsets (ragged last stage of both segments)
    x = rng.standard_normal((V, cin)).astype(np.float32)
    w = (rng.standard_normal((K, cin, cout)) / np.sqrt(cin * 8)).astype(np.float32)
    w2 = (rng.standard_normal((cin, cout)) / np.sqrt(cin)).astype(np.float32)
    shift = rng.standard_normal(cout).astype(np.float32)
    ld = (V + 31) // 32 * 32
    m = np.full((K, ld), -1, np.int32)
    m[:, :V] = nbr
    dev = lambda a, dt=None: torch.as_tensor(np.ascontiguousarray(a)).cuda().to(dt) if dt else torch.as_tensor(np.ascontiguousarray(a)).cuda()
    n_out = torch.tensor([V], dtype=torch.int32, device="cuda")
    half = torch.float16
    for fused in (False, True):
        outs = []
        for cs in (0, cin_split):
            wt = convops.pack_kmajor_f16x(dev(w), dev(w2) if fused else None, cin_split=cs, fold_lo=cout == 8)
            out = convops.conv_fwd(dev(x, half), dev(w), n_out, map=dev(m), map_ld=ld, shift=dev(shift), relu=True,
                                   in2=dev(x, half) if fused else None, weight2=dev(w2) if fused else None,
                                   weight_kmajor=wt, io_f16=True, backend=3, cin_split=cs, flags=1 if cout == 8 else 0)
            torch.cuda.synchronize()
            outs.append(out[:V].float().cpu().numpy())
        ref = ref_conv(x, nbr, w, shift=shift, x2=x if fused else None, w2=w2 if fused else None, relu=True, quant=f16)
        scale = max(1.0, np.abs(ref).max())
        if cout != 8:   # (folded low weight parts are more exact than the fp16-weight reference)
            assert np.abs(outs[1] - ref).max() < 2e-3 * scale, (fused, np.abs(outs[1] - ref).max())
        assert np.abs(outs[1] - outs[0]).max() < 2e-3 * scale, (fused, np.abs(outs[1] - outs[0]).max())


@pytest.mark.parametrize("cin,cout,io_f16", [(64, 64, True), (32, 16, True), (16, 8, True), (64, 32, False)])
def test_transposed_conv_reads_the_parent_array(cin, cout, io_f16):
    """SPS_CONV_MAP_PARENT: `map` is the fine level's parent array (coarse row * 8 + child class) instead of a dense
    [8][ld] up-map; the result equals the dense-map call bit for bit (same products, same order) and the float64
    reference at the operand precision.  Ragged last tile, rows of every class."""
    from sps_b200 import convops, _cabi
    rng = np.random.default_rng(cin + cout)
    v_out, v_in, K = 1000, 300, 8
    parent_row = rng.integers(0, v_in, v_out)
    cls = rng.integers(0, 8, v_out)
    parent = (parent_row * 8 + cls).astype(np.int32)
    ld = (v_out + 31) // 32 * 32
    up = np.full((K, ld), -1, np.int32)
    up[cls, np.arange(v_out)] = parent_row
    x = rng.standard_normal((v_in, cin)).astype(np.float32)
    w = (rng.standard_normal((K, cin, cout)) / np.sqrt(cin)).astype(np.float32)
    shift = rng.standard_normal(cout).astype(np.float32)
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a)).cuda()
    n_out = torch.tensor([v_out], dtype=torch.int32, device="cuda")
    masks = torch.zeros(((v_out + 127) // 128, 4), dtype=torch.int32, device="cuda")
    masks[:, 0] = 0xFF
    if io_f16:
        wt = convops.pack_kmajor_f16x(t(w), fold_lo=cout == 8)
        xin = t(x).half()
        kw = dict(io_f16=True, backend=3, flags=1 if cout == 8 else 0)
        quant = f16
    else:
        wt = convops.pack_kmajor(t(w))
        xin = t(x)
        kw = dict(backend=2)
        quant = tf32
    dense = convops.conv_fwd(xin, t(w), n_out, map=t(up), map_ld=ld, shift=t(shift), relu=True, weight_kmajor=wt,
                             tile_mask=masks, **kw)
    kw["flags"] = kw.get("flags", 0) | _cabi.SPS_CONV_MAP_PARENT
    par = convops.conv_fwd(xin, t(w), n_out, map=t(parent), map_ld=ld, shift=t(shift), relu=True, weight_kmajor=wt,
                           tile_mask=masks, **kw)
    torch.cuda.synchronize()
    assert torch.equal(dense[:v_out], par[:v_out])
    ref = ref_conv(x, up[:, :v_out], w, shift=shift, relu=True, quant=quant)
    got = par[:v_out].float().cpu().numpy()
    if cout != 8:
        assert np.abs(got - ref).max() < 2e-3 * max(1.0, np.abs(ref).max())
