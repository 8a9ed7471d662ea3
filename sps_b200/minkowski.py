"""MinkowskiEngine-shaped layer API on the B200 engine -- the part of ME's Python surface that the
reference touches: ``TensorField``/``.sparse()``, ``SparseTensor``/``.slice()``,
``MinkowskiConvolution`` / ``MinkowskiConvolutionTranspose`` / ``MinkowskiBatchNorm`` /
``MinkowskiReLU``, ``cat``, ``utils.kaiming_normal_``, ``modules.resnet_block.BasicBlock``.

``install()`` registers this module as ``MinkowskiEngine`` so that the reference's own
``src/sps/models/MinkowskiEngine/{resnet,minkunet,customminkunet}.py`` import and run unchanged.
It is the compatibility path: one kernel launch per layer, BatchNorm/ReLU as separate passes; the
product path for SPS is the fused ``sps_forward`` behind ``sps_b200.models.SPSModel``.

Kernel shapes served (anything else raises): 1 (any stride-1 layer), 3 in all four dimensions,
[5,5,5,1] at tensor stride 1, [2,2,2,1] with stride [2,2,2,1] (conv and transposed conv) -- i.e.
every layer of ``MinkUNetBase`` (minkunet.py:52-159).  ``MinkowskiUnion`` / ``MinkowskiPruning``
(only used by ``util.prune``) are replaced by ``sps_b200.util.prune`` on the map hash.
"""
from __future__ import annotations

import ctypes as C
import math
import sys
import types

import torch
import torch.nn as nn

from . import _cabi, convops
from ._cabi import check
from .engine import Engine, _ptr, _stream

__version__ = "0.5.4+sps_b200"


def _as_list(v, D):
    return [int(v)] * D if isinstance(v, int) else [int(x) for x in v]


class CoordinateManager:
    """Owns one engine context: the coordinate sets of tensor strides 1..16 and their kernel maps."""

    def __init__(self, n_points: int, device):
        self.engine = Engine(max(int(n_points), 1), device)
        self.n_points = int(n_points)
        self.counts = None

    def build(self, coordinates: torch.Tensor):
        self.engine.voxelize(coordinates, 1.0)         # coordinates arrive already divided by the voxel size
        self.engine.build_maps()
        self.engine.status()
        self.counts = [self.engine.count(L) for L in range(5)]


class _SetManager:
    """Coordinate manager of SparseTensors built directly from integer coordinates (util.prune, util.py:86-95): the
    tensors that share it live on one integer lattice; there are no strided levels or kernel maps behind it."""


def _unique_rows(coords: torch.Tensor, feats: torch.Tensor, reduce: str):
    """Unique integer rows in first-occurrence order through the voxel hash of the library (columns padded to the
    5-column key (b, x, y, z, t)); features of coincident rows are summed (``sum``) or the first one is kept (``first``,
    ME's default RANDOM_SUBSAMPLE picks an arbitrary member: the first is a deterministic choice)."""
    n, d = coords.shape
    if d > 3:
        raise NotImplementedError("directly constructed SparseTensors take up to 3 coordinate columns (util.py:86-95)")
    dev = coords.device
    rows = torch.zeros((n, 5), dtype=torch.float32, device=dev)
    rows[:, 1:1 + d] = coords.to(torch.float32)          # exact: |c| < 2^24
    eng = Engine(max(n, 1), dev)
    eng.voxelize(rows, 1.0)
    eng.status()
    v = eng.count(0)
    uc = torch.as_tensor(eng.coords(0)[:, 1:1 + d], device=dev).to(coords.dtype)
    c = feats.shape[1]
    f = feats.contiguous().to(torch.float32)
    if reduce == "sum":
        out = torch.empty((max(n, 1), c), dtype=torch.float32, device=dev)
        cnt = torch.empty(max(n, 1), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(eng.lib.sps_voxel_sum(eng.handle, _ptr(f), f.stride(0), c, _ptr(out), _ptr(cnt), _stream()), "sps_voxel_sum")
        uf = out[:v].clone()
    else:
        inv = torch.as_tensor(eng.inverse_map(), device=dev).long()
        first = torch.full((v,), n, dtype=torch.long, device=dev)
        first.scatter_reduce_(0, inv, torch.arange(n, device=dev), reduce="amin")
        uf = f[first]
    return uc, uf


class SparseTensor:
    def __init__(self, features, coordinates=None, coordinate_manager=None, tensor_stride=1, level=0):
        self.level = level
        self.tensor_stride = [2 ** level] * 3 + [1]
        self._coords = None
        if coordinates is not None:
            # ME.SparseTensor(features=, coordinates=[, coordinate_manager=]) on integer coordinates (util.py:86-95):
            # duplicates collapse to one row
            if not coordinates.is_cuda:
                raise RuntimeError("sps_b200 has no CPU path: SparseTensor needs CUDA tensors")
            self._coords, self.F = _unique_rows(coordinates, features, "first")
            self.coordinate_manager = coordinate_manager if coordinate_manager is not None else _SetManager()
            return
        if coordinate_manager is None:
            raise NotImplementedError("a SparseTensor needs coordinates or the coordinate manager of TensorField(...).sparse()")
        self.F = features
        self.coordinate_manager = coordinate_manager

    @classmethod
    def _from_set(cls, coords, feats, manager):
        t = cls.__new__(cls)
        t.level, t.tensor_stride, t._coords, t.F, t.coordinate_manager = 0, [1, 1, 1, 1], coords, feats, manager
        return t

    @property
    def features(self):
        return self.F

    @property
    def C(self):
        if self._coords is not None:
            return self._coords
        return torch.as_tensor(self.coordinate_manager.engine.coords(self.level), device=self.F.device)

    coordinates = C

    @property
    def D(self):
        return 4

    def __len__(self):
        return self.F.shape[0]

    def slice(self, field: "TensorField") -> "TensorField":
        """models.py:28: features of the voxel each original point fell into."""
        eng = self.coordinate_manager.engine
        n = field.coordinates.shape[0]
        out = torch.empty((n, self.F.shape[1]), dtype=torch.float32, device=self.F.device)
        check(eng.lib.sps_gather_rows(_ptr(self.F), self.F.stride(0), self.F.shape[1],
                                      C.c_void_p(eng.lib.sps_ctx_inverse_map(eng.handle)), n, _ptr(out), _stream()),
              "sps_gather_rows")
        return TensorField(features=out, coordinates=field.coordinates, _manager=self.coordinate_manager)


class TensorField:
    def __init__(self, features, coordinates, _manager=None, **kwargs):
        if not features.is_cuda:
            raise RuntimeError("sps_b200 has no CPU path: TensorField needs CUDA tensors")
        self.F = features.contiguous().to(torch.float32)
        self.coordinates = coordinates.contiguous().to(torch.float32)
        self.coordinate_manager = _manager

    @property
    def features(self):
        return self.F

    def sparse(self) -> SparseTensor:
        """Floor the coordinates, de-duplicate, average the features per voxel (models.py:24-25)."""
        n = self.coordinates.shape[0]
        mgr = CoordinateManager(n, self.coordinates.device)
        mgr.build(self.coordinates)
        self.coordinate_manager = mgr
        eng = mgr.engine
        c = self.F.shape[1]
        out = torch.empty((max(n, 1), c), dtype=torch.float32, device=self.F.device)
        cnt = torch.empty(max(n, 1), dtype=torch.float32, device=self.F.device)
        check(eng.lib.sps_voxel_mean(eng.handle, _ptr(self.F), self.F.stride(0), c, _ptr(out), _ptr(cnt), _stream()),
              "sps_voxel_mean")
        return SparseTensor(out[: mgr.counts[0]], coordinate_manager=mgr, level=0)


class _ConvBase(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=-1, stride=1, dilation=1, bias=False, dimension=None,
                 transpose=False, **kwargs):
        super().__init__()
        assert dimension is not None and dimension > 0
        if dimension != 4:
            raise NotImplementedError("sps_b200 serves the 4-D layers of SPS (D=4, models.py:17)")
        self.in_channels, self.out_channels, self.dimension = in_channels, out_channels, dimension
        self.kernel_size = _as_list(kernel_size, dimension)
        self.stride = _as_list(stride, dimension)
        if _as_list(dilation, dimension) != [1] * dimension:
            raise NotImplementedError("dilation != 1")
        self.is_transpose = transpose
        self.kernel_volume = math.prod(self.kernel_size)
        self.use_mm = self.kernel_volume == 1 and self.stride == [1] * dimension and not transpose
        shape = (in_channels, out_channels) if self.use_mm else (self.kernel_volume, in_channels, out_channels)
        self.kernel = nn.Parameter(torch.empty(shape))
        self.bias = nn.Parameter(torch.empty(1, out_channels)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):
        with torch.no_grad():
            n = (self.out_channels if self.is_transpose else self.in_channels) * self.kernel_volume
            stdv = 1.0 / math.sqrt(n)
            self.kernel.uniform_(-stdv, stdv)
            if self.bias is not None:
                self.bias.uniform_(-stdv, stdv)

    def _run(self, x: SparseTensor, mode, map_ptr, map_ld, out_level, K):
        mgr = x.coordinate_manager
        eng = mgr.engine
        n_out = mgr.counts[out_level]
        lib = eng.lib
        w = self.kernel if self.kernel.dim() == 3 else self.kernel.unsqueeze(0)
        w = w.detach().contiguous()
        out = torch.empty((max(n_out, 1), self.out_channels), dtype=torch.float32, device=x.F.device)
        a = _cabi.ConvArgs()
        a.mode, a.K, a.cin, a.cout = mode, K, self.in_channels, self.out_channels
        a.map, a.map_ld = map_ptr, map_ld
        lv_count = eng.level(out_level if mode == _cabi.SPS_CONV_NBR else out_level + 1).count
        a.n_out = lv_count
        a.n_out_max = mgr.counts[out_level if mode == _cabi.SPS_CONV_NBR else out_level + 1]
        feat = x.F if x.F.stride(1) == 1 else x.F.contiguous()
        if self.in_channels % 4 and self.in_channels != 1:
            raise NotImplementedError("channel counts must be 1 or a multiple of 4")
        a.in_, a.in_ld = feat.data_ptr(), feat.stride(0)
        a.weight = w.data_ptr()
        if self.bias is not None:
            b = self.bias.detach().reshape(-1).contiguous()
            a.shift = b.data_ptr()
        a.out, a.out_ld = out.data_ptr(), out.stride(0)
        a.backend = _cabi.SPS_BACKEND_FP32                       # layer API: exact fp32 kernels, chosen per call
        with torch.cuda.device(x.F.device):
            check(lib.sps_conv_fwd(C.byref(a), _stream()), "sps_conv_fwd")
        return SparseTensor(out[:n_out], coordinate_manager=mgr, level=out_level)


class MinkowskiConvolution(_ConvBase):
    """ME.MinkowskiConvolution (minkunet.py:55-60,64-70,152-158; resnet.py:100-106; BasicBlock)."""

    def __init__(self, in_channels, out_channels, kernel_size=-1, stride=1, dilation=1, bias=False, dimension=None,
                 **kwargs):
        super().__init__(in_channels, out_channels, kernel_size, stride, dilation, bias, dimension, transpose=False)

    def forward(self, x: SparseTensor) -> SparseTensor:
        v = x.coordinate_manager.engine.level(x.level)
        ks, st = self.kernel_size, self.stride
        if self.use_mm:
            return self._run(x, _cabi.SPS_CONV_NBR, None, 0, x.level, 1)
        if ks == [3, 3, 3, 3] and st == [1, 1, 1, 1]:
            return self._run(x, _cabi.SPS_CONV_NBR, v.nbr3, v.ld, x.level, 81)
        if ks == [5, 5, 5, 1] and st == [1, 1, 1, 1] and x.level == 0:
            return self._run(x, _cabi.SPS_CONV_NBR, v.nbr5, v.ld, 0, 125)
        if ks == [2, 2, 2, 1] and st == [2, 2, 2, 1] and x.level < 4:
            up = x.coordinate_manager.engine.level(x.level + 1)
            return self._run(x, _cabi.SPS_CONV_NBR, up.child, up.ld, x.level + 1, 8)
        raise NotImplementedError(f"kernel_size={ks} stride={st} at tensor stride {x.tensor_stride}")


class MinkowskiConvolutionTranspose(_ConvBase):
    """ME.MinkowskiConvolutionTranspose k=[2,2,2,1], s=[2,2,2,1] onto the existing finer map
    (minkunet.py:107-113)."""

    def __init__(self, in_channels, out_channels, kernel_size=-1, stride=1, dilation=1, bias=False, dimension=None,
                 **kwargs):
        super().__init__(in_channels, out_channels, kernel_size, stride, dilation, bias, dimension, transpose=True)

    def forward(self, x: SparseTensor) -> SparseTensor:
        if self.kernel_size != [2, 2, 2, 1] or self.stride != [2, 2, 2, 1] or x.level < 1:
            raise NotImplementedError(f"transposed kernel_size={self.kernel_size} stride={self.stride}")
        v = x.coordinate_manager.engine.level(x.level)
        return self._run(x, _cabi.SPS_CONV_UP, v.child, v.ld, x.level - 1, 8)


def _affine(x: SparseTensor, scale, shift, relu):
    eng = x.coordinate_manager.engine
    out = torch.empty_like(x.F)
    check(eng.lib.sps_affine_relu(_ptr(x.F), x.F.stride(0), x.F.shape[1], x.F.shape[0], _ptr(scale), _ptr(shift),
                                  int(relu), _ptr(out), out.stride(0), _stream()), "sps_affine_relu")
    return SparseTensor(out, coordinate_manager=x.coordinate_manager, level=x.level)


class MinkowskiBatchNorm(nn.Module):
    """ME.MinkowskiBatchNorm: ``nn.BatchNorm1d`` on the feature matrix (attribute ``bn``); eval only."""

    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True):
        super().__init__()
        self.bn = nn.BatchNorm1d(num_features, eps=eps, momentum=momentum, affine=affine,
                                 track_running_stats=track_running_stats)

    def forward(self, x: SparseTensor) -> SparseTensor:
        if self.training:
            raise RuntimeError("sps_b200 is inference-only: call .eval() (training is out of scope)")
        bn = self.bn
        inv = 1.0 / torch.sqrt(bn.running_var + bn.eps)
        scale = (bn.weight * inv).detach().contiguous()
        shift = (bn.bias - bn.running_mean * bn.weight * inv).detach().contiguous()
        return _affine(x, scale, shift, False)


class MinkowskiReLU(nn.Module):
    def __init__(self, inplace=False):
        super().__init__()

    def forward(self, x: SparseTensor) -> SparseTensor:
        return _affine(x, None, None, True)


def cat(*tensors: SparseTensor) -> SparseTensor:
    """ME.cat (minkunet.py:192): channel concatenation of tensors on the same coordinate map."""
    assert len({t.level for t in tensors}) == 1
    return SparseTensor(torch.cat([t.F for t in tensors], dim=1), coordinate_manager=tensors[0].coordinate_manager,
                        level=tensors[0].level)


def _add(a: SparseTensor, b: SparseTensor) -> SparseTensor:
    return SparseTensor(a.F + b.F, coordinate_manager=a.coordinate_manager, level=a.level)


SparseTensor.__add__ = _add
SparseTensor.__iadd__ = _add


class MinkowskiUnion(nn.Module):
    """ME.MinkowskiUnion (util.py:98-99): union of the coordinate sets of tensors that share a coordinate manager, the
    features of coincident coordinates add.  Rows come in first-occurrence order over the arguments."""

    def forward(self, *args):
        if len(args) < 2 or any(a._coords is None for a in args):
            raise NotImplementedError("MinkowskiUnion takes SparseTensors built from integer coordinates (util.py:86-99)")
        if any(a.coordinate_manager is not args[0].coordinate_manager for a in args):
            raise RuntimeError("MinkowskiUnion: all inputs must share one coordinate manager")
        coords = torch.cat([a._coords for a in args], dim=0)
        feats = torch.cat([a.F for a in args], dim=0)
        uc, uf = _unique_rows(coords, feats, "sum")
        return SparseTensor._from_set(uc, uf, args[0].coordinate_manager)


class MinkowskiPruning(nn.Module):
    """ME.MinkowskiPruning (util.py:105-106): keep the rows whose mask entry is true."""

    def forward(self, x: SparseTensor, mask: torch.Tensor):
        if x._coords is None:
            raise NotImplementedError("MinkowskiPruning takes SparseTensors built from integer coordinates (util.py:86-106)")
        mask = mask.to(device=x.F.device, dtype=torch.bool).reshape(-1)
        assert mask.numel() == len(x)
        return SparseTensor._from_set(x._coords[mask], x.F[mask], x.coordinate_manager)


class _Utils(types.ModuleType):
    @staticmethod
    def kaiming_normal_(tensor, a=0, mode="fan_in", nonlinearity="leaky_relu"):
        from .models import kaiming_normal_ as impl
        return impl(tensor, mode=mode, nonlinearity=nonlinearity)


utils = _Utils("MinkowskiEngine.utils")


class BasicBlock(nn.Module):
    """ME ``modules.resnet_block.BasicBlock`` (mirrored at c_ws/src/mapmos/scripts/minkunet.py:31-82)."""
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, dilation=1, downsample=None, bn_momentum=0.1, dimension=-1):
        super().__init__()
        assert dimension > 0
        self.conv1 = MinkowskiConvolution(inplanes, planes, kernel_size=3, stride=stride, dilation=dilation,
                                          dimension=dimension)
        self.norm1 = MinkowskiBatchNorm(planes, momentum=bn_momentum)
        self.conv2 = MinkowskiConvolution(planes, planes, kernel_size=3, stride=1, dilation=dilation, dimension=dimension)
        self.norm2 = MinkowskiBatchNorm(planes, momentum=bn_momentum)
        self.relu = MinkowskiReLU(inplace=True)
        self.downsample = downsample

    def forward(self, x):
        residual = x
        out = self.relu(self.norm1(self.conv1(x)))
        out = self.norm2(self.conv2(out))
        if self.downsample is not None:
            residual = self.downsample(x)
        out = out + residual
        return self.relu(out)


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, *args, **kwargs):
        raise NotImplementedError("Bottleneck blocks are not used by CustomMinkUNet (MinkUNet14 = BasicBlock)")


def _unsupported(name):
    class _U(nn.Module):
        def __init__(self, *a, **k):
            raise NotImplementedError(f"ME.{name} is not on the SPS hot path")
    _U.__name__ = name
    return _U


for _n in ("MinkowskiInstanceNorm", "MinkowskiMaxPooling", "MinkowskiDropout", "MinkowskiGELU",
           "MinkowskiGlobalMaxPooling", "MinkowskiLinear", "MinkowskiSumPooling", "MinkowskiGlobalSumPooling"):
    globals()[_n] = _unsupported(_n)


def install():
    """Register this module as ``MinkowskiEngine`` (+ ``.utils``, ``.modules.resnet_block``)."""
    me = sys.modules[__name__]
    modules = types.ModuleType("MinkowskiEngine.modules")
    rb = types.ModuleType("MinkowskiEngine.modules.resnet_block")
    rb.BasicBlock, rb.Bottleneck = BasicBlock, Bottleneck
    modules.resnet_block = rb
    me.modules = modules
    sys.modules["MinkowskiEngine"] = me
    sys.modules["MinkowskiEngine.utils"] = utils
    sys.modules["MinkowskiEngine.modules"] = modules
    sys.modules["MinkowskiEngine.modules.resnet_block"] = rb
    return me
