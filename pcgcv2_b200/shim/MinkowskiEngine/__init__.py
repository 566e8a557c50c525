"""Drop-in ``MinkowskiEngine`` for the PCGCv2 hot path, backed by libpcgc (sm_100a CUDA).

NOT a port of MinkowskiEngine: only the operator surface the reference's
``autoencoder.py`` / ``pcc_model.py`` / ``coder.py`` / ``data_utils.py`` /
``trainer.py`` use is provided (SURVEY.md section 8 b), with the same names, keyword
arguments, parameter names/shapes (``kernel`` [K,Cin,Cout] -- 2-D for K=1 --
and ``bias`` [1,Cout], so the reference checkpoints load with ``strict=True``)
and error behaviour.  Put ``pcgcv2_b200/shim`` on ``sys.path`` (or call
``pcgcv2_b200.install_shims()``) and the reference files import this module as
``import MinkowskiEngine as ME`` unchanged.

Semantics follow SURVEY.md Appendix A (``oracle/sparse_ref.py`` is the checker):
row order of user-supplied coordinates is preserved (A.2), strided maps come out
in ascending Morton-key order, generative up-sampling lays children out as
``8*i + k`` (A.6), pruning is stable (A.9).  There is no CPU path: tensors on the
CPU raise ``ValueError``.
"""
from __future__ import annotations

import itertools
import math
from typing import Optional

import torch

from pcgcv2_b200 import ops as _ops
from . import utils  # noqa: F401  (ME.utils.sparse_collate)

__version__ = "0.5.4+pcgc.b200"


def _to_stride(s):
    if isinstance(s, (list, tuple)):
        assert len(set(s)) == 1, "anisotropic tensor strides are not supported"
        return int(s[0])
    return int(s)


class CoordinateMapKey:
    """Opaque handle of one coordinate map inside a CoordinateManager."""
    _ids = itertools.count()

    def __init__(self, tensor_stride: int):
        self._stride = int(tensor_stride)
        self._id = next(CoordinateMapKey._ids)

    def get_tensor_stride(self):
        return [self._stride] * 3

    def get_key(self):
        return ([self._stride] * 3, str(self._id))

    def __repr__(self):
        return f"coordinate map key:[{self._stride}, {self._stride}, {self._stride}]#{self._id}"


class _CoordMap:
    """keys + lazily built hash table / kernel maps of one coordinate set (owned by the manager)."""

    def __init__(self, keys: torch.Tensor, stride: int, table=None, sorted_keys=False):
        self.keys = keys
        self.stride = stride
        self._table = table
        self.sorted = sorted_keys
        self._nbr = None
        self._coords = None
        self.down = None          # (child key, child_rows, child_off) of the k2s2 stride map
        self.up = None            # child key of the generative k2s2 map

    def __len__(self):
        return self.keys.shape[0]

    @property
    def table(self):
        if self._table is None:
            self._table = _ops.HashTable(self.keys)
        return self._table

    @property
    def nbr(self):                # kernel map of every k=3 conv on this set (cached, Appendix A.8)
        if self._nbr is None:
            self._nbr = _ops.kernel_map_k3(self.keys, self.table)
        return self._nbr

    @property
    def coords(self):
        if self._coords is None:
            self._coords = _ops.unpack_keys(self.keys, self.stride)
        return self._coords


class CoordinateManager:
    def __init__(self, D: int = 3):
        assert D == 3, "only 3-D coordinates are supported"
        self.D = D
        self._maps = {}

    def _insert(self, cmap: _CoordMap) -> CoordinateMapKey:
        key = CoordinateMapKey(cmap.stride)
        self._maps[key] = cmap
        return key

    def _get(self, key: CoordinateMapKey) -> _CoordMap:
        try:
            return self._maps[key]
        except KeyError:
            raise RuntimeError(f"{key} does not belong to this coordinate manager") from None

    def get_coordinates(self, key):
        return self._get(key).coords

    def size(self, key):
        return len(self._get(key))

    def stride(self, key: CoordinateMapKey) -> CoordinateMapKey:
        """output map of a kernel_size=2, stride=2 convolution (Appendix A.5); cached."""
        cmap = self._get(key)
        if cmap.down is None:
            pk, rows, off, parent_of = _ops.stride_down(cmap.keys, keys_are_sorted=cmap.sorted, with_parent_of=True)
            child = self._insert(_CoordMap(pk, cmap.stride * 2, sorted_keys=True))
            cmap.down = (child, rows, off, parent_of)
        return cmap.down[0]

    def stride_region(self, key: CoordinateMapKey) -> CoordinateMapKey:
        """output map of the generative kernel_size=2, stride=2 transposed convolution (A.6); cached."""
        cmap = self._get(key)
        if cmap.up is None:
            assert cmap.stride % 2 == 0, "cannot up-sample a tensor_stride-1 tensor"
            cmap.up = self._insert(_CoordMap(_ops.upsample_keys(cmap.keys), cmap.stride // 2, sorted_keys=cmap.sorted))
        return cmap.up


class SparseTensor:
    def __init__(self, features: torch.Tensor, coordinates: Optional[torch.Tensor] = None, tensor_stride=1,
                 coordinate_map_key: Optional[CoordinateMapKey] = None,
                 coordinate_manager: Optional[CoordinateManager] = None, quantization_mode=None,
                 requires_grad=None, device=None):
        if not isinstance(features, torch.Tensor):
            raise ValueError("Features must be a torch.Tensor")
        if features.dim() != 2:
            raise ValueError(f"The feature should be a matrix, the dimension is {features.dim()}")
        if (coordinates is None) == (coordinate_map_key is None):
            raise ValueError("Provide exactly one of coordinates and coordinate_map_key")
        if device is None:
            device = features.device if coordinates is None or features.is_cuda else coordinates.device
        device = torch.device(device)
        if device.type != "cuda":
            raise ValueError("pcgcv2_b200's MinkowskiEngine has no CPU backend: pass device='cuda'")
        features = features.to(device)
        if features.dtype != torch.float32:
            raise ValueError("features must be float32")
        if coordinates is not None:
            if not isinstance(coordinates, torch.Tensor) or coordinates.dtype != torch.int32:
                raise ValueError("coordinates must be an int32 torch.Tensor (use ME.utils.sparse_collate)")
            if coordinates.dim() != 2 or coordinates.shape[1] != 4:
                raise ValueError("coordinates must be [N, 1+3] (batch index first)")
            if coordinates.shape[0] != features.shape[0]:
                raise ValueError("coordinates and features have different numbers of rows")
            stride = _to_stride(tensor_stride)
            coords = coordinates.to(device).contiguous()
            keys = _ops.pack_keys(coords, stride)
            table = _ops.HashTable(keys)
            if table.n_dup:                                   # A.2: one row per coordinate, first seen wins
                keep = table.keep_flags(keys)
                keys, features = _ops.prune(keep, keys, features)
                cmap = _CoordMap(keys, stride)
            else:
                cmap = _CoordMap(keys, stride, table=table)
                cmap._coords = coords
            coordinate_manager = coordinate_manager or CoordinateManager()
            coordinate_map_key = coordinate_manager._insert(cmap)
        else:
            if coordinate_manager is None:
                raise ValueError("coordinate_map_key needs its coordinate_manager")
            if coordinate_manager.size(coordinate_map_key) != features.shape[0]:
                raise ValueError("features do not match the coordinate map size")
        self._F = features
        self.coordinate_map_key = coordinate_map_key
        self.coordinate_manager = coordinate_manager
        if requires_grad is not None:
            self._F.requires_grad_(requires_grad)

    # -- reference-visible attributes -------------------------------------------------------
    @property
    def F(self):
        return self._F

    features = F

    @property
    def C(self):
        return self.coordinate_manager.get_coordinates(self.coordinate_map_key)

    coordinates = C

    @property
    def tensor_stride(self):
        return self.coordinate_map_key.get_tensor_stride()

    @property
    def device(self):
        return self._F.device

    @property
    def dtype(self):
        return self._F.dtype

    @property
    def D(self):
        return 3

    @property
    def shape(self):
        return self._F.shape

    def size(self):
        return self._F.shape

    def __len__(self):
        return self._F.shape[0]

    @property
    def requires_grad(self):
        return self._F.requires_grad

    @property
    def _cmap(self) -> _CoordMap:
        return self.coordinate_manager._get(self.coordinate_map_key)

    @property
    def _batchwise_row_indices(self):
        b = self.C[:, 0]
        nb = int(b.max().item()) + 1 if len(b) else 0
        return [torch.nonzero(b == i, as_tuple=False).reshape(-1) for i in range(nb)]

    @property
    def decomposed_coordinates(self):
        c = self.C
        return [c[idx, 1:] for idx in self._batchwise_row_indices]

    @property
    def decomposed_features(self):
        return [self._F[idx] for idx in self._batchwise_row_indices]

    def detach(self):
        return SparseTensor(self._F.detach(), coordinate_map_key=self.coordinate_map_key,
                            coordinate_manager=self.coordinate_manager)

    def _like(self, feats):
        return SparseTensor(feats, coordinate_map_key=self.coordinate_map_key,
                            coordinate_manager=self.coordinate_manager)

    def _same_map(self, other, what):
        if not isinstance(other, SparseTensor):
            raise TypeError(f"{what}: expected a SparseTensor")
        if other.coordinate_manager is not self.coordinate_manager or \
                other.coordinate_map_key is not self.coordinate_map_key:
            raise ValueError(f"{what}: the operands live on different coordinate maps")

    def __add__(self, other):
        if isinstance(other, SparseTensor):
            self._same_map(other, "SparseTensor + SparseTensor")
            return self._like(self._F + other._F)
        return self._like(self._F + other)

    def __sub__(self, other):
        if isinstance(other, SparseTensor):
            self._same_map(other, "SparseTensor - SparseTensor")
            return self._like(self._F - other._F)
        return self._like(self._F - other)

    def __repr__(self):
        return (f"SparseTensor(\n  coordinates={self.C}\n  features={self._F}\n  {self.coordinate_map_key}"
                f"  spatial dimension=3)")


def cat(*sparse_tensors):
    """column concatenation of tensors on one coordinate map (autoencoder.py:55)."""
    if len(sparse_tensors) == 1 and isinstance(sparse_tensors[0], (list, tuple)):
        sparse_tensors = tuple(sparse_tensors[0])
    first = sparse_tensors[0]
    for t in sparse_tensors[1:]:
        first._same_map(t, "ME.cat")
    return first._like(torch.cat([t.F for t in sparse_tensors], dim=1))


# ---- autograd: forward = libpcgc kernels; backward = libpcgc kernels -------------
class _ConvK3Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats, kernel, bias, nbr, packed=None):
        ctx.save_for_backward(feats, kernel)
        ctx.nbr, ctx.has_bias = nbr, bias is not None
        if packed is not None and feats.stride(0) % 4 == 0 and feats.data_ptr() % 16 == 0:
            return _ops.conv_k3_packed(feats, nbr, packed, bias)            # 3xTF32 tensor-core kernel
        return _ops.conv_k3(feats, nbr, kernel, bias)

    @staticmethod
    def backward(ctx, go):
        feats, kernel = ctx.saved_tensors
        go = go.contiguous()
        gi = gw = gb = None
        if ctx.needs_input_grad[0]:      # stride-1 maps are symmetric: a forward conv with W'[k] = W[26-k]^T
            gi = _ops.conv_k3(go, ctx.nbr, kernel.flip(0).transpose(1, 2).contiguous())
        if ctx.needs_input_grad[1]:
            gw = _ops.conv_bwd_weight(feats, ctx.nbr, go, 27, kernel.shape[1], kernel.shape[2])
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = _ops.colsum(go)
        return gi, gw, gb, None, None


class _ConvK1Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats, kernel, bias):
        ctx.save_for_backward(feats, kernel)
        ctx.has_bias = bias is not None
        return _ops.conv_k1(feats, kernel, bias)

    @staticmethod
    def backward(ctx, go):
        feats, kernel = ctx.saved_tensors
        go = go.contiguous()
        gi = gw = gb = None
        if ctx.needs_input_grad[0]:
            gi = _ops.conv_k1(go, kernel.t().contiguous())
        if ctx.needs_input_grad[1]:
            gw = _ops.conv_bwd_weight(feats, None, go, 1, kernel.shape[0], kernel.shape[1]).view_as(kernel)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = _ops.colsum(go)
        return gi, gw, gb


class _ConvDownFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats, kernel, bias, keys, rows, off, parent_of):
        ctx.save_for_backward(feats, kernel)
        ctx.maps, ctx.has_bias = (keys, parent_of), bias is not None
        return _ops.conv_k2s2(feats, keys, rows, off, kernel, bias)

    @staticmethod
    def backward(ctx, go):
        feats, kernel = ctx.saved_tensors
        keys, parent_of = ctx.maps
        go = go.contiguous()
        gi, gw = _ops.conv_k2s2_bwd(feats, keys, parent_of, go, kernel, need_input_grad=ctx.needs_input_grad[0])
        gb = _ops.colsum(go) if ctx.has_bias and ctx.needs_input_grad[2] else None
        return gi, gw, gb, None, None, None, None


class _ConvUpFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats, kernel, bias):
        ctx.save_for_backward(feats, kernel)
        ctx.has_bias = bias is not None
        return _ops.convT_k2s2(feats, kernel, bias)

    @staticmethod
    def backward(ctx, go):
        feats, kernel = ctx.saved_tensors
        go = go.contiguous()
        gi, gw = _ops.convT_k2s2_bwd(feats, go, kernel, need_input_grad=ctx.needs_input_grad[0])
        gb = _ops.colsum(go) if ctx.has_bias and ctx.needs_input_grad[2] else None
        return gi, gw, gb


class _PruneFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats, mask, keys):
        new_keys, out = _ops.prune(mask, keys, feats)
        ctx.kept = torch.nonzero(mask, as_tuple=False).reshape(-1)
        ctx.n = feats.shape[0]
        ctx.mark_non_differentiable(new_keys)
        return out, new_keys

    @staticmethod
    def backward(ctx, go, _):
        gi = torch.zeros((ctx.n, go.shape[1]), dtype=go.dtype, device=go.device)
        gi.index_copy_(0, ctx.kept, go.contiguous())          # backward scatters grads to the kept rows (A.9)
        return gi, None, None


class _ConvBase(torch.nn.Module):
    _transpose = False

    def __init__(self, in_channels, out_channels, kernel_size=-1, stride=1, dilation=1, bias=False,
                 kernel_generator=None, expand_coordinates=False, convolution_mode=None, dimension=None):
        super().__init__()
        if dimension != 3:
            raise ValueError("dimension must be 3")
        if kernel_generator is not None or _to_stride(dilation) != 1:
            raise NotImplementedError("custom kernel generators / dilation are outside the PCGCv2 hot path")
        self.in_channels, self.out_channels = int(in_channels), int(out_channels)
        self.kernel_size, self.stride_, self.dimension = _to_stride(kernel_size), _to_stride(stride), 3
        supported = {(3, 1), (1, 1)} if not self._transpose else set()
        supported |= {(2, 2)}
        if (self.kernel_size, self.stride_) not in supported:
            raise NotImplementedError(
                f"kernel_size={self.kernel_size}, stride={self.stride_} is outside the PCGCv2 hot path "
                f"(supported: k3 s1, k1 s1, k2 s2)")
        kvol = self.kernel_size ** 3
        shape = (self.in_channels, self.out_channels) if kvol == 1 else (kvol, self.in_channels, self.out_channels)
        self.kernel = torch.nn.Parameter(torch.empty(shape, dtype=torch.float32))
        self.bias = torch.nn.Parameter(torch.empty((1, self.out_channels), dtype=torch.float32)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):                                # Appendix A.7
        kvol = self.kernel_size ** 3
        n = (self.out_channels if self._transpose else self.in_channels) * kvol
        stdv = 1.0 / math.sqrt(n)
        with torch.no_grad():
            self.kernel.uniform_(-stdv, stdv)
            if self.bias is not None:
                self.bias.uniform_(-stdv, stdv)

    def extra_repr(self):
        return (f"in={self.in_channels}, out={self.out_channels}, kernel_size=[{self.kernel_size}]*3, "
                f"stride=[{self.stride_}]*3, dilation=[1, 1, 1]")

    def _check(self, x):
        if not isinstance(x, SparseTensor):
            raise TypeError("input must be a SparseTensor")
        if x.F.shape[1] != self.in_channels:
            raise ValueError(f"input has {x.F.shape[1]} channels, the layer expects {self.in_channels}")
        if self.kernel.device != x.device:
            raise RuntimeError("module parameters and input are on different devices (call model.to(device))")


class MinkowskiConvolution(_ConvBase):
    """k=3 s=1, k=1 s=1 and k=2 s=2 convolutions (autoencoder.py:13-48,71-134,162-234)."""

    def _packed_weights(self):
        """k=3 weights in tensor-core fragment order, re-packed whenever the parameter changes."""
        tag = (self.kernel._version, self.kernel.data_ptr())
        if getattr(self, "_packed_tag", None) != tag:
            pw = _ops.PackedK3(self.kernel.detach())
            self._packed, self._packed_tag = (pw if pw.packed is not None else None), tag
        return self._packed

    def forward(self, x: SparseTensor) -> SparseTensor:
        self._check(x)
        cm, cmap = x.coordinate_manager, x._cmap
        if self.kernel_size == 1:
            return x._like(_ConvK1Fn.apply(x.F, self.kernel, self.bias))
        if self.kernel_size == 3:
            return x._like(_ConvK3Fn.apply(x.F, self.kernel, self.bias, cmap.nbr, self._packed_weights()))
        out_key = cm.stride(x.coordinate_map_key)
        _, rows, off, parent_of = cmap.down
        out = _ConvDownFn.apply(x.F, self.kernel, self.bias, cmap.keys, rows, off, parent_of)
        return SparseTensor(out, coordinate_map_key=out_key, coordinate_manager=cm)


class MinkowskiGenerativeConvolutionTranspose(_ConvBase):
    """generative k=2 s=2 transposed convolution (autoencoder.py:155,182,209; Appendix A.6)."""
    _transpose = True

    def forward(self, x: SparseTensor) -> SparseTensor:
        self._check(x)
        cm = x.coordinate_manager
        out_key = cm.stride_region(x.coordinate_map_key)
        return SparseTensor(_ConvUpFn.apply(x.F, self.kernel, self.bias), coordinate_map_key=out_key,
                            coordinate_manager=cm)


class MinkowskiConvolutionTranspose(MinkowskiGenerativeConvolutionTranspose):
    """k=2 s=2 transposed convolution; without a target map it generates all 8 children, exactly
    like the generative variant the reference uses."""


class MinkowskiReLU(torch.nn.Module):
    def __init__(self, inplace=False):
        super().__init__()
        self.inplace = inplace

    def forward(self, x: SparseTensor) -> SparseTensor:
        if self.inplace and not (torch.is_grad_enabled() and x.F.requires_grad):
            return x._like(torch.relu_(x.F))
        return x._like(torch.relu(x.F))


class MinkowskiPruning(torch.nn.Module):
    """keep the rows where mask is True, order preserved (autoencoder.py:237,247; Appendix A.9)."""

    def forward(self, x: SparseTensor, mask: torch.Tensor) -> SparseTensor:
        if mask.dtype != torch.bool or mask.dim() != 1 or mask.shape[0] != len(x):
            raise ValueError("mask must be a bool vector with one entry per row")
        cmap = x._cmap
        feats, keys = _PruneFn.apply(x.F, mask.to(x.device), cmap.keys)
        key = x.coordinate_manager._insert(_CoordMap(keys, cmap.stride, sorted_keys=cmap.sorted))
        return SparseTensor(feats, coordinate_map_key=key, coordinate_manager=x.coordinate_manager)
