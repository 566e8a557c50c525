"""Host-side operators over libpcgc (C ABI in include/pcgc.h).

PyTorch is used for device memory, streams and the caching allocator only; every
computation below is a call into the hand-written sm_100a kernels.  Keys are kept
in ``torch.int64`` tensors holding the uint64 bit pattern described in pcgc.h.
"""
from __future__ import annotations

import threading

import numpy as np
import torch

from . import _lib
from ._lib import check

EPI_RELU = 1
EB_PPC = 48


def _p(t):
    return None if t is None else t.data_ptr()


_TLS = threading.local()   # .stream: raw handle of the thread's current stream, looked up once per pipeline pass (stream_scope)


def _stream():
    h = getattr(_TLS, "stream", None)
    return h if h is not None else torch.cuda.current_stream().cuda_stream


class stream_scope:
    """with ops.stream_scope(): ... -- every operator this thread calls inside launches on the stream that is current on
    entry without asking torch for it again (about 2 us per launch); do not switch streams inside."""

    def __enter__(self):
        self.prev = getattr(_TLS, "stream", None)
        _TLS.stream = torch.cuda.current_stream().cuda_stream
        return self

    def __exit__(self, *exc):
        _TLS.stream = self.prev
        return False


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise ValueError("pcgcv2_b200 operators run on CUDA tensors only (there is no CPU path)")


def _ws(nbytes, device):
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


# ------------------------------------------------------------------ coordinates

def pack_keys(coords: torch.Tensor, tensor_stride: int) -> torch.Tensor:
    """int32 [N,4] (b,x,y,z) -> int64 [N] Morton keys (validates range; one sync)."""
    _need_cuda(coords)
    if coords.dtype != torch.int32:
        raise ValueError("coordinates must be int32")
    coords = coords.contiguous()
    n = coords.shape[0]
    keys = torch.empty(n, dtype=torch.int64, device=coords.device)
    err = torch.zeros(1, dtype=torch.int32, device=coords.device)
    check(_lib.lib().pcgc_pack_keys(_p(coords), n, int(tensor_stride), _p(keys), _p(err), _stream()), "pcgc_pack_keys")
    if n and int(err.item()):
        raise ValueError("coordinates out of range: need 0 <= c/stride <= %d, batch <= 126, c %% stride == 0"
                         % ((1 << 19) - 1))
    return keys


def pack_keys_async(coords: torch.Tensor, tensor_stride: int, err: torch.Tensor, batch: int = 0, hint_bits: int = 0) -> torch.Tensor:
    """int32 [N,4] (b,x,y,z) or [N,3] (x,y,z; all rows in ``batch``) -> int64 [N] Morton keys WITHOUT reading the range flag
    back: ``err`` (device int32 [1], zeroed by the caller) is raised on a bad coordinate (bit 0) or, for [N,3] input with
    ``hint_bits`` > 0, on a coordinate / stride >= 2^hint_bits (bit 1); the caller reads it with its next synchronising read."""
    _need_cuda(coords)
    if coords.dtype != torch.int32:
        raise ValueError("coordinates must be int32")
    coords = coords.contiguous()
    n = coords.shape[0]
    keys = torch.empty(n, dtype=torch.int64, device=coords.device)
    if coords.shape[1] == 3:
        check(_lib.lib().pcgc_pack_keys3(_p(coords), n, int(tensor_stride), int(batch), int(hint_bits), _p(keys), _p(err), _stream()),
              "pcgc_pack_keys3")
    else:
        check(_lib.lib().pcgc_pack_keys(_p(coords), n, int(tensor_stride), _p(keys), _p(err), _stream()), "pcgc_pack_keys")
    return keys


def scale_coords(coords: torch.Tensor, factor: float) -> torch.Tensor:
    """int32 coordinate values -> round_half_even(float32(v) * float32(factor)) (scale_sparse_tensor, data_utils.py:112-118);
    duplicates are NOT removed here."""
    _need_cuda(coords)
    if coords.dtype != torch.int32:
        raise ValueError("coordinates must be int32")
    coords = coords.contiguous()
    out = torch.empty_like(coords)
    check(_lib.lib().pcgc_scale_coords(_p(coords), coords.numel(), float(factor), _p(out), _stream()), "pcgc_scale_coords")
    return out


def unpack_keys(keys: torch.Tensor, tensor_stride: int) -> torch.Tensor:
    n = keys.shape[0]
    coords = torch.empty((n, 4), dtype=torch.int32, device=keys.device)
    check(_lib.lib().pcgc_unpack_keys(_p(keys), n, int(tensor_stride), _p(coords), _stream()), "pcgc_unpack_keys")
    return coords


class HashTable:
    """key -> row table of one coordinate map."""

    def __init__(self, keys: torch.Tensor):
        n = keys.shape[0]
        L = _lib.lib()
        self.cap = int(L.pcgc_hash_capacity(n))
        self.tkeys = torch.empty(self.cap, dtype=torch.int64, device=keys.device)
        self.tvals = torch.empty(self.cap, dtype=torch.int32, device=keys.device)
        self._ndup = torch.zeros(1, dtype=torch.int32, device=keys.device)
        check(L.pcgc_hash_build(_p(keys), n, _p(self.tkeys), _p(self.tvals), self.cap, _p(self._ndup), _stream()),
              "pcgc_hash_build")

    @property
    def n_dup(self) -> int:
        return int(self._ndup.item())

    def keep_flags(self, keys):
        keep = torch.empty(keys.shape[0], dtype=torch.uint8, device=keys.device)
        check(_lib.lib().pcgc_hash_keep_flags(_p(keys), keys.shape[0], _p(self.tkeys), _p(self.tvals), self.cap,
                                              _p(keep), _stream()), "pcgc_hash_keep_flags")
        return keep

    def contains(self, query):
        found = torch.empty(query.shape[0], dtype=torch.uint8, device=query.device)
        check(_lib.lib().pcgc_hash_contains(_p(query), query.shape[0], _p(self.tkeys), self.cap, _p(found), _stream()),
              "pcgc_hash_contains")
        return found.bool()


def kernel_map_k3(keys: torch.Tensor, table: HashTable, count_pairs=False):
    """-> nbr int32 [27, N] (offset-major, -1 = missing) [, device int64 pair count]."""
    n = keys.shape[0]
    nbr = torch.empty((27, n), dtype=torch.int32, device=keys.device)
    npairs = torch.zeros(1, dtype=torch.int64, device=keys.device) if count_pairs else None
    check(_lib.lib().pcgc_kernel_map_k3(_p(keys), n, _p(table.tkeys), _p(table.tvals), table.cap, _p(nbr), _p(npairs),
                                        _stream()), "pcgc_kernel_map_k3")
    return (nbr, npairs) if count_pairs else nbr


def stride_down(keys: torch.Tensor, keys_are_sorted=False, with_parent_of=False):
    """-> (parent_keys [P], child_rows int32 [N], child_off int32 [P+1][, parent_of int32 [N]]);
    one sync for P."""
    n = keys.shape[0]
    L = _lib.lib()
    dev = keys.device
    parent = torch.empty(n, dtype=torch.int64, device=dev)
    n_par = torch.zeros(1, dtype=torch.int32, device=dev)
    rows = torch.empty(n, dtype=torch.int32, device=dev)
    off = torch.empty(n + 1, dtype=torch.int32, device=dev)
    parent_of = torch.empty(n, dtype=torch.int32, device=dev) if with_parent_of else None
    nbytes = L.pcgc_stride_down_ws_bytes(n)
    ws = _ws(nbytes, dev)
    check(L.pcgc_stride_down(_p(keys), n, int(bool(keys_are_sorted)), _p(parent), _p(n_par), _p(rows), _p(off),
                             _p(parent_of), _p(ws), nbytes, _stream()), "pcgc_stride_down")
    p = int(n_par.item())
    if with_parent_of:
        return parent[:p], rows, off[:p + 1], parent_of
    return parent[:p], rows, off[:p + 1]


def parent_info(child_keys: torch.Tensor, child_off: torch.Tensor) -> torch.Tensor:
    """per parent: (first child row << 8) | occupancy byte (children sorted)."""
    n_par = child_off.shape[0] - 1
    info = torch.empty(n_par, dtype=torch.int64, device=child_keys.device)
    check(_lib.lib().pcgc_parent_info(_p(child_keys), _p(child_off), n_par, _p(info), _stream()), "pcgc_parent_info")
    return info


def kernel_map_k3_from_parent(parent_nbr: torch.Tensor, n: int, child_keys=None, parent_of=None, info=None):
    """child kernel map [27, n] from the parent's [27, P]; info=None: full octets (n == 8 P)."""
    n_par = parent_nbr.shape[1]
    nbr = torch.empty((27, n), dtype=torch.int32, device=parent_nbr.device)
    check(_lib.lib().pcgc_kernel_map_k3_from_parent(_p(child_keys), _p(parent_of), _p(info), _p(parent_nbr), n_par, n,
                                                    _p(nbr), _stream()), "pcgc_kernel_map_k3_from_parent")
    return nbr


def conv_k3_ones_from_parent(parent_nbr, child_keys, parent_of, info, weight, bias, relu=True, want_f32=True, want_h2=False,
                             overflow=None):
    """encoder.conv0 on constant-one features straight from the parent's kernel map (no child map): weight [27, 1, 16]
    -> (fp32 [n, 16] or None, h2 [n, 16] or None)."""
    n, n_par, cout = child_keys.shape[0], parent_nbr.shape[1], weight.shape[2]
    assert weight.shape[0] == 27 and weight.shape[1] == 1 and weight.is_contiguous()
    out = torch.empty((n, cout), dtype=torch.float32, device=child_keys.device) if want_f32 else None
    out_h2 = torch.empty((n, cout), dtype=torch.int32, device=child_keys.device) if want_h2 else None
    check(_lib.lib().pcgc_conv_k3_ones_from_parent_fwd(_p(child_keys), _p(parent_of), _p(info), _p(parent_nbr), n_par, n, _p(weight),
                                                       _p(bias), cout, _p(out), cout, _p(out_h2), cout, EPI_RELU if relu else 0,
                                                       _p(overflow), _stream()), "pcgc_conv_k3_ones_from_parent_fwd")
    return out, out_h2


def upsample_keys(keys: torch.Tensor) -> torch.Tensor:
    n = keys.shape[0]
    out = torch.empty(8 * n, dtype=torch.int64, device=keys.device)
    check(_lib.lib().pcgc_upsample_keys(_p(keys), n, _p(out), _stream()), "pcgc_upsample_keys")
    return out


def argsort_u64(keys: torch.Tensor, end_bit=64):
    """stable ascending argsort of uint64 bit patterns -> (sorted keys, order int32)."""
    n = keys.shape[0]
    L = _lib.lib()
    ks = torch.empty_like(keys)
    order = torch.empty(n, dtype=torch.int32, device=keys.device)
    nbytes = L.pcgc_argsort_ws_bytes(n)
    ws = _ws(nbytes, keys.device)
    check(L.pcgc_argsort_u64(_p(keys), n, int(end_bit), _p(ks), _p(order), _p(ws), nbytes, _stream()), "pcgc_argsort_u64")
    return ks, order


# ------------------------------------------------------------------ convolutions

def _feat(t):
    _need_cuda(t)
    if t.dtype != torch.float32:
        raise ValueError("features must be float32")
    if t.dim() != 2 or t.stride(1) != 1:
        t = t.contiguous()
    return t


def _out_slice(out, n, cout, device):
    if out is None:
        out = torch.empty((n, cout), dtype=torch.float32, device=device)
    assert out.shape[0] == n and out.shape[1] == cout and out.stride(1) == 1
    return out


def conv_k3(feats, nbr, weight, bias=None, residual=None, relu=False, out=None):
    feats = _feat(feats)
    n, cin = feats.shape
    cout = weight.shape[2]
    assert weight.shape[0] == 27 and weight.shape[1] == cin and weight.is_contiguous()
    out = _out_slice(out, n, cout, feats.device)
    residual = None if residual is None else _feat(residual)
    check(_lib.lib().pcgc_conv_k3_fwd(_p(feats), feats.stride(0), _p(nbr), n, _p(weight), _p(bias), cin, cout,
                                      _p(residual), 0 if residual is None else residual.stride(0), _p(out), out.stride(0),
                                      EPI_RELU if relu else 0, _stream()), "pcgc_conv_k3_fwd")
    return out


class PackedK3:
    """k=3 weights pre-packed for the tensor-core kernel (None if the shape has no such kernel)."""

    def __init__(self, weight: torch.Tensor):
        assert weight.dim() == 3 and weight.shape[0] == 27 and weight.is_contiguous()
        self.cin, self.cout = int(weight.shape[1]), int(weight.shape[2])
        n = int(_lib.lib().pcgc_conv_k3_packed_floats(self.cin, self.cout))
        self.packed = None
        if n:
            self.packed = torch.empty(n, dtype=torch.float32, device=weight.device)
            check(_lib.lib().pcgc_conv_k3_pack_weights(_p(weight), self.cin, self.cout, _p(self.packed), _stream()),
                  "pcgc_conv_k3_pack_weights")


def conv_k3_packed(feats, nbr, pw: PackedK3, bias=None, residual=None, relu=False, out=None):
    feats = _feat(feats)
    n, cin = feats.shape
    assert cin == pw.cin and pw.packed is not None
    out = _out_slice(out, n, pw.cout, feats.device)
    residual = None if residual is None else _feat(residual)
    check(_lib.lib().pcgc_conv_k3_fwd_packed(_p(feats), feats.stride(0), _p(nbr), n, _p(pw.packed), _p(bias), cin, pw.cout,
                                             _p(residual), 0 if residual is None else residual.stride(0), _p(out),
                                             out.stride(0), EPI_RELU if relu else 0, _stream()), "pcgc_conv_k3_fwd_packed")
    return out


class PackedK3Octet:
    """k=3 weights packed for the full-octet kernels (None if the shape has no such kernel)."""

    def __init__(self, weight: torch.Tensor):
        assert weight.dim() == 3 and weight.shape[0] == 27 and weight.is_contiguous()
        self.cin, self.cout = int(weight.shape[1]), int(weight.shape[2])
        n = int(_lib.lib().pcgc_conv_k3_octet_packed_floats(self.cin, self.cout))
        self.packed = None
        if n:
            self.packed = torch.empty(n, dtype=torch.float32, device=weight.device)
            check(_lib.lib().pcgc_conv_k3_octet_pack_weights(_p(weight), self.cin, self.cout, _p(self.packed), _stream()),
                  "pcgc_conv_k3_octet_pack_weights")


def conv_k3_octet(feats, parent_nbr, pw: PackedK3Octet, bias=None, residual=None, relu=False, out=None):
    """k=3 convolution on the 8-child expansion of a parent set: feats [8P, cin] (row 8i+c = child c of
    parent row i), parent_nbr int32 [27, P] = the PARENT set's kernel map."""
    feats = _feat(feats)
    n, cin = feats.shape
    n_par = parent_nbr.shape[1]
    assert cin == pw.cin and pw.packed is not None and n == 8 * n_par and parent_nbr.is_contiguous()
    out = _out_slice(out, n, pw.cout, feats.device)
    residual = None if residual is None else _feat(residual)
    check(_lib.lib().pcgc_conv_k3_octet_fwd(_p(feats), feats.stride(0), _p(parent_nbr), n_par, _p(pw.packed), _p(bias), cin,
                                            pw.cout, _p(residual), 0 if residual is None else residual.stride(0), _p(out),
                                            out.stride(0), EPI_RELU if relu else 0, _stream()), "pcgc_conv_k3_octet_fwd")
    return out


# ---- pre-split half-precision ("h2") features: see include/pcgc.h and csrc/conv_h2.cuh
def _h2(t):
    _need_cuda(t)
    if t.dtype != torch.int32 or t.dim() != 2 or t.stride(1) != 1:
        raise ValueError("h2 features are int32 [n, c] tensors with unit column stride")
    return t


def split_h2(feats, out=None, overflow=None):
    """fp32 [n, c] -> h2 int32 [n, c] (x = f16 hi + f16 lo, 16-byte groups of four channels)."""
    feats = _feat(feats)
    n, c = feats.shape
    if out is None:
        out = torch.empty((n, c), dtype=torch.int32, device=feats.device)
    assert out.shape == feats.shape and out.dtype == torch.int32 and out.stride(1) == 1
    check(_lib.lib().pcgc_split_h2(_p(feats), feats.stride(0), n, c, _p(out), out.stride(0), _p(overflow), _stream()), "pcgc_split_h2")
    return out


def join_h2(h2, out=None):
    h2 = _h2(h2)
    n, c = h2.shape
    out = _out_slice(out, n, c, h2.device)
    check(_lib.lib().pcgc_join_h2(_p(h2), h2.stride(0), n, c, _p(out), out.stride(0), _stream()), "pcgc_join_h2")
    return out


class PackedK3H2:
    """k=3 weights split into f16 hi/lo MMA fragments after a power-of-two scale (None: no h2 kernel for the shape)."""

    @staticmethod
    def supported(cin, cout) -> bool:
        return int(_lib.lib().pcgc_conv_k3_h2_packed_words(int(cin), int(cout))) > 0

    @property
    def gather(self) -> bool:
        """a child-map (gather) kernel exists for the shape; cin = 4 has the full-octet kernel only."""
        return self.cin % 16 == 0 or self.cin == 8

    def __init__(self, weight: torch.Tensor):
        assert weight.dim() == 3 and weight.shape[0] == 27 and weight.is_contiguous()
        self.cin, self.cout = int(weight.shape[1]), int(weight.shape[2])
        n = int(_lib.lib().pcgc_conv_k3_h2_packed_words(self.cin, self.cout))
        self.packed = None
        if n:
            wmax = float(weight.abs().max())
            k = int(np.floor(np.log2(16384.0 / wmax))) if wmax > 0 and np.isfinite(wmax) else 0
            k = max(-24, min(24, k))
            self.scale, self.inv_scale = float(2.0 ** k), float(2.0 ** -k)
            self.packed = torch.empty(n, dtype=torch.int32, device=weight.device)
            check(_lib.lib().pcgc_conv_k3_h2_pack_weights(_p(weight), self.cin, self.cout, self.scale, _p(self.packed), _stream()),
                  "pcgc_conv_k3_h2_pack_weights")


def conv_k3_h2(feats_h2, nbr, pw: PackedK3H2, bias=None, residual=None, relu=False, out=None, out_h2=None, want_f32=True,
               want_h2=False, overflow=None):
    """k=3 convolution over h2 features -> (fp32 out or None, h2 out or None)."""
    x = _h2(feats_h2)
    n, cin = x.shape
    assert cin == pw.cin and pw.packed is not None and nbr.shape[0] == 27 and nbr.shape[1] == n and nbr.is_contiguous()
    if want_f32 or out is not None:
        out = _out_slice(out, n, pw.cout, x.device)
    if want_h2 and out_h2 is None:
        out_h2 = torch.empty((n, pw.cout), dtype=torch.int32, device=x.device)
    if out_h2 is not None:
        assert out_h2.shape[0] == n and out_h2.shape[1] == pw.cout and out_h2.dtype == torch.int32 and out_h2.stride(1) == 1
    residual = None if residual is None else _feat(residual)
    check(_lib.lib().pcgc_conv_k3_h2_fwd(_p(x), x.stride(0), _p(nbr), n, _p(pw.packed), pw.inv_scale, _p(bias), cin, pw.cout,
                                         _p(residual), 0 if residual is None else residual.stride(0), _p(out),
                                         0 if out is None else out.stride(0), _p(out_h2), 0 if out_h2 is None else out_h2.stride(0),
                                         EPI_RELU if relu else 0, _p(overflow), _stream()), "pcgc_conv_k3_h2_fwd")
    return out, out_h2


class PackedK3Wide:
    """k=3 weights as the swizzled tcgen05 operand tiles [W_hi ; W_lo] of csrc/conv_wide.cuh (None: no kernel for the shape)."""

    @staticmethod
    def supported(cin, cout) -> bool:
        return int(_lib.lib().pcgc_conv_k3_wide_packed_bytes(int(cin), int(cout))) > 0

    def __init__(self, weight: torch.Tensor):
        assert weight.dim() == 3 and weight.shape[0] == 27 and weight.is_contiguous()
        self.cin, self.cout = int(weight.shape[1]), int(weight.shape[2])
        n = int(_lib.lib().pcgc_conv_k3_wide_packed_bytes(self.cin, self.cout))
        self.packed = None
        if n:
            self.scale, self.inv_scale = _h2_scale(weight)
            self.packed = torch.empty(n, dtype=torch.uint8, device=weight.device)
            check(_lib.lib().pcgc_conv_k3_wide_pack_weights(_p(weight), self.cin, self.cout, self.scale, _p(self.packed), _stream()),
                  "pcgc_conv_k3_wide_pack_weights")


def conv_k3_wide(feats_h2, nbr, pw: PackedK3Wide, bias=None, residual=None, relu=False, out=None, out_h2=None, want_f32=True,
                 want_h2=False, overflow=None):
    """k=3 convolution over h2 features on tcgen05 / TMA -> (fp32 out or None, h2 out or None)."""
    x = _h2(feats_h2)
    n, cin = x.shape
    assert cin == pw.cin and pw.packed is not None and nbr.shape[0] == 27 and nbr.shape[1] == n and nbr.is_contiguous()
    if want_f32 or out is not None:
        out = _out_slice(out, n, pw.cout, x.device)
    if want_h2 and out_h2 is None:
        out_h2 = torch.empty((n, pw.cout), dtype=torch.int32, device=x.device)
    if out_h2 is not None:
        assert out_h2.shape[0] == n and out_h2.shape[1] == pw.cout and out_h2.dtype == torch.int32 and out_h2.stride(1) == 1
    residual = None if residual is None else _feat(residual)
    check(_lib.lib().pcgc_conv_k3_wide_fwd(_p(x), x.stride(0), _p(nbr), n, _p(pw.packed), pw.inv_scale, _p(bias), cin, pw.cout,
                                           _p(residual), 0 if residual is None else residual.stride(0), _p(out),
                                           0 if out is None else out.stride(0), _p(out_h2), 0 if out_h2 is None else out_h2.stride(0),
                                           EPI_RELU if relu else 0, _p(overflow), _stream()), "pcgc_conv_k3_wide_fwd")
    return out, out_h2


class PackedK3OctetTc05:
    """k=3 weights as the resident tcgen05 operand tiles of csrc/conv_octet_tc05.cuh (None: no kernel for the shape)."""

    @staticmethod
    def supported(cin, cout) -> bool:
        return int(_lib.lib().pcgc_conv_k3_octet_tc05_packed_bytes(int(cin), int(cout))) > 0

    def __init__(self, weight: torch.Tensor):
        assert weight.dim() == 3 and weight.shape[0] == 27 and weight.is_contiguous()
        self.cin, self.cout = int(weight.shape[1]), int(weight.shape[2])
        n = int(_lib.lib().pcgc_conv_k3_octet_tc05_packed_bytes(self.cin, self.cout))
        self.packed = None
        if n:
            self.scale, self.inv_scale = _h2_scale(weight)
            self.packed = torch.empty(n, dtype=torch.uint8, device=weight.device)
            check(_lib.lib().pcgc_conv_k3_octet_tc05_pack_weights(_p(weight), self.cin, self.cout, self.scale, _p(self.packed), _stream()),
                  "pcgc_conv_k3_octet_tc05_pack_weights")


def conv_k3_octet_tc05(feats_h2, parent_nbr, pw: PackedK3OctetTc05, bias=None, residual=None, relu=False, out=None, out_h2=None,
                       want_f32=True, want_h2=False, overflow=None):
    """k=3 convolution over h2 features on the 8-child expansion of a parent set, tcgen05 with descriptor-addressed halos."""
    x = _h2(feats_h2)
    n, cin = x.shape
    n_par = parent_nbr.shape[1]
    assert cin == pw.cin and pw.packed is not None and n == 8 * n_par and parent_nbr.is_contiguous()
    if want_f32 or out is not None:
        out = _out_slice(out, n, pw.cout, x.device)
    if want_h2 and out_h2 is None:
        out_h2 = torch.empty((n, pw.cout), dtype=torch.int32, device=x.device)
    if out_h2 is not None:
        assert out_h2.shape[0] == n and out_h2.shape[1] == pw.cout and out_h2.dtype == torch.int32 and out_h2.stride(1) == 1
    residual = None if residual is None else _feat(residual)
    check(_lib.lib().pcgc_conv_k3_octet_tc05_fwd(_p(x), x.stride(0), _p(parent_nbr), n_par, _p(pw.packed), pw.inv_scale, _p(bias), cin,
                                                 pw.cout, _p(residual), 0 if residual is None else residual.stride(0), _p(out),
                                                 0 if out is None else out.stride(0), _p(out_h2),
                                                 0 if out_h2 is None else out_h2.stride(0), EPI_RELU if relu else 0, _p(overflow),
                                                 _stream()), "pcgc_conv_k3_octet_tc05_fwd")
    return out, out_h2


_SUPPORT = {}


def _h2_scale(weight):
    wmax = float(weight.abs().max())
    k = int(np.floor(np.log2(16384.0 / wmax))) if wmax > 0 and np.isfinite(wmax) else 0
    k = max(-24, min(24, k))
    return float(2.0 ** k), float(2.0 ** -k)


class PackedDownH2:
    """k=2 stride-2 weights [8, cin, cout] split into f16 hi/lo fragments for pcgc_conv_k2s2_h2_fwd (None: no kernel)."""

    def __init__(self, weight: torch.Tensor):
        assert weight.dim() == 3 and weight.shape[0] == 8 and weight.is_contiguous()
        self.cin, self.cout = int(weight.shape[1]), int(weight.shape[2])
        n = int(_lib.lib().pcgc_conv_k2s2_h2_packed_words(self.cin, self.cout))
        self.packed = None
        if n:
            self.scale, self.inv_scale = _h2_scale(weight)
            self.packed = torch.empty(n, dtype=torch.int32, device=weight.device)
            check(_lib.lib().pcgc_conv_k2s2_h2_pack_weights(_p(weight), self.cin, self.cout, self.scale, _p(self.packed), _stream()),
                  "pcgc_conv_k2s2_h2_pack_weights")


def child_map_k2(child_keys, parent_of, n_parents):
    """int32 [8, n_parents]: row of child k = key & 7 of every parent, -1 where the child is absent."""
    cm = torch.empty((8, n_parents), dtype=torch.int32, device=child_keys.device)
    check(_lib.lib().pcgc_child_map_k2(_p(child_keys), _p(parent_of), child_keys.shape[0], int(n_parents), _p(cm), _stream()),
          "pcgc_child_map_k2")
    return cm


def conv_k2s2_h2(feats_h2, child_map, pw: PackedDownH2, bias=None, relu=False, want_f32=True, want_h2=True, overflow=None):
    """k=2 stride-2 convolution over h2 features on the tensor cores -> (fp32 out or None, h2 out or None)."""
    x = _h2(feats_h2)
    n_par = child_map.shape[1]
    assert x.shape[1] == pw.cin and pw.packed is not None and child_map.shape[0] == 8 and child_map.is_contiguous()
    out = torch.empty((n_par, pw.cout), dtype=torch.float32, device=x.device) if want_f32 else None
    out_h2 = torch.empty((n_par, pw.cout), dtype=torch.int32, device=x.device) if want_h2 else None
    check(_lib.lib().pcgc_conv_k2s2_h2_fwd(_p(x), x.stride(0), _p(child_map), n_par, _p(pw.packed), pw.inv_scale, _p(bias), pw.cin,
                                           pw.cout, _p(out), pw.cout, _p(out_h2), pw.cout, EPI_RELU if relu else 0, _p(overflow),
                                           _stream()), "pcgc_conv_k2s2_h2_fwd")
    return out, out_h2


class PackedUpH2:
    """transposed k=2 stride-2 weights [8, cin, cout] as the dense [cin, 8*cout] operand, split and packed (None: no kernel)."""

    def __init__(self, weight: torch.Tensor, bias: torch.Tensor | None):
        assert weight.dim() == 3 and weight.shape[0] == 8 and weight.is_contiguous()
        self.cin, self.cout = int(weight.shape[1]), int(weight.shape[2])
        n = int(_lib.lib().pcgc_convT_k2s2_h2_packed_words(self.cin, self.cout))
        self.packed = None
        if n:
            self.scale, self.inv_scale = _h2_scale(weight)
            self.packed = torch.empty(n, dtype=torch.int32, device=weight.device)
            ws = torch.empty(8 * self.cin * self.cout, dtype=torch.float32, device=weight.device)
            check(_lib.lib().pcgc_convT_k2s2_h2_pack_weights(_p(weight), self.cin, self.cout, self.scale, _p(ws), _p(self.packed),
                                                             _stream()), "pcgc_convT_k2s2_h2_pack_weights")
            self.bias8 = None if bias is None else bias.reshape(1, -1).repeat(1, 8).contiguous()


def convT_k2s2_h2(feats_h2, pw: PackedUpH2, relu=False, want_f32=True, want_h2=True, overflow=None):
    """generative transposed k=2 stride-2 convolution as one dense tensor-core product -> ([8n, cout] fp32, h2)."""
    x = _h2(feats_h2)
    n = x.shape[0]
    assert x.shape[1] == pw.cin and pw.packed is not None
    out = torch.empty((8 * n, pw.cout), dtype=torch.float32, device=x.device) if want_f32 else None
    out_h2 = torch.empty((8 * n, pw.cout), dtype=torch.int32, device=x.device) if want_h2 else None
    check(_lib.lib().pcgc_convT_k2s2_h2_fwd(_p(x), x.stride(0), n, _p(pw.packed), pw.inv_scale, _p(pw.bias8), pw.cin, pw.cout, _p(out),
                                            pw.cout, _p(out_h2), pw.cout, EPI_RELU if relu else 0, _p(overflow), _stream()),
          "pcgc_convT_k2s2_h2_fwd")
    return out, out_h2


def octet_h2_supported(cin, cout) -> bool:
    key = ("octet_h2", int(cin), int(cout))
    if key not in _SUPPORT:
        _SUPPORT[key] = bool(_lib.lib().pcgc_conv_k3_octet_h2_supported(int(cin), int(cout)))
    return _SUPPORT[key]


def conv_k3_octet_h2(feats_h2, parent_nbr, pw: PackedK3H2, bias=None, residual=None, relu=False, out=None, out_h2=None,
                     want_f32=True, want_h2=False, overflow=None):
    """k=3 convolution over h2 features on the 8-child expansion of a parent set (parent_nbr int32 [27, P])."""
    x = _h2(feats_h2)
    n, cin = x.shape
    n_par = parent_nbr.shape[1]
    assert cin == pw.cin and pw.packed is not None and n == 8 * n_par and parent_nbr.is_contiguous()
    if want_f32 or out is not None:
        out = _out_slice(out, n, pw.cout, x.device)
    if want_h2 and out_h2 is None:
        out_h2 = torch.empty((n, pw.cout), dtype=torch.int32, device=x.device)
    if out_h2 is not None:
        assert out_h2.shape[0] == n and out_h2.shape[1] == pw.cout and out_h2.dtype == torch.int32 and out_h2.stride(1) == 1
    residual = None if residual is None else _feat(residual)
    check(_lib.lib().pcgc_conv_k3_octet_h2_fwd(_p(x), x.stride(0), _p(parent_nbr), n_par, _p(pw.packed), pw.inv_scale, _p(bias), cin,
                                               pw.cout, _p(residual), 0 if residual is None else residual.stride(0), _p(out),
                                               0 if out is None else out.stride(0), _p(out_h2),
                                               0 if out_h2 is None else out_h2.stride(0), EPI_RELU if relu else 0, _p(overflow),
                                               _stream()), "pcgc_conv_k3_octet_h2_fwd")
    return out, out_h2


def h2out_supported(kind: str, cin, cout) -> bool:
    """kind in {"k1", "down", "up"}: the layer's kernel can write the h2 copy of its output in its epilogue."""
    key = (kind, int(cin), int(cout))
    if key not in _SUPPORT:
        _SUPPORT[key] = bool(_lib.lib().pcgc_conv_h2out_supported({"k1": 1, "down": 2, "up": 3}[kind], int(cin), int(cout)))
    return _SUPPORT[key]


def _h2_out(out_h2, n, cout, device):
    if out_h2 is None:
        out_h2 = torch.empty((n, cout), dtype=torch.int32, device=device)
    assert out_h2.shape[0] == n and out_h2.shape[1] == cout and out_h2.dtype == torch.int32 and out_h2.stride(1) == 1
    return out_h2


def conv_k1(feats, weight, bias=None, residual=None, relu=False, out=None, out_h2=None, overflow=None):
    """out_h2: int32 [n, cout] tensor (or True to allocate one) that receives the h2 copy of the output -> (out, out_h2)."""
    feats = _feat(feats)
    n, cin = feats.shape
    assert weight.dim() == 2 and weight.shape[0] == cin and weight.is_contiguous()
    cout = weight.shape[1]
    out = _out_slice(out, n, cout, feats.device)
    residual = None if residual is None else _feat(residual)
    if out_h2 is not None:
        out_h2 = _h2_out(None if out_h2 is True else out_h2, n, cout, feats.device)
        check(_lib.lib().pcgc_conv_k1_fwd_h2out(_p(feats), feats.stride(0), n, _p(weight), _p(bias), cin, cout, _p(residual),
                                                0 if residual is None else residual.stride(0), _p(out), out.stride(0),
                                                _p(out_h2), out_h2.stride(0), EPI_RELU if relu else 0, _p(overflow), _stream()),
              "pcgc_conv_k1_fwd_h2out")
        return out, out_h2
    check(_lib.lib().pcgc_conv_k1_fwd(_p(feats), feats.stride(0), n, _p(weight), _p(bias), cin, cout, _p(residual),
                                      0 if residual is None else residual.stride(0), _p(out), out.stride(0),
                                      EPI_RELU if relu else 0, _stream()), "pcgc_conv_k1_fwd")
    return out


def conv_k2s2(feats, in_keys, child_rows, child_off, weight, bias=None, relu=False, out=None, out_h2=None, overflow=None):
    feats = _feat(feats)
    cin = feats.shape[1]
    n_par = child_off.shape[0] - 1
    cout = weight.shape[2]
    assert weight.shape[0] == 8 and weight.shape[1] == cin and weight.is_contiguous()
    out = _out_slice(out, n_par, cout, feats.device)
    if out_h2 is not None:
        out_h2 = _h2_out(None if out_h2 is True else out_h2, n_par, cout, feats.device)
        check(_lib.lib().pcgc_conv_k2s2_fwd_h2out(_p(feats), feats.stride(0), _p(in_keys), _p(child_rows), _p(child_off), n_par,
                                                  _p(weight), _p(bias), cin, cout, _p(out), out.stride(0), _p(out_h2),
                                                  out_h2.stride(0), EPI_RELU if relu else 0, _p(overflow), _stream()),
              "pcgc_conv_k2s2_fwd_h2out")
        return out, out_h2
    check(_lib.lib().pcgc_conv_k2s2_fwd(_p(feats), feats.stride(0), _p(in_keys), _p(child_rows), _p(child_off), n_par,
                                        _p(weight), _p(bias), cin, cout, _p(out), out.stride(0),
                                        EPI_RELU if relu else 0, _stream()), "pcgc_conv_k2s2_fwd")
    return out


def convT_k2s2(feats, weight, bias=None, relu=False, out=None, out_h2=None, overflow=None):
    feats = _feat(feats)
    n, cin = feats.shape
    cout = weight.shape[2]
    assert weight.shape[0] == 8 and weight.shape[1] == cin and weight.is_contiguous()
    out = _out_slice(out, 8 * n, cout, feats.device)
    if out_h2 is not None:
        out_h2 = _h2_out(None if out_h2 is True else out_h2, 8 * n, cout, feats.device)
        check(_lib.lib().pcgc_convT_k2s2_fwd_h2out(_p(feats), feats.stride(0), n, _p(weight), _p(bias), cin, cout, _p(out),
                                                   out.stride(0), _p(out_h2), out_h2.stride(0), EPI_RELU if relu else 0,
                                                   _p(overflow), _stream()), "pcgc_convT_k2s2_fwd_h2out")
        return out, out_h2
    check(_lib.lib().pcgc_convT_k2s2_fwd(_p(feats), feats.stride(0), n, _p(weight), _p(bias), cin, cout, _p(out),
                                         out.stride(0), EPI_RELU if relu else 0, _stream()), "pcgc_convT_k2s2_fwd")
    return out


# ------------------------------------------------------------------ backward passes (row a16)

def conv_bwd_weight(feats, nbr, grad_out, kvol, cin, cout):
    """grad of `kernel` for k=3 (kvol 27, nbr given) / k=1 (kvol 1, nbr None) -> [kvol, cin, cout]."""
    feats, grad_out = _feat(feats), _feat(grad_out)
    gw = torch.empty((kvol, cin, cout), dtype=torch.float32, device=feats.device)
    ws, nbytes = _bwd_ws(feats.shape[0], kvol, cin, cout, feats.device)
    check(_lib.lib().pcgc_conv_bwd_weight(_p(feats), feats.stride(0), _p(nbr), feats.shape[0], kvol, _p(grad_out),
                                          grad_out.stride(0), cin, cout, _p(gw), _p(ws), nbytes, _stream()), "pcgc_conv_bwd_weight")
    return gw


def _bwd_ws(n, kvol, cin, cout, device):
    """workspace of the deterministic two-pass weight gradient (partials per block, summed in block order)."""
    nbytes = int(_lib.lib().pcgc_conv_bwd_weight_ws_bytes(n, kvol, cin, cout))
    return _ws(max(nbytes, 16), device), nbytes


def conv_k2s2_bwd(feats, in_keys, parent_of, grad_out, weight, need_input_grad=True):
    feats, grad_out = _feat(feats), _feat(grad_out)
    n, cin = feats.shape
    cout = weight.shape[2]
    gi = torch.empty((n, cin), dtype=torch.float32, device=feats.device) if need_input_grad else None
    gw = torch.empty_like(weight)
    ws, nbytes = _bwd_ws(n, 8, cin, cout, feats.device)
    check(_lib.lib().pcgc_conv_k2s2_bwd(_p(feats), feats.stride(0), _p(in_keys), _p(parent_of), n, _p(grad_out),
                                        grad_out.stride(0), _p(weight), cin, cout, _p(gi), cin, _p(gw), _p(ws), nbytes,
                                        _stream()), "pcgc_conv_k2s2_bwd")
    return gi, gw


def convT_k2s2_bwd(feats, grad_out, weight, need_input_grad=True):
    feats, grad_out = _feat(feats), _feat(grad_out)
    n, cin = feats.shape
    cout = weight.shape[2]
    gi = torch.empty((n, cin), dtype=torch.float32, device=feats.device) if need_input_grad else None
    gw = torch.empty_like(weight)
    ws, nbytes = _bwd_ws(n, 8, cin, cout, feats.device)
    check(_lib.lib().pcgc_convT_k2s2_bwd(_p(feats), feats.stride(0), n, _p(grad_out), grad_out.stride(0), _p(weight), cin,
                                         cout, _p(gi), cin, _p(gw), _p(ws), nbytes, _stream()), "pcgc_convT_k2s2_bwd")
    return gi, gw


def colsum(x):
    x = _feat(x)
    out = torch.empty((1, x.shape[1]), dtype=torch.float32, device=x.device)
    nbytes = int(_lib.lib().pcgc_colsum_ws_bytes())
    ws = _ws(nbytes, x.device)
    check(_lib.lib().pcgc_colsum(_p(x), x.stride(0), x.shape[0], x.shape[1], _p(out), _p(ws), nbytes, _stream()), "pcgc_colsum")
    return out


def bce_isin(logits: torch.Tensor, cand_keys: torch.Tensor, gt_table: "HashTable", need_grad=True):
    """fused loss.py:7-15: -> (sum of BCE-with-logits over the isin mask in bits, float32 [1] on the device;
    d(sum)/d(logit) float32 [n] or None; the 0/1 mask uint8 [n])."""
    _need_cuda(logits, cand_keys)
    x = logits.reshape(logits.shape[0], -1) if logits.dim() > 1 else logits
    assert x.dim() == 1 or x.shape[1] == 1, "one logit per candidate row"
    n = x.shape[0]
    ld = x.stride(0) if n > 1 else 1
    L = _lib.lib()
    loss = torch.empty(1, dtype=torch.float32, device=x.device)
    grad = torch.empty(n, dtype=torch.float32, device=x.device) if need_grad else None
    target = torch.empty(n, dtype=torch.uint8, device=x.device)
    nbytes = int(L.pcgc_bce_isin_ws_bytes())
    ws = _ws(nbytes, x.device)
    check(L.pcgc_bce_isin(_p(x), max(int(ld), 1), _p(cand_keys), n, _p(gt_table.tkeys), gt_table.cap, _p(loss), _p(grad), _p(target),
                          _p(ws), nbytes, _stream()), "pcgc_bce_isin")
    return loss, grad, target


# ------------------------------------------------------------------ selection / pruning

def topk_mask(logits: torch.Tensor, k: int) -> torch.Tensor:
    """bool [n]: True on the k largest entries of logits ([n] or [n,1] / strided column)."""
    _need_cuda(logits)
    if logits.dim() == 2:
        assert logits.shape[1] == 1
        ld = logits.stride(0)
    else:
        ld = logits.stride(0) if logits.numel() > 1 else 1
    n = logits.shape[0]
    L = _lib.lib()
    mask = torch.empty(n, dtype=torch.uint8, device=logits.device)
    nbytes = L.pcgc_topk_mask_ws_bytes(n)
    ws = _ws(nbytes, logits.device)
    check(L.pcgc_topk_mask(_p(logits), max(int(ld), 1), n, int(k), _p(mask), _p(ws), nbytes, _stream()), "pcgc_topk_mask")
    return mask.bool()


def prune(mask: torch.Tensor, keys, feats, nbr=None, n_kept_hint=None):
    """stable compaction -> (keys_kept, feats_kept[, nbr_kept]); one sync for the count unless the
    caller already knows it (``n_kept_hint``: e.g. an exact top-k mask keeps exactly k rows)."""
    feats = _feat(feats)
    n, c = feats.shape
    L = _lib.lib()
    dev = feats.device
    m8 = mask.to(torch.uint8) if mask.dtype != torch.uint8 else mask
    m8 = m8.contiguous()
    keys_out = torch.empty(n, dtype=torch.int64, device=dev) if keys is not None else None
    feats_out = torch.empty((n, c), dtype=torch.float32, device=dev)
    n_kept = torch.zeros(1, dtype=torch.int32, device=dev)
    nbytes = L.pcgc_prune_ws_bytes(n)
    ws = _ws(nbytes, dev)
    nbr_out = torch.empty(27 * n, dtype=torch.int32, device=dev) if nbr is not None else None
    check(L.pcgc_prune(_p(m8), n, _p(keys), _p(feats), feats.stride(0), c, _p(keys_out), _p(feats_out), c, _p(n_kept),
                       _p(nbr), _p(nbr_out), _p(ws), nbytes, _stream()), "pcgc_prune")
    k = int(n_kept.item()) if n_kept_hint is None else int(n_kept_hint)
    if nbr is not None:
        return (None if keys is None else keys_out[:k]), feats_out[:k], nbr_out[:27 * k].view(27, k)
    return (None if keys is None else keys_out[:k]), feats_out[:k]


# ------------------------------------------------------------------ entropy bottleneck

def pack_eb_params(matrices, biases, factors, device) -> torch.Tensor:
    """12 reference tensors -> float32 [C, 48] block (layout in pcgc.h)."""
    c = matrices[0].shape[0]
    cols = [matrices[0].reshape(c, 3), matrices[1].reshape(c, 9), matrices[2].reshape(c, 9), matrices[3].reshape(c, 3),
            biases[0].reshape(c, 3), biases[1].reshape(c, 3), biases[2].reshape(c, 3), biases[3].reshape(c, 1),
            factors[0].reshape(c, 3), factors[1].reshape(c, 3), factors[2].reshape(c, 3), factors[3].reshape(c, 1)]
    packed = torch.cat([t.detach().float().to(device) for t in cols] +
                       [torch.zeros((c, EB_PPC - 44), dtype=torch.float32, device=device)], dim=1)
    return packed.contiguous()


def eb_likelihood(values: torch.Tensor, params: torch.Tensor) -> torch.Tensor:
    values = _feat(values).contiguous()
    n, c = values.shape
    out = torch.empty_like(values)
    check(_lib.lib().pcgc_eb_likelihood_fwd(_p(values), n, c, _p(params), _p(out), _stream()), "pcgc_eb_likelihood_fwd")
    return out


def eb_likelihood_bwd(values: torch.Tensor, params: torch.Tensor, grad_lik: torch.Tensor, need_values_grad=True):
    """-> (grad_values [n,c] or None, grad_params [c,48] wrt the RAW packed parameters)."""
    values, grad_lik = _feat(values).contiguous(), _feat(grad_lik).contiguous()
    n, c = values.shape
    gv = torch.empty_like(values) if need_values_grad else None
    gp = torch.empty_like(params)
    check(_lib.lib().pcgc_eb_likelihood_bwd(_p(values), n, c, _p(params), _p(grad_lik), _p(gv), _p(gp), _stream()),
          "pcgc_eb_likelihood_bwd")
    return gv, gp


def unpack_eb_param_grads(gp: torch.Tensor):
    """[C,48] packed gradient -> (matrices, biases, factors) lists shaped like the reference parameters."""
    c = gp.shape[0]
    cut = lambda a, b, shape: gp[:, a:b].reshape((c,) + shape)
    m = [cut(0, 3, (3, 1)), cut(3, 12, (3, 3)), cut(12, 21, (3, 3)), cut(21, 24, (1, 3))]
    b = [cut(24, 27, (3, 1)), cut(27, 30, (3, 1)), cut(30, 33, (3, 1)), cut(33, 34, (1, 1))]
    f = [cut(34, 37, (3, 1)), cut(37, 40, (3, 1)), cut(40, 43, (3, 1)), cut(43, 44, (1, 1))]
    return m, b, f


def eb_cdf_table(params: torch.Tensor, min_v: int, max_v: int):
    """-> (cdf_float [C, L+1] float32, cdf_u16 [C, L+1] int16-bits) on the device."""
    c = params.shape[0]
    lp = int(max_v) - int(min_v) + 2
    cdf = torch.empty((c, lp), dtype=torch.float32, device=params.device)
    u16 = torch.empty((c, lp), dtype=torch.int16, device=params.device)
    check(_lib.lib().pcgc_eb_cdf_table(_p(params), c, int(min_v), int(max_v), _p(cdf), _p(u16), _stream()),
          "pcgc_eb_cdf_table")
    return cdf, u16


_MM_INIT = {}


def eb_quantize_async(feats: torch.Tensor):
    """round -> (symbols int16 [N,C], minmax int32 [2]) on the device, no synchronisation."""
    feats = _feat(feats).contiguous()
    count = feats.numel()
    init = _MM_INIT.get(feats.device)
    if init is None:                       # (a scalar assignment would be a pageable H2D copy: a stream synchronisation)
        init = _MM_INIT[feats.device] = torch.tensor([2 ** 31 - 1, -2 ** 31], dtype=torch.int32, device=feats.device)
    mm = init.clone()
    L = _lib.lib()
    check(L.pcgc_eb_round_minmax(_p(feats), count, _p(mm), _stream()), "pcgc_eb_round_minmax")
    sym = torch.empty(feats.shape, dtype=torch.int16, device=feats.device)
    check(L.pcgc_eb_symbols(_p(feats), count, _p(mm), _p(sym), _stream()), "pcgc_eb_symbols")
    return sym, mm


def eb_quantize(feats: torch.Tensor):
    """round -> (symbols int16 [N,C] on device, min, max); one sync for min/max."""
    sym, mm = eb_quantize_async(feats)
    lo, hi = mm.tolist()
    return sym, lo, hi


# ------------------------------------------------------------------ range coder (host)

def rc_encode_float(cdf_float: np.ndarray, sym: np.ndarray) -> bytes:
    """cdf_float float32 [T, Lp] (symbol i uses row i % T), sym int16 [n]."""
    cdf_float = np.ascontiguousarray(cdf_float, dtype=np.float32)
    sym = np.ascontiguousarray(sym, dtype=np.int16).reshape(-1)
    T, lp = cdf_float.shape
    cap = 4 * sym.size + 64
    out = np.empty(cap, dtype=np.uint8)
    n = _lib.lib().pcgc_rc_encode_host(cdf_float.ctypes.data, T, lp, sym.ctypes.data, sym.size, out.ctypes.data, cap)
    check(n, "pcgc_rc_encode_host")
    if n > cap:
        out = np.empty(n, dtype=np.uint8)
        check(_lib.lib().pcgc_rc_encode_host(cdf_float.ctypes.data, T, lp, sym.ctypes.data, sym.size, out.ctypes.data, n))
    return out[:n].tobytes()


def rc_decode_float(cdf_float: np.ndarray, data: bytes, n_sym: int) -> np.ndarray:
    cdf_float = np.ascontiguousarray(cdf_float, dtype=np.float32)
    T, lp = cdf_float.shape
    buf = np.frombuffer(data, dtype=np.uint8)
    out = np.empty(n_sym, dtype=np.int16)
    check(_lib.lib().pcgc_rc_decode_host(cdf_float.ctypes.data, T, lp, buf.ctypes.data if buf.size else None, buf.size,
                                         out.ctypes.data, n_sym), "pcgc_rc_decode_host")
    return out


def rc_encode_u16(table: np.ndarray, sym: np.ndarray) -> bytes:
    table = np.ascontiguousarray(table).view(np.uint16)
    sym = np.ascontiguousarray(sym, dtype=np.int16).reshape(-1)
    T, lp = table.shape
    cap = 4 * sym.size + 64
    out = np.empty(cap, dtype=np.uint8)
    n = _lib.lib().pcgc_rc_encode_u16_host(table.ctypes.data, T, lp, sym.ctypes.data, sym.size, out.ctypes.data, cap)
    check(n, "pcgc_rc_encode_u16_host")
    if n > cap:
        out = np.empty(n, dtype=np.uint8)
        check(_lib.lib().pcgc_rc_encode_u16_host(table.ctypes.data, T, lp, sym.ctypes.data, sym.size, out.ctypes.data, n))
    return out[:n].tobytes()


def symbol_ranges(sym: torch.Tensor, table_u16: torch.Tensor) -> torch.Tensor:
    """per-symbol coding intervals on the device (pcgc_symbol_ranges): sym int16 [n] (row-major [N3, C]), table uint16-as-int16
    [C, L+1] on the same device -> int32 [n] holding c_low | (c_high - 1) << 16.  Raises on symbols outside the alphabet."""
    _need_cuda(sym, table_u16)
    sym = sym.contiguous().reshape(-1)
    assert sym.dtype == torch.int16 and table_u16.dtype in (torch.int16, torch.uint16) and table_u16.dim() == 2
    table_u16 = table_u16.contiguous()
    ranges = torch.empty(sym.shape[0], dtype=torch.int32, device=sym.device)
    bad = torch.zeros(1, dtype=torch.int32, device=sym.device)
    check(_lib.lib().pcgc_symbol_ranges(_p(sym), sym.shape[0], _p(table_u16), table_u16.shape[0], table_u16.shape[1], _p(ranges),
                                        _p(bad), _stream()), "pcgc_symbol_ranges")
    if sym.shape[0] and int(bad.item()):
        raise ValueError("symbol outside the table's alphabet")
    return ranges


def rc_encode_ranges(ranges: np.ndarray) -> bytes:
    """the torchac-compatible stream from per-symbol intervals (host; pcgc_rc_encode_ranges_host)."""
    ranges = np.ascontiguousarray(ranges).view(np.uint32).reshape(-1)
    cap = 4 * ranges.size + 64
    out = np.empty(cap, dtype=np.uint8)
    n = _lib.lib().pcgc_rc_encode_ranges_host(ranges.ctypes.data, ranges.size, out.ctypes.data, cap)
    check(n, "pcgc_rc_encode_ranges_host")
    if n > cap:
        out = np.empty(n, dtype=np.uint8)
        check(_lib.lib().pcgc_rc_encode_ranges_host(ranges.ctypes.data, ranges.size, out.ctypes.data, n))
    return out[:n].tobytes()


def rc_decode_u16(table: np.ndarray, data: bytes, n_sym: int, out: np.ndarray | None = None) -> np.ndarray:
    table = np.ascontiguousarray(table).view(np.uint16)
    T, lp = table.shape
    buf = np.frombuffer(data, dtype=np.uint8)
    if out is None:
        out = np.empty(n_sym, dtype=np.int16)
    assert out.dtype == np.int16 and out.size == n_sym and out.flags.c_contiguous
    check(_lib.lib().pcgc_rc_decode_u16_host(table.ctypes.data, T, lp, buf.ctypes.data if buf.size else None, buf.size,
                                             out.ctypes.data, n_sym), "pcgc_rc_decode_u16_host")
    return out


# ------------------------------------------------------------------ ASCII PLY geometry I/O (host, row f2)

def ply_read_ascii(path, pinned=False) -> torch.Tensor:
    """ASCII PLY file -> int32 [n, 3] CPU tensor (pinned on request, so that ``.to(device, non_blocking=True)`` is one DMA).
    Line semantics of the reference's read_ply_ascii_geo (data_utils.py:19-34): see csrc/ply.cpp."""
    text = np.fromfile(path, dtype=np.uint8)
    L = _lib.lib()
    cap = check(L.pcgc_ply_count_lines_host(text.ctypes.data if text.size else None, text.size), "pcgc_ply_count_lines_host") if text.size else 0
    out = torch.empty((max(cap, 1), 3), dtype=torch.int32, pin_memory=bool(pinned))
    n = check(L.pcgc_ply_parse_ascii_host(text.ctypes.data, text.size, out.data_ptr(), cap), "pcgc_ply_parse_ascii_host") if text.size else 0
    return out[:n]


def ply_write_ascii(path, coords) -> int:
    """int [n, 3] host coordinates -> ASCII PLY file in the reference's layout (data_utils.py:36-48); returns the byte count."""
    c = np.ascontiguousarray(np.asarray(coords.cpu() if isinstance(coords, torch.Tensor) else coords).astype(np.int32)).reshape(-1, 3)
    buf = np.empty(160 + 36 * c.shape[0], dtype=np.uint8)
    n = check(_lib.lib().pcgc_ply_format_ascii_host(c.ctypes.data if c.size else None, c.shape[0], buf.ctypes.data, buf.size),
              "pcgc_ply_format_ascii_host")
    with open(path, "wb") as f:
        f.write(memoryview(buf[:n]))
    return int(n)
