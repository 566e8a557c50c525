"""Encode / decode pipeline of the PCGCv2 geometry codec on libpcgc (the path bench.py measures).

``Codec`` mirrors the reference's ``Coder`` (coder.py:73-112): ``encode`` runs the analysis
network, sorts the bottleneck into the canonical symbol order, quantises and range-codes the
features; ``decode`` inverts it and runs the synthesis network with top-k pruning.  The byte
layouts of ``_F.bin`` / ``_H.bin`` / ``_num_points.bin`` are the reference's (coder.py:49-55,
85-87).  The stride-8 coordinate side channel (``_C.bin``, coder.py:23-36)
is produced by a pluggable coordinate coder (``coords_coder.py``): the in-process octree coder by
default, or the reference's external ``tmc3`` for byte-identical G-PCC streams; it runs on a side
thread while this thread range-codes the features.

Unlike the per-operator shim, the pipeline keeps every coordinate set in ascending Morton-key
order (one radix sort of the input; stride-2 parents, 8-child expansion and stable pruning all
preserve it), fuses bias / ReLU / concat / residual into the convolution epilogues and never
leaves the device between layers.
"""
from __future__ import annotations

import ctypes
import os
import threading
from dataclasses import dataclass, field

import numpy as np
import torch

from concurrent.futures import ThreadPoolExecutor

from . import _lib, ops
from .coords_coder import OctreeCoordinateCoder
from .entropy_host import HostTable


@dataclass
class Stream:
    """the four pieces the reference writes to disk (coder.py:81-91)."""
    F: bytes
    H: bytes
    num_points: bytes
    coords: np.ndarray                  # int32 [N3, 3], stride-8 coordinates / 8, canonical order (what C encodes)
    stats: dict = field(default_factory=dict)
    C: bytes | None = None              # the coded coordinates (`_C.bin`, coder.py:23-29); None: no coordinate coder configured

    def bits(self, coords_bits: int = 0) -> int:
        """total stream size in bits: the four files of coder.py:169-170 (``coords_bits`` stands in for C when it is absent)."""
        return 8 * (len(self.F) + len(self.H) + len(self.num_points)) + (8 * len(self.C) if self.C is not None else coords_bits)


class _Level:
    """one coordinate set (sorted Morton keys) with its lazily built k3 kernel map.

    Maps are hierarchical: a set that knows its parent set (stride-2 coarser) derives its map from
    the parent's map with pure index arithmetic (``pcgc_kernel_map_k3_from_parent``); only a root
    set (the ~14 k-row bottleneck) is hashed.  A pruned set receives its map from ``pcgc_prune``."""

    def __init__(self, keys: torch.Tensor, stride: int, parent=None, parent_of=None, info=None, nbr=None):
        self.keys, self.stride = keys, stride
        self.parent, self.parent_of, self.info = parent, parent_of, info      # info None + parent => full octets
        self._nbr = nbr

    @property
    def full_octets(self):
        """this set is the 8-child expansion of ``parent`` (row 8i + c = child c of parent row i)."""
        return self.parent is not None and self.info is None

    def __len__(self):
        return self.keys.shape[0]

    @property
    def nbr(self):
        if self._nbr is None:
            if self.parent is None:
                self._nbr = ops.kernel_map_k3(self.keys, ops.HashTable(self.keys))
            else:
                self._nbr = ops.kernel_map_k3_from_parent(self.parent.nbr, len(self), self.keys, self.parent_of,
                                                          self.info)
        return self._nbr


class _F:
    """one feature tensor in one or both storage formats: ``f`` = float32 [n, c]; ``h`` = the pre-split
    half-precision copy (int32 [n, c], csrc/conv_h2.cuh) that the h2 k=3 kernels gather from."""
    __slots__ = ("f", "h")

    def __init__(self, f=None, h=None):
        self.f, self.h = f, h


class _FullOctets:
    full_octets = True


_FULL = _FullOctets()            # "a full-octet level" for kernel-routing questions asked before the level exists


_SPIN_LOCK = threading.Lock() if os.environ.get("PCGC_SYNC_LOCK", "0") == "1" else None


class Codec:
    WIDE_SHAPES = {(64, 64)}            # (cin, cout) of the k=3 layers routed to the tcgen05 / TMA kernel

    def __init__(self, state_dict, device="cuda", use_tensor_cores=True, use_octet_kernels=True, use_h2=True, fuse_irn=True,
                 coords_coder="octree", wide_shapes=None, merge_first=True, coord_bits=None, fuse_tail=True, dual_second=True, fuse_conv0=True):
        """``coords_coder``: "octree" (in-process, default), None (hand the coordinates over raw, no ``Stream.C``) or any
        object with ``encode(int32 [n,3]) -> bytes`` / ``decode(bytes) -> int32 [n,3]`` (e.g. ``Tmc3CoordinateCoder``)."""
        # ``coord_bits``: the caller's promise that every input coordinate is below 2^coord_bits (coder.py's --res: 10 for vox10).
        # The radix sorts then run over 3 * coord_bits key bits instead of 57; an input that breaks the promise raises a device
        # flag that is read with the pass's synchronising read, and the frame is coded again at full width (coord_bits_fallbacks).
        self.coord_bits = None if coord_bits is None else max(4, min(19, int(coord_bits)))
        self.coord_bits_fallbacks = 0
        self.dual_second = dual_second      # ... and conv0_1 with them: a 16-channel block is two launches (needs merge_first, fuse_tail)
        self.fuse_conv0 = fuse_conv0        # encoder.conv0 on its constant-one input straight from the parent's kernel map
        self.fuse_tail = fuse_tail          # conv1_1 (k=3) + ReLU + conv1_2 (k=1) of the 16-channel blocks in one kernel
        self.merge_first = merge_first      # conv0_0 + conv1_0 of the 16-channel blocks as one k=3 convolution (see _merged_first)
        self.coords_coder = OctreeCoordinateCoder() if coords_coder == "octree" else coords_coder
        self._side = ThreadPoolExecutor(max_workers=1, thread_name_prefix="pcgc-coords") if self.coords_coder is not None else None
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("pcgcv2_b200.Codec runs on a CUDA device only (there is no CPU path)")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        if wide_shapes is None:
            env = os.environ.get("PCGC_WIDE_SHAPES")
            wide_shapes = self.WIDE_SHAPES if env is None else (env if env in ("all", "none") else
                                                                {tuple(int(v) for v in t.split("x")) for t in env.split(",") if t})
        self.wide_shapes = wide_shapes          # "all" | "none" | set of (cin, cout): k=3 layers routed to the tcgen05 / TMA kernel
        with torch.cuda.device(self.device):             # weight packing launches on THIS device's current stream
            self._init(state_dict, use_tensor_cores, use_octet_kernels, use_h2, fuse_irn)

    def _init(self, state_dict, use_tensor_cores, use_octet_kernels, use_h2, fuse_irn):
        self.w = {k: v.detach().float().contiguous().to(self.device) for k, v in state_dict.items()
                  if k.startswith(("encoder.", "decoder."))}
        g = lambda name: [state_dict[f"entropy_bottleneck.{name}.{i}"] for i in range(4)]
        self.eb_params = ops.pack_eb_params(g("_matrices"), g("_biases"), g("_factors"), self.device)
        self.eb_host = HostTable(g("_matrices"), g("_biases"), g("_factors"))       # the range coder's table is host data
        self.channels = self.eb_params.shape[0]
        # k=3 weights pre-packed for the tensor-core kernel where one exists (cin >= 8)
        self.packed = {}
        if use_tensor_cores:
            for k, v in self.w.items():
                if k.endswith(".kernel") and v.dim() == 3 and v.shape[0] == 27:
                    pw = ops.PackedK3(v)
                    if pw.packed is not None:
                        self.packed[k[:-len(".kernel")]] = pw
        # synthesis-side k=3 weights packed for the full-octet kernels (halo in shared memory, parent's map)
        self.packed_octet = {}
        if use_octet_kernels:
            for k, v in self.w.items():
                if k.startswith("decoder.") and k.endswith(".kernel") and v.dim() == 3 and v.shape[0] == 27:
                    pw = ops.PackedK3Octet(v)
                    if pw.packed is not None:
                        self.packed_octet[k[:-len(".kernel")]] = pw
        # k=3 weights split into f16 hi/lo fragments for the pre-split half-precision kernels (cin in {16,32,64})
        self.packed_h2 = {}
        if use_h2 and use_tensor_cores:
            for k, v in self.w.items():
                if k.endswith(".kernel") and v.dim() == 3 and v.shape[0] == 27:
                    pw = ops.PackedK3H2(v)
                    if pw.packed is not None:
                        self.packed_h2[k[:-len(".kernel")]] = pw
        # cout = 64 does not fit the resident-weight h2 kernel: run it as four 16-wide output slices of the same gather
        self.packed_h2_slices = {}
        if use_h2 and use_tensor_cores:
            for k, v in self.w.items():
                name = k[:-len(".kernel")]
                if (k.endswith(".kernel") and v.dim() == 3 and v.shape[0] == 27 and name not in self.packed_h2
                        and v.shape[2] % 16 == 0 and ops.PackedK3H2.supported(v.shape[1], 16)):
                    self.packed_h2_slices[name] = [(ops.PackedK3H2(v[:, :, j:j + 16].contiguous()),
                                                    self.w[name + ".bias"][:, j:j + 16]) for j in range(0, v.shape[2], 16)]
        # tcgen05 / TMA kernel (csrc/conv_wide.cuh) for the shapes where it beats the mma.sync kernels (tools/bench_wide.py,
        # profiles/r02_wide_tcgen05.txt); PCGC_WIDE_SHAPES="64x64,32x32" | "all" | "none" overrides the routed set
        self.packed_wide = {}
        if use_h2 and use_tensor_cores:
            for k, v in self.w.items():
                if k.endswith(".kernel") and v.dim() == 3 and v.shape[0] == 27:
                    shape = (int(v.shape[1]), int(v.shape[2]))
                    routed = self.wide_shapes == "all" or (self.wide_shapes != "none" and shape in self.wide_shapes)
                    if routed and ops.PackedK3Wide.supported(*shape):
                        self.packed_wide[k[:-len(".kernel")]] = ops.PackedK3Wide(v)
        # k=2 stride-2 / transposed layers on the tensor cores (gather over the 8 child slots / one dense product)
        self.packed_down_h2, self.packed_up_h2 = {}, {}
        if use_h2 and use_tensor_cores:
            for k, v in self.w.items():
                name = k[:-len(".kernel")]
                if k.endswith(".kernel") and v.dim() == 3 and v.shape[0] == 8:
                    pw = ops.PackedDownH2(v) if ".down" in name else ops.PackedUpH2(v, self.w.get(name + ".bias"))
                    if pw.packed is not None:
                        (self.packed_down_h2 if ".down" in name else self.packed_up_h2)[name] = pw
        self._h2_on = bool(self.packed_h2)
        self._merged = {}
        self._sync_event, self._blocking_sync = None, os.environ.get("PCGC_BLOCKING_SYNC", "0") == "1"
        self.use_octet = use_octet_kernels
        self._overflow = torch.zeros(1, dtype=torch.int32, device=self.device)   # raised by an h2 producer: re-run in fp32
        self._bad = torch.zeros(1, dtype=torch.int32, device=self.device)        # raised by pack_keys: coordinate out of range
        self.h2_fallbacks = 0
        self.fuse_irn = fuse_irn        # InceptionResNet blocks through pcgc_irn_fwd (one C call per block)
        self._irn_plans = {}
        self._pinned = {}               # reusable pinned host staging buffers (decoded coordinates, symbols, flags)
        self._tables = {}               # (lo, hi) -> host CDF table
        self._dedupe_next = False
        self.keep_bottleneck = False    # True: Stream.stats["y_F"] = the float bottleneck [N3, 8] in symbol order (parity tests)
        self.record = None              # set to a dict to capture per-layer activations (parity tests)
        self.probe = {}                 # layer name -> list of (start, end) CUDA event pairs (bench.py roofline)

    # ---------------------------------------------------------------- layers
    def _rec(self, name, t, level=None):
        if self.record is not None:
            if isinstance(t, _F):
                t = t.f if t.f is not None else ops.join_h2(t.h)
            self.record[name] = (t.clone(), None if level is None else level.keys.clone(),
                                 None if level is None else level.stride)

    def _h(self, x: _F):
        """the h2 copy of x (split once, on first use, when the producer did not write it)."""
        if x.h is None:
            x.h = ops.split_h2(x.f, overflow=self._overflow)
        return x.h

    def _k3(self, name, x: _F, level, relu=False, residual=None, out=None, out_h=None, want_f=True, want_h=False) -> _F:
        """one k=3 layer.  ``out`` / ``out_h``: column slices to write into (fp32 / h2); ``want_h``: the consumer is an
        h2 kernel, so an h2 producer writes that format in its epilogue (others are split lazily by ``_h``)."""
        aligned = x.f is None or x.f.stride(0) % 4 == 0
        pwide = self.packed_wide.get(name) if self._h2_on else None
        ph = self.packed_h2.get(name) if (self._h2_on and pwide is None) else None
        oh = ph if (ph is not None and level.full_octets and self.use_octet and ops.octet_h2_supported(ph.cin, ph.cout)) else None
        if ph is not None and oh is None and not ph.gather:
            ph = None
        po = self.packed_octet.get(name) if (pwide is None and oh is None and level.full_octets and aligned and x.f is not None) else None
        if po is not None:
            ph = None
        pw = self.packed.get(name) if (aligned and x.f is not None) else None
        ev = self.probe.get(name)
        sl = self.packed_h2_slices.get(name) if (self._h2_on and po is None and pwide is None) else None
        if ev is not None:
            nbr = level.parent.nbr if (po is not None or oh is not None) else level.nbr   # keep the (cached) map build outside the probe
            if ph is not None or pwide is not None or sl is not None:
                self._h(x)
            start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            start.record()
        y = _F()
        if pwide is not None:    # tcgen05.mma over TMA-streamed weight tiles, accumulators in tensor memory
            y.f, y.h = ops.conv_k3_wide(self._h(x), level.nbr, pwide, self.w[name + ".bias"], residual=residual, relu=relu, out=out,
                                        out_h2=out_h, want_f32=want_f, want_h2=want_h and pwide.cout % 4 == 0, overflow=self._overflow)
        elif oh is not None:     # 8-child expansion + pre-split f16 features: halo of h2 rows, LDS.128 -> HMMA.16816
            y.f, y.h = ops.conv_k3_octet_h2(self._h(x), level.parent.nbr, oh, self.w[name + ".bias"], residual=residual, relu=relu,
                                            out=out, out_h2=out_h, want_f32=want_f, want_h2=want_h and oh.cout % 4 == 0,
                                            overflow=self._overflow)
        elif po is not None:     # 8-child expansion: halo kernels addressed by the parent's map (no child map at all)
            y.f = ops.conv_k3_octet(x.f, level.parent.nbr, po, self.w[name + ".bias"], residual=residual, relu=relu, out=out)
        elif ph is not None:     # pre-split f16 features: LDG.128 -> HMMA.16816, no split arithmetic in the loop
            y.f, y.h = ops.conv_k3_h2(self._h(x), level.nbr, ph, self.w[name + ".bias"], residual=residual, relu=relu, out=out,
                                      out_h2=out_h, want_f32=want_f, want_h2=want_h and ph.cout % 4 == 0,
                                      overflow=self._overflow)
        elif sl is not None:     # wide output: 16-channel slices through the h2 kernel
            cout = 16 * len(sl)
            n = len(level)
            y.f = out if out is not None else (torch.empty((n, cout), dtype=torch.float32, device=self.device) if want_f else None)
            y.h = out_h if out_h is not None else (torch.empty((n, cout), dtype=torch.int32, device=self.device) if want_h else None)
            for j, (ps, bias) in enumerate(sl):
                cs = slice(16 * j, 16 * j + 16)
                ops.conv_k3_h2(self._h(x), level.nbr, ps, bias, residual=None if residual is None else residual[:, cs], relu=relu,
                               out=None if y.f is None else y.f[:, cs], out_h2=None if y.h is None else y.h[:, cs],
                               want_f32=False, overflow=self._overflow)
        elif pw is not None:
            y.f = ops.conv_k3_packed(x.f, level.nbr, pw, self.w[name + ".bias"], residual=residual, relu=relu, out=out)
        else:
            y.f = ops.conv_k3(x.f, level.nbr, self.w[name + ".kernel"], self.w[name + ".bias"], residual=residual,
                              relu=relu, out=out)
        if ev is not None:
            end.record()
            ev.append((start, end))
        if out is None:
            self._rec(name, y, level)
        return y

    def _k1(self, name, x, relu=False, residual=None, out=None, out_h=None):
        """k=1 layer on fp32 rows; ``out_h`` (a slice, or True): also write the h2 copy in the epilogue when the kernel can
        (-> fp32 out, h2 out or None when the caller has to split)."""
        w = self.w[name + ".kernel"]
        if out_h is not None and self._h2_on and ops.h2out_supported("k1", w.shape[0], w.shape[1]):
            return ops.conv_k1(x, w, self.w[name + ".bias"], residual=residual, relu=relu, out=out, out_h2=out_h,
                               overflow=self._overflow)
        return ops.conv_k1(x, w, self.w[name + ".bias"], residual=residual, relu=relu, out=out), None

    def _uses_h2(self, name, level):
        """the layer will run on an h2 kernel (tcgen05, gather or full-octet variant)."""
        if self._h2_on and name in self.packed_wide:
            return True
        if self._h2_on and name in self.packed_h2_slices and not (level.full_octets and name in self.packed_octet):
            return True
        if not (self._h2_on and name in self.packed_h2):
            return False
        ph = self.packed_h2[name]
        if level.full_octets and self.use_octet and ops.octet_h2_supported(ph.cin, ph.cout):
            return True
        if level.full_octets and name in self.packed_octet:
            return False
        return ph.gather

    # ---- one InceptionResNet block per C call (csrc/irn.cpp): same kernels, same order, no Python between the launches
    def _route(self, name, full_octets, aligned=True):
        """(route code, packed weights tensor, inverse weight scale) of one k=3 layer -- the decision ``_k3`` takes."""
        pwide = self.packed_wide.get(name) if self._h2_on else None
        if pwide is not None:
            return _lib.ROUTE_WIDE, pwide.packed, pwide.inv_scale
        ph = self.packed_h2.get(name) if self._h2_on else None
        if ph is not None and full_octets and self.use_octet and ops.octet_h2_supported(ph.cin, ph.cout):
            return _lib.ROUTE_H2_OCTET, ph.packed, ph.inv_scale
        if ph is not None and not ph.gather:
            ph = None
        po = self.packed_octet.get(name) if (full_octets and aligned) else None
        if po is not None:
            return _lib.ROUTE_TF32_OCTET, po.packed, 1.0
        if ph is not None:
            return _lib.ROUTE_H2_GATHER, ph.packed, ph.inv_scale
        pw = self.packed.get(name) if aligned else None
        if pw is not None:
            return _lib.ROUTE_TF32_GATHER, pw.packed, 1.0
        return _lib.ROUTE_FP32, self.w[name + ".kernel"], 1.0

    def _merged_first(self, prefix):
        """conv0_0 (k=3, c -> c/4) and conv1_0 (k=1, c -> c/4) of one InceptionResNet block as ONE k=3 convolution c -> c/2:
        output channels c/4.. carry conv1_0's weights at the centre offset.  Worth it where the 8-wide MMA tile of conv0_0 is
        half empty (c = 16): -> (packed h2 weights, bias [1, c/2]) or None."""
        m = self._merged.get(prefix)
        if m is None:
            w0, w1 = self.w[prefix + ".conv0_0.kernel"], self.w[prefix + ".conv1_0.kernel"]
            c, q = int(w0.shape[1]), int(w0.shape[2])
            m = False
            if self.merge_first and c == 16 and w1.shape == (c, q) and ops.octet_h2_supported(c, 2 * q):
                wm = torch.zeros((27, c, 2 * q), dtype=torch.float32, device=self.device)
                wm[:, :, :q] = w0
                wm[13, :, q:] = w1                                    # k = ix + 3 iy + 9 iz with i = 1: the centre offset
                bias = torch.cat([self.w[prefix + ".conv0_0.bias"], self.w[prefix + ".conv1_0.bias"]], dim=1).contiguous()
                pk = ops.PackedK3H2(wm.contiguous())
                m = (pk, bias) if pk.packed is not None else False
            self._merged[prefix] = m
        return m or None

    def _irn_plan(self, prefix, full_octets):
        key = (prefix, full_octets, self._h2_on)
        plan = self._irn_plans.get(key)
        if plan is None:
            args = _lib.IrnArgs()
            keep = []                                                 # tensors the struct points into
            for i, leaf in enumerate((".conv0_0", ".conv0_1", ".conv1_1")):
                route, w, inv = self._route(prefix + leaf, full_octets)
                args.route[i], args.inv_scale[i] = route, inv
                args.w3[i], args.b3[i] = w.data_ptr(), self.w[prefix + leaf + ".bias"].data_ptr()
                keep.append(w)
            h2_routes = (_lib.ROUTE_H2_GATHER, _lib.ROUTE_H2_OCTET, _lib.ROUTE_WIDE)
            merged = self._merged_first(prefix) if (self._h2_on and full_octets and self.use_octet and
                                                    all(r in h2_routes for r in args.route)) else None
            if merged is not None:
                pk, bias = merged
                args.reserved = 1                                     # PCGC_IRN_MERGED_FIRST
                args.route[0], args.inv_scale[0] = _lib.ROUTE_H2_OCTET, pk.inv_scale
                args.w3[0], args.b3[0] = pk.packed.data_ptr(), bias.data_ptr()
                keep += [pk.packed, bias]
            q = int(self.w[prefix + ".conv1_1.kernel"].shape[2])
            if (self.fuse_tail and self._h2_on and args.route[2] == _lib.ROUTE_H2_OCTET and
                    _lib.lib().pcgc_conv_k3_octet_h2_k1_supported(q, q, 2 * q)):
                args.reserved |= 2                                    # PCGC_IRN_FUSED_TAIL: conv1_1 + ReLU + conv1_2 in one kernel
                if (self.dual_second and (args.reserved & 1) and args.route[1] == _lib.ROUTE_H2_OCTET and
                        int(self.w[prefix + ".conv0_0.kernel"].shape[1]) == 16):
                    args.reserved |= 4                                # PCGC_IRN_DUAL_SECOND: conv0_1 and conv1_1 + conv1_2 in one kernel
            for i, leaf in enumerate((".conv1_0", ".conv1_2")):
                args.w1[i], args.b1[i] = self.w[prefix + leaf + ".kernel"].data_ptr(), self.w[prefix + leaf + ".bias"].data_ptr()
            routes = list(args.route)
            plan = self._irn_plans[key] = {
                "args": args, "keep": keep,
                "child_map": any(r in (_lib.ROUTE_H2_GATHER, _lib.ROUTE_TF32_GATHER, _lib.ROUTE_FP32, _lib.ROUTE_WIDE) for r in routes),
                "parent_map": any(r in (_lib.ROUTE_H2_OCTET, _lib.ROUTE_TF32_OCTET) for r in routes),
                "x_h2": routes[0] in (_lib.ROUTE_H2_GATHER, _lib.ROUTE_H2_OCTET, _lib.ROUTE_WIDE)}
        return plan

    def _irn_fused(self, prefix, x: _F, level) -> _F:
        n, c = x.f.shape
        plan = self._irn_plan(prefix, level.full_octets)
        a = plan["args"]
        out = _F(torch.empty((n, c), dtype=torch.float32, device=self.device),
                 torch.empty((n, c), dtype=torch.int32, device=self.device) if self._h2_on else None)
        ws = torch.empty(int(_lib.lib().pcgc_irn_ws_bytes(n, c)), dtype=torch.uint8, device=self.device)
        xh = self._h(x) if plan["x_h2"] else x.h
        a.n, a.c = n, c
        a.nbr = level.nbr.data_ptr() if plan["child_map"] else None
        a.parent_nbr = level.parent.nbr.data_ptr() if plan["parent_map"] else None
        a.x, a.x_ld = x.f.data_ptr(), x.f.stride(0)
        a.x_h2, a.x_h2_ld = (xh.data_ptr(), xh.stride(0)) if xh is not None else (None, 0)
        a.out, a.out_ld = out.f.data_ptr(), c
        a.out_h2, a.out_h2_ld = (out.h.data_ptr(), c) if out.h is not None else (None, 0)
        a.ws, a.ws_bytes = ws.data_ptr(), ws.numel()
        a.overflow = self._overflow.data_ptr()
        _lib.check(_lib.lib().pcgc_irn_fwd(ctypes.byref(a), ops._stream()), "pcgc_irn_fwd")
        self._rec(prefix, out, level)
        return out

    def _irn(self, prefix, x: _F, level) -> _F:
        """InceptionResNet (autoencoder.py:52-57) as 5 fused launches: the two branch outputs are
        written straight into the halves of the result with the residual added in the epilogue."""
        if (self.fuse_irn and self.record is None and x.f is not None and x.f.stride(0) % 4 == 0
                and not any(k.startswith(prefix) for k in self.probe)):      # per-layer recording / probing: layer by layer
            return self._irn_fused(prefix, x, level)
        c = x.f.shape[1]
        h = c // 2
        out = _F(torch.empty_like(x.f))
        a_h2 = self._uses_h2(prefix + ".conv0_1", level)             # conv0_0's output feeds an h2 kernel only
        a = self._k3(prefix + ".conv0_0", x, level, relu=True, want_f=not a_h2 or self.record is not None, want_h=a_h2)
        if a_h2:                                                     # this producer writes its half of the h2 copy too
            out.h = torch.empty((x.f.shape[0], c), dtype=torch.int32, device=self.device)
        self._k3(prefix + ".conv0_1", a, level, residual=x.f[:, :h], out=out.f[:, :h], out_h=None if out.h is None else out.h[:, :h])
        b = _F(*self._k1(prefix + ".conv1_0", x.f, relu=True, out_h=True if self._uses_h2(prefix + ".conv1_1", level) else None))
        cc = self._k3(prefix + ".conv1_1", b, level, relu=True)
        if out.h is None and self._h2_on:                            # the block output always feeds an h2 layer
            out.h = torch.empty((x.f.shape[0], c), dtype=torch.int32, device=self.device)
            first_half_done = False
        else:
            first_half_done = out.h is not None
        _, hh = self._k1(prefix + ".conv1_2", cc.f, residual=x.f[:, h:], out=out.f[:, h:],
                         out_h=None if out.h is None else out.h[:, h:])
        if out.h is not None:
            if hh is None and not first_half_done:
                ops.split_h2(out.f, out=out.h, overflow=self._overflow)
            else:
                if hh is None:
                    ops.split_h2(out.f[:, h:], out=out.h[:, h:], overflow=self._overflow)
                if not first_half_done:
                    ops.split_h2(out.f[:, :h], out=out.h[:, :h], overflow=self._overflow)
        self._rec(prefix, out, level)
        return out

    # ---------------------------------------------------------------- analysis / synthesis
    def _sorted_input(self, coords: torch.Tensor, dedupe: bool):
        """int32 [N,4] on the device -> (level-0 coordinate set in Morton order, device flag "has duplicates").
        Duplicates are rare: the first pass only raises the flag (read together with the symbol range, no extra
        synchronisation); ``dedupe`` drops them (``unique_consecutive`` synchronises for the output size)."""
        bits = self.coord_bits if (self.coord_bits is not None and coords.shape[1] == 3) else 0
        keys = ops.pack_keys_async(coords, 1, self._bad, hint_bits=bits)   # range flags: read with the pass's synchronising read
        keys, _ = ops.argsort_u64(keys, end_bit=3 * bits if bits else 64)
        if dedupe:
            keys = torch.unique_consecutive(keys)
        dup = (keys[1:] == keys[:-1]).any().to(torch.int32).reshape(1) if keys.numel() > 1 else self._overflow.new_zeros(1)
        return _Level(keys, 1), dup

    def analysis(self, level0: _Level):
        """autoencoder.py:138-147 -> (y [N3,8], level3, [N2, N1, N0])."""
        # coordinate pyramid first (integer work on keys only), so that the kernel maps can be
        # derived top-down from the coarsest set
        levels, down = [level0], []
        for i in range(3):
            pk, rows, off, parent_of = ops.stride_down(levels[-1].keys, keys_are_sorted=True, with_parent_of=True)
            levels[-1].parent_of, levels[-1].info = parent_of, ops.parent_info(levels[-1].keys, off)
            down.append((rows, off))
            levels.append(_Level(pk, levels[-1].stride * 2))
        for child, par in zip(levels[:-1], levels[1:]):
            child.parent = par
        w0 = self.w["encoder.conv0.kernel"]
        if (self.fuse_conv0 and self.record is None and "encoder.conv0" not in self.probe and w0.shape[1] == 1 and w0.shape[2] == 16
                and level0.info is not None):
            # the input features are the constant 1 (coder.py:131): conv0 = bias + the sum of the weights of the PRESENT neighbours,
            # read off the parent's kernel map -- the 27 x N0 kernel map of the finest level is never built
            want_h = self._h2_on and "encoder.down0" in self.packed_down_h2
            x = _F(*ops.conv_k3_ones_from_parent(level0.parent.nbr, level0.keys, level0.parent_of, level0.info, w0,
                                                 self.w["encoder.conv0.bias"], relu=True, want_f32=not want_h, want_h2=want_h,
                                                 overflow=self._overflow))
        else:
            x = _F(torch.ones((len(level0), 1), dtype=torch.float32, device=self.device))
            x = self._k3("encoder.conv0", x, level0, relu=True)
        level, sizes = level0, [len(level0)]
        for i in range(3):
            rows, off = down[i]
            wd = self.w[f"encoder.down{i}.kernel"]
            pd = self.packed_down_h2.get(f"encoder.down{i}") if self._h2_on else None
            if pd is not None:                                                           # tensor cores over the 8 child slots
                cmap = ops.child_map_k2(level.keys, level.parent_of, len(levels[i + 1]))
                x = _F(*ops.conv_k2s2_h2(self._h(x), cmap, pd, self.w[f"encoder.down{i}.bias"], relu=True,
                                         want_f32=True, want_h2=True, overflow=self._overflow))
            elif self._h2_on and ops.h2out_supported("down", wd.shape[1], wd.shape[2]):  # the IRN's h2 layers read this
                x = _F(*ops.conv_k2s2(x.f, level.keys, rows, off, wd, self.w[f"encoder.down{i}.bias"], relu=True, out_h2=True,
                                      overflow=self._overflow))
            else:
                x = _F(ops.conv_k2s2(x.f, level.keys, rows, off, wd, self.w[f"encoder.down{i}.bias"], relu=True))
            level = levels[i + 1]
            self._rec(f"encoder.down{i}", x, level)
            for j in range(3):
                x = self._irn(f"encoder.block{i}.{j}", x, level)
            sizes.append(len(level))
            x = self._k3(f"encoder.conv{i + 1}", x, level, relu=(i < 2),
                         want_h=i < 2 and self._h2_on and f"encoder.down{i + 1}" in self.packed_down_h2)
        return x.f, level, [sizes[2], sizes[1], sizes[0]]

    def synthesis(self, y: torch.Tensor, level: _Level, nums):
        """autoencoder.py:251-273 with training=False -> (final level, classifier logits per scale)."""
        x, cls_list = y, []
        for i in range(3):
            wu = self.w[f"decoder.up{i}.kernel"]
            pu = self.packed_up_h2.get(f"decoder.up{i}") if (self._h2_on and self._uses_h2(f"decoder.conv{i}", _FULL)) else None
            if pu is not None:                               # one dense tensor-core product; only the h2 copy is consumed
                x = _F(*ops.convT_k2s2_h2(ops.split_h2(x, overflow=self._overflow), pu, relu=True,
                                          want_f32=self.record is not None, want_h2=True, overflow=self._overflow))
            elif self._h2_on and self._uses_h2(f"decoder.conv{i}", _FULL) and ops.h2out_supported("up", wu.shape[1], wu.shape[2]):
                x = _F(*ops.convT_k2s2(x, wu, self.w[f"decoder.up{i}.bias"], relu=True, out_h2=True, overflow=self._overflow))
            else:
                x = _F(ops.convT_k2s2(x, wu, self.w[f"decoder.up{i}.bias"], relu=True))
            level = _Level(ops.upsample_keys(level.keys), level.stride // 2, parent=level)    # full octets
            self._rec(f"decoder.up{i}", x, level)
            x = self._k3(f"decoder.conv{i}", x, level, relu=True, want_h=True)
            for j in range(3):
                x = self._irn(f"decoder.block{i}.{j}", x, level)
            cls = self._k3(f"decoder.conv{i}_cls", x, level).f
            cls_list.append((cls, level))
            k = min(len(level), int(nums[i]))
            mask = ops.topk_mask(cls, k)                    # exactly k rows survive: no size read-back needed
            if i < 2:                                       # the pruned set parents the next up-sampling
                keys, x, nbr = ops.prune(mask, level.keys, x.f, nbr=level.nbr, n_kept_hint=k)
                level = _Level(keys, level.stride, nbr=nbr)
            else:
                keys, x = ops.prune(mask, level.keys, x.f, n_kept_hint=k)
                level = _Level(keys, level.stride)
        return level, x, cls_list

    # ---------------------------------------------------------------- codec
    @staticmethod
    def _canonical_order(coords3: torch.Tensor, bits: int = 20) -> torch.Tensor:
        """argsort of the reference's sort key b + x*S + y*S^2 + z*S^3 (data_utils.py:55-61,91-101):
        z most significant, x least -- any S > max gives the same order.  ``bits``: every coordinate is below 2^bits (the
        radix sort then runs over 3 * bits key bits)."""
        c = coords3.long()
        key = (c[:, 2] << (2 * bits)) | (c[:, 1] << bits) | c[:, 0]
        return ops.argsort_u64(key.contiguous(), end_bit=3 * bits)[1].long()

    def _h2_overflowed(self) -> bool:
        """True when an h2 producer met a value outside the f16 range during the pass that just finished (call
        after a synchronising read): the caller repeats the pass on the 3xTF32 kernels instead."""
        if not self._h2_on or not int(self._overflow.item()):
            return False
        self._overflow.zero_()
        self.h2_fallbacks += 1
        return True

    def _without_h2(self, fn, *args, **kw):
        on, self._h2_on = self._h2_on, False
        try:
            return fn(*args, **kw)
        finally:
            self._h2_on = on

    def scale(self, coords, factor: float) -> torch.Tensor:
        """``scale_sparse_tensor`` (data_utils.py:112-118; coder.py:149-152,166-167) on the device: int32 [N,3] -> the
        scaled, rounded and de-duplicated voxel set, int32 [M,3] (Morton order), on the device."""
        with torch.cuda.device(self.device), torch.no_grad(), ops.stream_scope():
            c = torch.as_tensor(coords, dtype=torch.int32).to(self.device, non_blocking=True)
            c = ops.scale_coords(c[:, -3:].contiguous(), factor)
            keys, _ = ops.argsort_u64(ops.pack_keys(torch.nn.functional.pad(c, (1, 0)), 1))
            return ops.unpack_keys(torch.unique_consecutive(keys), 1)[:, 1:].contiguous()

    def encode(self, coords) -> Stream:
        """coords: int32 [N,3] (or [N,4] with the batch column; batch 0 only) host array or tensor."""
        st = self._encode(coords)
        while not isinstance(st, Stream):                                # "dup": duplicates in the input; "h2": f16 range left;
            if st == "bits":                                             # "bits": a coordinate above the coord_bits promise
                self.coord_bits, self.coord_bits_fallbacks = None, self.coord_bits_fallbacks + 1
                st = self._encode(coords)
            else:
                st = self._encode(coords, dedupe=True) if st == "dup" else self._without_h2(self._encode, coords, self._dedupe_next)
        return st

    def decode(self, stream: Stream, rho: float = 1.0, to_host: bool = True):
        """-> decoded voxel coordinates int32 [N_out, 3] (host array, or device tensor if not to_host)."""
        out = self._decode(stream, rho, to_host)
        return out if out is not None else self._without_h2(self._decode, stream, rho, to_host)

    def _host_table(self, lo: int, hi: int) -> np.ndarray:
        """uint16 CDF table [C, L+1] of the symbol range: a pure function of the model and (lo, hi), built once per range
        on the host with the reference's own float32 operator sequence (entropy_host.py; entropy_model.py:151-171 runs on
        the CPU too, coder.py:44) and converted like torchac does (Appendix B.1), so that ``F`` is byte-identical to the
        reference's for identical symbols.  Cached: after the first frame of a symbol range no table work is left."""
        t = self._tables.get((lo, hi))
        if t is None:
            if len(self._tables) > 64:
                self._tables.clear()
            cdf = self.eb_host.cdf_float(lo, hi)
            lp = cdf.shape[1]
            scaled = np.round(cdf * np.float32(65536 - (lp - 1)))                       # float32 multiply, half-to-even
            t = self._tables[(lo, hi)] = (scaled.astype(np.int64) + np.arange(lp)).astype(np.uint16)
        return t

    def _sync(self):
        """wait for this frame's stream with a BLOCKING event: the host thread sleeps instead of spinning in
        cudaStreamSynchronize, so the `depth` frame threads of a rank do not each burn a core while the GPU works (eight ranks
        share the host).  Opt-in (PCGC_BLOCKING_SYNC=1): measured on B200 it LOSES -- 113.7 vs 118.2 Mpoints/s on one GPU with 16 cores,
        720.7 vs 855.7 on eight GPUs with 4 cores per rank (the wake-up through the interrupt path costs more than the spinning
        threads take from the enqueueing one)."""
        if not self._blocking_sync:
            if _SPIN_LOCK is not None:                            # PCGC_SYNC_LOCK=1: at most one thread of the process spins at a time
                # (measured: one host core less per rank -- 3.0 -> 2.0 busy with 4 frames in flight -- at unchanged throughput)
                with _SPIN_LOCK:
                    torch.cuda.current_stream().synchronize()
            else:
                torch.cuda.current_stream().synchronize()
            return
        ev = self._sync_event
        if ev is None:
            ev = self._sync_event = torch.cuda.Event(blocking=True)
        ev.record()
        ev.synchronize()

    def _staging(self, name, shape, dtype):
        """reusable pinned host buffer (grown geometrically): D2H copies are asynchronous and share one synchronise."""
        n = int(np.prod(shape))
        buf = self._pinned.get(name)
        if buf is None or buf.numel() < n or buf.dtype != dtype:
            buf = self._pinned[name] = torch.empty(max(n, 1) * 5 // 4 + 16, dtype=dtype, pin_memory=True)
        return buf[:n].view(shape)

    def _encode(self, coords, dedupe=False):
        with torch.cuda.device(self.device), torch.no_grad(), ops.stream_scope():
            return self._encode_pass(coords, dedupe)

    def _decode(self, stream: Stream, rho: float = 1.0, to_host: bool = True):
        with torch.cuda.device(self.device), torch.no_grad(), ops.stream_scope():
            return self._decode_pass(stream, rho, to_host)

    def _encode_pass(self, coords, dedupe=False):
        coords = torch.as_tensor(coords, dtype=torch.int32).to(self.device, non_blocking=True)   # async from pinned host memory
        if coords.dim() != 2 or coords.shape[1] not in (3, 4):
            raise ValueError("coordinates must be int32 [N,3] (x,y,z) or [N,4] (batch,x,y,z)")
        self._dedupe_next = dedupe
        level0, dup = self._sorted_input(coords, dedupe)
        y, level3, num_points = self.analysis(level0)
        c3 = ops.unpack_keys(level3.keys, 1)[:, 1:]                       # stride-8 coordinates / 8
        order = self._canonical_order(c3, max(1, self.coord_bits - 3) if self.coord_bits is not None else 20)   # c3 = coordinates / 8
        y, c3 = y[order].contiguous(), c3[order].contiguous()
        sym, mm = ops.eb_quantize_async(y)
        # one synchronising read for everything the host needs: symbol range + flags, symbols, coordinates
        flags_h = self._staging("flags", (5,), torch.int32)
        sym_h, c3_h = self._staging("sym", tuple(sym.shape), torch.int16), self._staging("c3", tuple(c3.shape), torch.int32)
        flags_h.copy_(torch.cat([mm, self._overflow, dup, self._bad]), non_blocking=True)
        sym_h.copy_(sym, non_blocking=True)
        c3_h.copy_(c3, non_blocking=True)
        self._sync()
        lo, hi, over, has_dup, bad = flags_h.tolist()
        if bad & 1:
            self._bad.zero_()
            self._overflow.zero_()
            raise ValueError("coordinates out of range: need 0 <= c <= %d, batch <= 126" % ((1 << 19) - 1))
        if bad & 2:                                                       # a coordinate above the coord_bits promise: the sorts were too
            self._bad.zero_()                                             # narrow; code the frame again at full key width
            self._overflow.zero_()
            return "bits"
        if has_dup:
            if over:
                self._overflow.zero_()                                    # the deduplicated re-run decides for itself
            return "dup"
        if over and self._h2_on:
            self._overflow.zero_()
            self.h2_fallbacks += 1
            return "h2"
        c3_np = c3_h.numpy().copy()
        y_keep = y.cpu().numpy() if self.keep_bottleneck else None
        c_job = self._side.submit(self.coords_coder.encode, c3_np) if self._side is not None else None   # overlaps the range coder
        f_bytes = ops.rc_encode_u16(self._host_table(lo, hi), sym_h.numpy())
        h_bytes = (np.array(y.shape, dtype=np.int32).tobytes() + np.array(1, dtype=np.int8).tobytes() +
                   np.array([lo], dtype=np.float32).tobytes() + np.array([hi], dtype=np.float32).tobytes())
        return Stream(F=f_bytes, H=h_bytes, num_points=np.array(num_points, dtype=np.int32).tobytes(),
                      coords=c3_np, stats={"N": num_points, "sym_range": (lo, hi), **({} if y_keep is None else {"y_F": y_keep})},
                      C=None if c_job is None else c_job.result())

    def _decode_pass(self, stream: Stream, rho: float = 1.0, to_host: bool = True):
        shape = np.frombuffer(stream.H[:8], dtype=np.int32)
        lo = int(np.frombuffer(stream.H[9:13], dtype=np.float32)[0])
        hi = int(np.frombuffer(stream.H[13:17], dtype=np.float32)[0])
        n3, ch = int(shape[0]), int(shape[1])
        sym_h = self._staging("sym_in", (n3, ch), torch.int16)
        if stream.C is not None:                                          # coordinates come out of the stream (coder.py:95-96):
            if self.coords_coder is None:                                 # decoded on the side thread while this one decodes F
                raise ValueError("the stream carries coded coordinates but this Codec has no coordinate coder")
            c_job = self._side.submit(self.coords_coder.decode, stream.C)
            ops.rc_decode_u16(self._host_table(lo, hi), stream.F, n3 * ch, out=sym_h.numpy().reshape(-1))
            coords_in = c_job.result()
            if coords_in.shape[0] != n3:
                raise ValueError(f"coordinate stream holds {coords_in.shape[0]} points, header says {n3}")
        else:
            coords_in = stream.coords
        c3_h = self._staging("c3_in", (n3, 3), torch.int32)
        c3_h.copy_(torch.as_tensor(coords_in, dtype=torch.int32))
        c3 = c3_h.to(self.device, non_blocking=True)
        cmax = int(c3_h.max()) if n3 else 0                               # host data: the exact key width of the two small sorts
        cbits = max(1, cmax.bit_length()) if 0 <= cmax < (1 << 19) and (n3 == 0 or int(c3_h.min()) >= 0) else 0
        c3 = c3[self._canonical_order(c3, cbits or 20)]                  # coder.py:97-99 (runs while the host decodes the symbols)
        keys = ops.pack_keys_async(c3, 1, self._bad)                      # = the stride-8 keys of 8 * c3 (keys hold coordinate / stride)
        if stream.C is None:
            ops.rc_decode_u16(self._host_table(lo, hi), stream.F, n3 * ch, out=sym_h.numpy().reshape(-1))
        y = sym_h.to(self.device, non_blocking=True).float() + float(lo)
        keys, order = ops.argsort_u64(keys, end_bit=3 * cbits if cbits else 64)   # Morton order for the synthesis network
        level3 = _Level(keys, 8)
        nums = np.frombuffer(stream.num_points, dtype=np.int32).tolist()
        nums[-1] = int(rho * nums[-1])                                   # coder.py:107
        level0, _, _ = self.synthesis(y[order.long()].contiguous(), level3, nums)
        out = ops.unpack_keys(level0.keys, 1)[:, 1:]
        flag_h = self._staging("flag_out", (2,), torch.int32)
        flag_h.copy_(torch.cat([self._overflow, self._bad]), non_blocking=True)
        if not to_host:
            self._sync()
        else:
            out = out.contiguous()
            host = self._staging("out", tuple(out.shape), torch.int32)   # D2H through a reusable pinned buffer
            host.copy_(out, non_blocking=True)
            self._sync()
        if int(flag_h[1]):
            self._bad.zero_()
            self._overflow.zero_()
            raise ValueError("bottleneck coordinates out of range")
        if not to_host:
            if self._h2_on and int(flag_h[0]):
                self._overflow.zero_()
                self.h2_fallbacks += 1
                return None
            return out
        if self._h2_on and int(flag_h[0]):
            self._overflow.zero_()
            self.h2_fallbacks += 1
            return None
        return host.numpy()                                              # view of the pinned buffer: valid until the next decode()
