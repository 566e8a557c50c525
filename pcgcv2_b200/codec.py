"""Encode / decode pipeline of the PCGCv2 geometry codec on libpcgc (the path bench.py measures).

``Codec`` mirrors the reference's ``Coder`` (coder.py:73-112): ``encode`` runs the analysis
network, sorts the bottleneck into the canonical symbol order, quantises and range-codes the
features; ``decode`` inverts it and runs the synthesis network with top-k pruning.  The byte
layouts of ``_F.bin`` / ``_H.bin`` / ``_num_points.bin`` are the reference's (coder.py:49-55,
85-87).  The stride-8 coordinate side channel is returned as an int32
array: the reference pipes it through the external ``tmc3`` binary (coder.py:23-36), which is
outside the hot path.

Unlike the per-operator shim, the pipeline keeps every coordinate set in ascending Morton-key
order (one radix sort of the input; stride-2 parents, 8-child expansion and stable pruning all
preserve it), fuses bias / ReLU / concat / residual into the convolution epilogues and never
leaves the device between layers.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import torch

from . import ops


@dataclass
class Stream:
    """the four pieces the reference writes to disk (coder.py:81-91)."""
    F: bytes
    H: bytes
    num_points: bytes
    coords: np.ndarray                  # int32 [N3, 3], stride-8 coordinates / 8, canonical order (-> tmc3 in the reference)
    stats: dict = field(default_factory=dict)

    def bits(self, coords_bits: int = 0) -> int:
        return 8 * (len(self.F) + len(self.H) + len(self.num_points)) + coords_bits


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


class Codec:
    def __init__(self, state_dict, device="cuda", use_tensor_cores=True, use_octet_kernels=True):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("pcgcv2_b200.Codec runs on a CUDA device only (there is no CPU path)")
        self.w = {k: v.detach().float().contiguous().to(self.device) for k, v in state_dict.items()
                  if k.startswith(("encoder.", "decoder."))}
        g = lambda name: [state_dict[f"entropy_bottleneck.{name}.{i}"] for i in range(4)]
        self.eb_params = ops.pack_eb_params(g("_matrices"), g("_biases"), g("_factors"), self.device)
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
        self._pinned_out = None         # reusable pinned host buffer for the decoded coordinates
        self.record = None              # set to a dict to capture per-layer activations (parity tests)
        self.probe = {}                 # layer name -> list of (start, end) CUDA event pairs (bench.py roofline)

    # ---------------------------------------------------------------- layers
    def _rec(self, name, t, level=None):
        if self.record is not None:
            self.record[name] = (t.clone(), None if level is None else level.keys.clone(),
                                 None if level is None else level.stride)

    def _k3(self, name, x, level, relu=False, residual=None, out=None):
        aligned = x.stride(0) % 4 == 0
        po = self.packed_octet.get(name) if (level.full_octets and aligned) else None
        pw = self.packed.get(name) if aligned else None
        ev = self.probe.get(name)
        if ev is not None:
            nbr = level.parent.nbr if po is not None else level.nbr   # keep the (cached) map build outside the probe
            start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            start.record()
        if po is not None:       # 8-child expansion: halo kernels addressed by the parent's map (no child map at all)
            y = ops.conv_k3_octet(x, level.parent.nbr, po, self.w[name + ".bias"], residual=residual, relu=relu, out=out)
        elif pw is not None:
            y = ops.conv_k3_packed(x, level.nbr, pw, self.w[name + ".bias"], residual=residual, relu=relu, out=out)
        else:
            y = ops.conv_k3(x, level.nbr, self.w[name + ".kernel"], self.w[name + ".bias"], residual=residual,
                            relu=relu, out=out)
        if ev is not None:
            end.record()
            ev.append((start, end))
        if out is None:
            self._rec(name, y, level)
        return y

    def _k1(self, name, x, relu=False, residual=None, out=None):
        return ops.conv_k1(x, self.w[name + ".kernel"], self.w[name + ".bias"], residual=residual, relu=relu, out=out)

    def _irn(self, prefix, x, level):
        """InceptionResNet (autoencoder.py:52-57) as 5 fused launches: the two branch outputs are
        written straight into the halves of the result with the residual added in the epilogue."""
        c = x.shape[1]
        h = c // 2
        out = torch.empty_like(x)
        a = self._k3(prefix + ".conv0_0", x, level, relu=True)
        self._k3(prefix + ".conv0_1", a, level, residual=x[:, :h], out=out[:, :h])
        b = self._k1(prefix + ".conv1_0", x, relu=True)
        cc = self._k3(prefix + ".conv1_1", b, level, relu=True)
        self._k1(prefix + ".conv1_2", cc, residual=x[:, h:], out=out[:, h:])
        self._rec(prefix, out, level)
        return out

    # ---------------------------------------------------------------- analysis / synthesis
    def _sorted_input(self, coords: torch.Tensor):
        """int32 [N,4] on the device -> level-0 coordinate set in Morton order (duplicates dropped)."""
        keys = ops.pack_keys(coords, 1)
        keys, _ = ops.argsort_u64(keys)
        if keys.numel() > 1 and bool((keys[1:] == keys[:-1]).any()):
            keys = torch.unique_consecutive(keys)
        return _Level(keys, 1)

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
        x = torch.ones((len(level0), 1), dtype=torch.float32, device=self.device)
        x = self._k3("encoder.conv0", x, level0, relu=True)
        level, sizes = level0, [len(level0)]
        for i in range(3):
            rows, off = down[i]
            x = ops.conv_k2s2(x, level.keys, rows, off, self.w[f"encoder.down{i}.kernel"],
                              self.w[f"encoder.down{i}.bias"], relu=True)
            level = levels[i + 1]
            self._rec(f"encoder.down{i}", x, level)
            for j in range(3):
                x = self._irn(f"encoder.block{i}.{j}", x, level)
            sizes.append(len(level))
            x = self._k3(f"encoder.conv{i + 1}", x, level, relu=(i < 2))
        return x, level, [sizes[2], sizes[1], sizes[0]]

    def synthesis(self, y: torch.Tensor, level: _Level, nums):
        """autoencoder.py:251-273 with training=False -> (final level, classifier logits per scale)."""
        x, cls_list = y, []
        for i in range(3):
            x = ops.convT_k2s2(x, self.w[f"decoder.up{i}.kernel"], self.w[f"decoder.up{i}.bias"], relu=True)
            level = _Level(ops.upsample_keys(level.keys), level.stride // 2, parent=level)    # full octets
            self._rec(f"decoder.up{i}", x, level)
            x = self._k3(f"decoder.conv{i}", x, level, relu=True)
            for j in range(3):
                x = self._irn(f"decoder.block{i}.{j}", x, level)
            cls = self._k3(f"decoder.conv{i}_cls", x, level)
            cls_list.append((cls, level))
            k = min(len(level), int(nums[i]))
            mask = ops.topk_mask(cls, k)                    # exactly k rows survive: no size read-back needed
            if i < 2:                                       # the pruned set parents the next up-sampling
                keys, x, nbr = ops.prune(mask, level.keys, x, nbr=level.nbr, n_kept_hint=k)
                level = _Level(keys, level.stride, nbr=nbr)
            else:
                keys, x = ops.prune(mask, level.keys, x, n_kept_hint=k)
                level = _Level(keys, level.stride)
        return level, x, cls_list

    # ---------------------------------------------------------------- codec
    @staticmethod
    def _canonical_order(coords3: torch.Tensor) -> torch.Tensor:
        """argsort of the reference's sort key b + x*S + y*S^2 + z*S^3 (data_utils.py:55-61,91-101):
        z most significant, x least -- any S > max gives the same order."""
        c = coords3.long()
        key = (c[:, 2] << 40) | (c[:, 1] << 20) | c[:, 0]
        return ops.argsort_u64(key.contiguous(), end_bit=60)[1].long()

    @torch.no_grad()
    def encode(self, coords) -> Stream:
        """coords: int32 [N,3] (or [N,4] with the batch column; batch 0 only) host array or tensor."""
        coords = torch.as_tensor(coords, dtype=torch.int32)
        if coords.shape[1] == 3:
            coords = torch.cat([torch.zeros((len(coords), 1), dtype=torch.int32, device=coords.device), coords], dim=1)
        level0 = self._sorted_input(coords.to(self.device, non_blocking=True))
        y, level3, num_points = self.analysis(level0)
        c3 = ops.unpack_keys(level3.keys, 1)[:, 1:]                       # stride-8 coordinates / 8
        order = self._canonical_order(c3)
        y, c3 = y[order].contiguous(), c3[order]
        sym, lo, hi = ops.eb_quantize(y)
        _, table = ops.eb_cdf_table(self.eb_params, lo, hi)
        f_bytes = ops.rc_encode_u16(table.cpu().numpy(), sym.cpu().numpy())
        h_bytes = (np.array(y.shape, dtype=np.int32).tobytes() + np.array(1, dtype=np.int8).tobytes() +
                   np.array([lo], dtype=np.float32).tobytes() + np.array([hi], dtype=np.float32).tobytes())
        return Stream(F=f_bytes, H=h_bytes, num_points=np.array(num_points, dtype=np.int32).tobytes(),
                      coords=c3.cpu().numpy(), stats={"N": num_points, "sym_range": (lo, hi)})

    @torch.no_grad()
    def decode(self, stream: Stream, rho: float = 1.0, to_host: bool = True):
        """-> decoded voxel coordinates int32 [N_out, 3] (host array, or device tensor if not to_host)."""
        shape = np.frombuffer(stream.H[:8], dtype=np.int32)
        lo = int(np.frombuffer(stream.H[9:13], dtype=np.float32)[0])
        hi = int(np.frombuffer(stream.H[13:17], dtype=np.float32)[0])
        n3, ch = int(shape[0]), int(shape[1])
        c3 = torch.as_tensor(stream.coords, dtype=torch.int32).to(self.device, non_blocking=True)
        _, table = ops.eb_cdf_table(self.eb_params, lo, hi)
        sym = ops.rc_decode_u16(table.cpu().numpy(), stream.F, n3 * ch).reshape(n3, ch)
        c3 = c3[self._canonical_order(c3)]                               # coder.py:97-99
        y = torch.from_numpy(sym.astype(np.float32)).to(self.device) + float(lo)
        keys = ops.pack_keys(torch.cat([torch.zeros((n3, 1), dtype=torch.int32, device=self.device), c3 * 8], dim=1), 8)
        keys, order = ops.argsort_u64(keys)                              # Morton order for the synthesis network
        level3 = _Level(keys, 8)
        nums = np.frombuffer(stream.num_points, dtype=np.int32).tolist()
        nums[-1] = int(rho * nums[-1])                                   # coder.py:107
        level0, _, _ = self.synthesis(y[order.long()].contiguous(), level3, nums)
        out = ops.unpack_keys(level0.keys, 1)[:, 1:]
        if not to_host:
            return out
        out = out.contiguous()
        n = out.shape[0]                                                 # D2H through a reusable pinned buffer
        if self._pinned_out is None or self._pinned_out.shape[0] < n:
            self._pinned_out = torch.empty((max(n, 1) * 5 // 4, 3), dtype=torch.int32, pin_memory=True)
        host = self._pinned_out[:n]
        host.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return host.numpy()                                              # view of the pinned buffer: valid until the next decode()
