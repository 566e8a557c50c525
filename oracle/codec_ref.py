"""CPU restatement of the PCGCv2 encode/decode flow -- TEST INFRASTRUCTURE.

Topology follows ``autoencoder.py:52-57`` (InceptionResNet), ``:138-147``
(Encoder.forward), ``:239-273`` (Decoder.prune_voxel / forward); the bitstream
flow follows ``coder.py:46-70,80-112``.  Operators come from ``sparse_ref``
(PARITY UNPINNED, see there); weights are a reference ``state_dict``.

The G-PCC side channel (``coder.py:23-36`` -> ``tmc3``) is out of the hot
path's scope: the stride-8 coordinates are handed to the
decoder directly, in the canonical ``sort_spare_tensor`` order the reference
imposes on both sides (``coder.py:84,97-99``).
"""
from __future__ import annotations

import numpy as np
import torch

from . import entropy_ref, rangecoder_ref, sparse_ref as S


class _Maps:
    """kernel-map cache per coordinate set (Appendix A.8)."""

    def __init__(self):
        self._m = {}

    def k3(self, coords, stride):
        key = (id(coords), stride)
        if key not in self._m:
            self._m[key] = (coords, S.kernel_map_k3(coords, stride))
        return self._m[key][1]


def _conv(sd, name, feats, coords, stride, maps, record=None):
    w = sd[name + ".kernel"].float()
    b = sd[name + ".bias"].float()
    if w.dim() == 2:
        out = S.conv_k1(feats, w, b)
    else:
        assert w.shape[0] == 27
        out = S.conv_from_map(feats, maps.k3(coords, stride), w, b)
    if record is not None:
        record[name] = out
        record[name + ".C"] = coords
    return out


def _irn(sd, prefix, x, coords, stride, maps, record=None):
    """``autoencoder.py:52-57``."""
    relu = torch.relu
    a = relu(_conv(sd, prefix + ".conv0_0", x, coords, stride, maps, record))
    out0 = _conv(sd, prefix + ".conv0_1", a, coords, stride, maps, record)
    b = relu(_conv(sd, prefix + ".conv1_0", x, coords, stride, maps, record))
    c = relu(_conv(sd, prefix + ".conv1_1", b, coords, stride, maps, record))
    out1 = _conv(sd, prefix + ".conv1_2", c, coords, stride, maps, record)
    out = torch.cat([out0, out1], dim=1) + x
    if record is not None:
        record[prefix] = out
        record[prefix + ".C"] = coords
    return out


def encoder_forward(sd, coords: np.ndarray, feats: torch.Tensor, record=None):
    """``autoencoder.py:138-147``.  Returns [(F, C, stride)] for out2, out1, out0."""
    relu = torch.relu
    maps = _Maps()
    stride = 1
    x = relu(_conv(sd, "encoder.conv0", feats, coords, stride, maps, record))
    outs = []
    for lvl in range(3):
        w = sd[f"encoder.down{lvl}.kernel"].float()
        b = sd[f"encoder.down{lvl}.bias"].float()
        x, coords = S.conv_k2s2(x, coords, stride, w, b)
        if record is not None:
            record[f"encoder.down{lvl}"] = x
            record[f"encoder.down{lvl}.C"] = coords
        stride *= 2
        x = relu(x)
        for i in range(3):
            x = _irn(sd, f"encoder.block{lvl}.{i}", x, coords, stride, maps, record)
        outs.append((x, coords, stride))
        if lvl < 2:
            x = relu(_conv(sd, f"encoder.conv{lvl + 1}", x, coords, stride, maps, record))
    y = _conv(sd, "encoder.conv3", outs[2][0], outs[2][1], outs[2][2], maps, record)
    return [(y, outs[2][1], 8), outs[1], outs[0]]


def decoder_forward(sd, coords: np.ndarray, feats: torch.Tensor, nums, record=None, ground_truths=None):
    """``autoencoder.py:251-273`` with ``training=False`` (top-k pruning only,
    ``:239-249``).  ``nums`` = [k0, k1, k2] voxels kept per scale."""
    relu = torch.relu
    stride = 8
    x = feats
    cls_list = []
    for lvl in range(3):
        w = sd[f"decoder.up{lvl}.kernel"].float()
        b = sd[f"decoder.up{lvl}.bias"].float()
        x, coords = S.convT_k2s2(x, coords, stride, w, b)
        stride //= 2
        if record is not None:
            record[f"decoder.up{lvl}"] = x
            record[f"decoder.up{lvl}.C"] = coords
        maps = _Maps()
        x = relu(x)
        x = relu(_conv(sd, f"decoder.conv{lvl}", x, coords, stride, maps, record))
        for i in range(3):
            x = _irn(sd, f"decoder.block{lvl}.{i}", x, coords, stride, maps, record)
        cls = _conv(sd, f"decoder.conv{lvl}_cls", x, coords, stride, maps, record)
        cls_list.append((cls, coords))
        mask = S.topk_mask(cls.detach(), nums[lvl])
        if ground_truths is not None:                      # training=True: autoencoder.py:241-244
            mask = mask | S.isin(coords, ground_truths[lvl])
        x, coords = S.prune(x, coords, mask)
        if record is not None:
            record[f"decoder.prune{lvl}.C"] = coords
    return cls_list, x, coords


def train_forward(sd, coords: np.ndarray, quantize="symbols"):
    """``PCCModel.forward(x, training=True)`` (``pcc_model.py:26-45``) + the losses of ``trainer.py:127-134``
    (``loss.py:7-20``) with the straight-through round quantiser (deterministic); differentiable torch ops, so
    autograd through it is the reference for the backward kernels."""
    coords, _ = S.unique_coords(np.asarray(coords, dtype=np.int32))
    feats = torch.ones((len(coords), 1), dtype=torch.float32)
    y_list = encoder_forward(sd, coords, feats)
    yF, yC, _ = y_list[0]
    gts = [y_list[1][1], y_list[2][1], coords]
    nums = [len(g) for g in gts]
    params = {"matrices": [sd[f"entropy_bottleneck._matrices.{i}"] for i in range(4)],
              "biases": [sd[f"entropy_bottleneck._biases.{i}"] for i in range(4)],
              "factors": [sd[f"entropy_bottleneck._factors.{i}"] for i in range(4)]}
    yq = yF + (yF.round() - yF).detach() if quantize == "symbols" else yF
    lik = torch.clamp(entropy_ref.likelihood(params, yq), min=entropy_ref.LIKELIHOOD_BOUND)
    cls_list, _, _ = decoder_forward(sd, yC, yq, nums, ground_truths=gts)
    bce = 0
    for (cls, cls_c), gt in zip(cls_list, gts):
        target = torch.from_numpy(S.isin(cls_c, gt).astype(np.float32))
        bce = bce + torch.nn.functional.binary_cross_entropy_with_logits(cls.squeeze(1), target) / np.log(2.0)
    bpp = -torch.log2(lik).sum() / float(len(coords))
    return bce + bpp, bce, bpp


def encode(sd, coords: np.ndarray, record=None):
    """``coder.py:80-91``: encoder -> canonical sort of the bottleneck ->
    num_points + feature bitstream (``coder.py:46-57``) + stride-8 coordinates
    (which the reference hands to tmc3, ``coder.py:89``)."""
    coords, _ = S.unique_coords(np.asarray(coords, dtype=np.int32))
    feats = torch.ones((len(coords), 1), dtype=torch.float32)
    y_list = encoder_forward(sd, coords, feats, record)
    yF, yC, ystride = y_list[0]
    order = np.argsort(S.sort_key(yC), kind="stable")
    yF, yC = yF[torch.from_numpy(order)], yC[order]
    num_points = [len(y_list[1][1]), len(y_list[2][1]), len(coords)]
    params = entropy_ref.params_from_state_dict(sd)
    sym, min_v, max_v = entropy_ref.quantize_symbols(yF)
    cdf = entropy_ref.cdf_table(params, min_v, max_v, yF.shape[1])
    table = rangecoder_ref.cdf_float_to_u16(cdf.numpy())
    rows = np.tile(np.arange(yF.shape[1], dtype=np.int32), yF.shape[0])
    f_bytes = rangecoder_ref.encode_u16(table, rows, sym.numpy().reshape(-1))
    h_bytes = (np.array(yF.shape, dtype=np.int32).tobytes() + np.array(1, dtype=np.int8).tobytes()
               + np.array([min_v], dtype=np.float32).tobytes() + np.array([max_v], dtype=np.float32).tobytes())
    np_bytes = np.array(num_points, dtype=np.int32).tobytes()
    return {"F": f_bytes, "H": h_bytes, "num_points": np_bytes,
            "C_coords": (yC // ystride)[:, 1:].astype(np.int32), "y_F": yF, "y_C": yC,
            "ideal_bits": entropy_ref.ideal_bits(params, yF.round())}


def decode(sd, stream, rho=1.0, record=None):
    """``coder.py:93-112``."""
    yC3 = np.asarray(stream["C_coords"], dtype=np.int32)
    yC = np.concatenate([np.zeros((len(yC3), 1), dtype=np.int32), yC3], axis=1)
    order = np.argsort(S.sort_key(yC), kind="stable")
    yC = yC[order]
    shape = np.frombuffer(stream["H"][:8], dtype=np.int32)
    min_v = float(np.frombuffer(stream["H"][9:13], dtype=np.float32)[0])
    max_v = float(np.frombuffer(stream["H"][13:17], dtype=np.float32)[0])
    params = entropy_ref.params_from_state_dict(sd)
    cdf = entropy_ref.cdf_table(params, min_v, max_v, int(shape[1]))
    table = rangecoder_ref.cdf_float_to_u16(cdf.numpy())
    rows = np.tile(np.arange(shape[1], dtype=np.int32), int(shape[0]))
    sym = rangecoder_ref.decode_u16(table, rows, stream["F"]).reshape(int(shape[0]), int(shape[1]))
    yF = torch.from_numpy(sym.astype(np.float32)) + min_v
    num_points = np.frombuffer(stream["num_points"], dtype=np.int32).tolist()
    num_points[-1] = int(rho * num_points[-1])
    cls_list, x, coords = decoder_forward(sd, yC * 8, yF, num_points, record)
    return coords, cls_list


def scale_coords(coords3: np.ndarray, factor: float) -> np.ndarray:
    """``scale_sparse_tensor`` (``data_utils.py:112-118``; ``coder.py:149-152,166-167``): float32 multiply,
    ``torch.round`` (half to even), int cast, then the coordinate-map insertion of ``ME.SparseTensor`` collapses
    duplicates (first seen kept, Appendix A.2).  int [N,3] -> int32 [M,3]."""
    c = (torch.from_numpy(np.asarray(coords3, dtype=np.int32)) * factor).round().int().numpy()
    c4 = np.concatenate([np.zeros((len(c), 1), dtype=np.int32), c], axis=1)
    return np.ascontiguousarray(S.unique_coords(c4)[0][:, 1:])


def stream_bits(stream) -> int:
    return 8 * (len(stream["F"]) + len(stream["H"]) + len(stream["num_points"]))
