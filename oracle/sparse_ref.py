"""CPU restatement of the MinkowskiEngine operator semantics used by PCGCv2.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  PARITY UNPINNED: there is no
MinkowskiEngine source or golden vector under /root/reference; the conventions
below follow SURVEY.md Appendix A and are pinned behaviourally by the shipped
checkpoints (Appendix E.7) and by the brute-force dictionary tests in
``tests/test_oracle.py``.

Algorithm shape follows what ME's CPU backend does for the call sites in
``autoencoder.py``: build a kernel map (in_row, out_row) per kernel offset,
then per offset ``out[out_rows] += in[in_rows] @ W[k]`` (gather -> mm ->
index_add), finally ``+ bias``.  float32 throughout, ascending-k summation
order (Appendix A.11).

Coordinates are int32 ``[N, 4]`` = (batch, x, y, z) (Appendix A.1).
"""
from __future__ import annotations

import numpy as np
import torch

_OFF = 1 << 19          # coordinate bias so small negatives stay orderable
_BITS = 20


def coord_key(coords: np.ndarray) -> np.ndarray:
    """Injective int64 key of (b, x, y, z) rows (any fixed order works: the
    oracle only needs equality / lookup, never ME's internal hash order)."""
    c = np.asarray(coords, dtype=np.int64)
    assert c.ndim == 2 and c.shape[1] == 4
    if c.size:
        assert c[:, 1:].min() >= -_OFF and c[:, 1:].max() < _OFF, "coordinate out of oracle range"
    return (((c[:, 0] << _BITS | (c[:, 3] + _OFF)) << _BITS | (c[:, 2] + _OFF)) << _BITS) | (c[:, 1] + _OFF)


def unique_coords(coords: np.ndarray):
    """``ME.SparseTensor(features, coordinates)`` coordinate-map insertion
    (call sites ``data_utils.py:96,108,116``, ``coder.py:102``): duplicates
    collapse to one row, first-seen representative, input order preserved for
    the survivors (Appendix A.2).  Returns (unique coords, index of the kept
    input row per output row)."""
    key = coord_key(coords)
    _, first = np.unique(key, return_index=True)
    first = np.sort(first)
    return np.ascontiguousarray(coords[first]), first


class _Lookup:
    """sorted-key lookup table: coordinate -> row index (or -1)."""

    def __init__(self, coords: np.ndarray):
        key = coord_key(coords)
        self.order = np.argsort(key, kind="stable")
        self.sorted = key[self.order]
        if len(self.sorted) > 1:
            assert (np.diff(self.sorted) > 0).all(), "coordinate map holds duplicates"

    def find(self, coords: np.ndarray) -> np.ndarray:
        c = np.asarray(coords, dtype=np.int64)
        inside = ((c[:, 1:] >= -_OFF) & (c[:, 1:] < _OFF)).all(axis=1)   # queries outside the key range miss
        key = coord_key(np.where(inside[:, None], c, 0))
        if len(self.sorted) == 0:
            return np.full(len(key), -1, dtype=np.int64)
        pos = np.searchsorted(self.sorted, key)
        pos_c = np.minimum(pos, len(self.sorted) - 1)
        hit = (self.sorted[pos_c] == key) & inside
        return np.where(hit, self.order[pos_c], -1).astype(np.int64)


def kernel_offsets(kernel_size: int, tensor_stride: int) -> np.ndarray:
    """Offsets of a HYPER_CUBE kernel region, x fastest (Appendix A.3/A.4):
    odd kernels are centred, even kernels span {0, +stride}; kernel index
    k = ix + K*iy + K*K*iz."""
    K = kernel_size
    rng = (np.arange(K) - (K - 1) // 2) if K % 2 == 1 else np.arange(K)
    offs = [(ix, iy, iz) for iz in rng for iy in rng for ix in rng]
    return np.asarray(offs, dtype=np.int64) * tensor_stride


def kernel_map_k3(coords: np.ndarray, tensor_stride: int) -> np.ndarray:
    """Kernel map of a k=3, stride-1 convolution (Appendix A.3): ``nbr[u, k]``
    is the row of coordinate ``coords[u] + offset_k`` or -1.  Output map ==
    input map.  Restates ``coordinate_manager.kernel_map`` as used by every k=3
    ``ME.MinkowskiConvolution`` in ``autoencoder.py:13,20,35,71,90,...``."""
    lut = _Lookup(coords)
    offs = kernel_offsets(3, tensor_stride)
    n = len(coords)
    nbr = np.empty((n, 27), dtype=np.int64)
    q = np.asarray(coords, dtype=np.int64)
    for k in range(27):
        qq = q.copy()
        qq[:, 1:] += offs[k]
        nbr[:, k] = lut.find(qq)
    return nbr


def conv_from_map(feats: torch.Tensor, nbr: np.ndarray, weight: torch.Tensor,
                  bias: torch.Tensor | None, n_out: int | None = None) -> torch.Tensor:
    """Per-offset gather -> mm -> index_add, ascending k (Appendix A.11).
    ``nbr[u, k]`` = input row feeding output row u through W[k] (or -1)."""
    n_out = nbr.shape[0] if n_out is None else n_out
    K = weight.shape[0]
    out = torch.zeros((n_out, weight.shape[2]), dtype=torch.float32)
    for k in range(K):
        out_rows = np.nonzero(nbr[:, k] >= 0)[0]
        if len(out_rows) == 0:
            continue
        in_rows = torch.from_numpy(nbr[out_rows, k])
        buf = feats.index_select(0, in_rows) @ weight[k]
        out.index_add_(0, torch.from_numpy(out_rows), buf)
    if bias is not None:
        out += bias.reshape(1, -1)
    return out


def conv_k3(feats, coords, tensor_stride, weight, bias, nbr=None):
    """``ME.MinkowskiConvolution(kernel_size=3, stride=1)`` forward (§8 a3)."""
    if nbr is None:
        nbr = kernel_map_k3(coords, tensor_stride)
    return conv_from_map(feats, nbr, weight, bias)


def conv_k1(feats, weight, bias):
    """``kernel_size=1`` convolution = ``F.mm(kernel)`` (Appendix A.7)."""
    out = feats @ weight
    if bias is not None:
        out = out + bias.reshape(1, -1)
    return out


def stride_down(coords: np.ndarray, tensor_stride: int):
    """Coordinate map of a k=2, s=2 convolution output (Appendix A.5):
    ``floor(c / 2s) * 2s`` per spatial axis, unique.  Returns (out coords in
    first-seen order, parent row per input row, kernel index per input row)."""
    s2 = 2 * tensor_stride
    c = np.asarray(coords, dtype=np.int64)
    par = c.copy()
    par[:, 1:] = np.floor_divide(c[:, 1:], s2) * s2
    out_coords, _ = unique_coords(par.astype(np.int32))
    parent = _Lookup(out_coords).find(par)
    d = (c[:, 1:] - par[:, 1:]) // tensor_stride          # in {0,1}^3
    kidx = d[:, 0] + 2 * d[:, 1] + 4 * d[:, 2]
    return out_coords, parent, kidx


def conv_k2s2(feats, coords, tensor_stride, weight, bias):
    """``ME.MinkowskiConvolution(kernel_size=2, stride=2)`` forward (§8 a5,
    ``autoencoder.py:78,97,116``): every input row is one (in,out) pair with
    kernel index k = child position inside its 2x2x2 parent cell."""
    out_coords, parent, kidx = stride_down(coords, tensor_stride)
    out = torch.zeros((len(out_coords), weight.shape[2]), dtype=torch.float32)
    for k in range(8):
        rows = np.nonzero(kidx == k)[0]
        if len(rows) == 0:
            continue
        buf = feats.index_select(0, torch.from_numpy(rows)) @ weight[k]
        out.index_add_(0, torch.from_numpy(parent[rows]), buf)
    if bias is not None:
        out += bias.reshape(1, -1)
    return out, out_coords


def convT_k2s2(feats, coords, tensor_stride, weight, bias):
    """``ME.MinkowskiGenerativeConvolutionTranspose(kernel_size=2, stride=2)``
    forward (§8 a6, ``autoencoder.py:155,182,209``; Appendix A.6): every input
    row spawns its 8 children ``c + {0, s/2}^3`` at stride s/2, row ``8*i + k``,
    ``out[8i+k] = in[i] @ W[k] + bias`` (kernel index not flipped)."""
    assert tensor_stride % 2 == 0
    half = tensor_stride // 2
    offs = kernel_offsets(2, half)
    n = len(coords)
    c = np.asarray(coords, dtype=np.int64)
    out_coords = np.repeat(c, 8, axis=0)
    out_coords[:, 1:] += np.tile(offs, (n, 1))
    out = torch.empty((n, 8, weight.shape[2]), dtype=torch.float32)
    for k in range(8):
        out[:, k, :] = feats @ weight[k]
    out = out.reshape(8 * n, -1)
    if bias is not None:
        out = out + bias.reshape(1, -1)
    return out, out_coords.astype(np.int32)


def prune(feats, coords, mask):
    """``ME.MinkowskiPruning()(x, mask)`` (§8 a8, ``autoencoder.py:237,247``):
    keep rows with mask True, order preserved (Appendix A.9)."""
    keep = np.nonzero(np.asarray(mask))[0]
    return feats.index_select(0, torch.from_numpy(keep)), np.ascontiguousarray(coords[keep])


def topk_mask(logits: torch.Tensor, k: int) -> np.ndarray:
    """``istopk`` for one batch item (``data_utils.py:77-89``): True on the k
    largest logits (``torch.topk`` on CPU, k = min(len, N*rho))."""
    n = logits.numel()
    k = int(min(n, k))
    mask = np.zeros(n, dtype=bool)
    if k > 0:
        _, idx = torch.topk(logits.reshape(-1), k)
        mask[idx.numpy()] = True
    return mask


def isin(coords: np.ndarray, ground_truth: np.ndarray) -> np.ndarray:
    """``isin`` (``data_utils.py:63-75``): membership of coordinate rows."""
    return _Lookup(unique_coords(ground_truth)[0]).find(coords) >= 0


def sort_key(coords: np.ndarray) -> np.ndarray:
    """``array2vector(C, C.max()+1)`` (``data_utils.py:55-61,91-101``):
    key = b + x*S + y*S^2 + z*S^3 with S = max(all entries)+1; argsort of it is
    the canonical symbol order of the bitstream."""
    c = np.asarray(coords, dtype=np.int64)
    step = int(c.max()) + 1
    return c[:, 0] + c[:, 1] * step + c[:, 2] * step ** 2 + c[:, 3] * step ** 3
