"""Same-named module as the reference's ``data_utils.py`` (imported by its autoencoder.py:4,
coder.py:7-8, loss.py:4) with the host hops replaced by libpcgc kernels: ``istopk`` (GPU radix
select instead of D2H + CPU torch.topk, data_utils.py:77-89), ``isin`` (hash probe instead of
np.isin on the host, :63-75), ``sort_spare_tensor`` (device argsort, :91-101).  PLY I/O is one
native pass over the text (csrc/ply.cpp) into pinned int32 memory; h5py is optional."""
import os

import numpy as np
import torch

import MinkowskiEngine as ME
from pcgcv2_b200 import ops as _ops


def read_h5_geo(filedir):
    import h5py
    return h5py.File(filedir, 'r')['data'][:][:, 0:3].astype('int')


def write_h5_geo(filedir, coords):
    import h5py
    data = coords.astype('uint8')
    with h5py.File(filedir, 'w') as h:
        h.create_dataset('data', data=data, shape=data.shape)


def read_ply_ascii_geo(filedir):
    """ASCII PLY -> int [N,3]: the first three values of every line whose tokens all parse as floats (the reference's
    line semantics, data_utils.py:19-34), parsed in one native pass (pcgc_ply_parse_ascii_host, csrc/ply.cpp)."""
    return _ops.ply_read_ascii(filedir).numpy().astype('int')


def write_ply_ascii_geo(filedir, coords):
    """data_utils.py:36-48 without the per-point Python loop (pcgc_ply_format_ascii_host)."""
    _ops.ply_write_ascii(filedir, np.asarray(coords).astype('int'))


def array2vector(array, step):
    """key = sum_i array[:, i] * step**i (kept on the array's device)."""
    array = array.long()
    step = step.long().to(array.device) if isinstance(step, torch.Tensor) else int(step)
    return sum([array[:, i] * (step ** i) for i in range(array.shape[-1])])


def isin(data, ground_truth):
    """bool [len(data)]: rows of ``data`` present in ``ground_truth`` (both int32 [N, 4] on the GPU)."""
    gt_keys = _ops.pack_keys(ground_truth.int().contiguous(), 1)
    return _ops.HashTable(gt_keys).contains(_ops.pack_keys(data.int().contiguous(), 1))


def istopk(data, nums, rho=1.0):
    """bool [len(data)]: True on the k = min(len, N*rho) largest logits of every batch item."""
    rows_per_batch = data._batchwise_row_indices
    if len(rows_per_batch) == 1:
        k = int(min(len(data), nums[0] * rho))
        return _ops.topk_mask(data.F, k)
    mask = torch.zeros(len(data), dtype=torch.bool, device=data.device)
    for rows, N in zip(rows_per_batch, nums):
        k = int(min(len(rows), N * rho))
        mask[rows] = _ops.topk_mask(data.F[rows].contiguous(), k)
    return mask


def sort_spare_tensor(sparse_tensor):
    """Rows re-ordered by key = b + x*S + y*S^2 + z*S^3, S = max+1 -- the bitstream's symbol order."""
    C = sparse_tensor.C
    key = array2vector(C, C.max() + 1)
    _, order = _ops.argsort_u64(key.contiguous())
    order = order.long()
    return ME.SparseTensor(features=sparse_tensor.F[order], coordinates=C[order],
                           tensor_stride=sparse_tensor.tensor_stride[0], device=sparse_tensor.device)


def load_sparse_tensor(filedir, device):
    coords = _ops.ply_read_ascii(filedir, pinned=torch.cuda.is_available())      # int32, pinned: the H2D copy is one DMA
    feats = torch.ones((len(coords), 1)).float()
    coords, feats = ME.utils.sparse_collate([coords], [feats])
    return ME.SparseTensor(features=feats, coordinates=coords, tensor_stride=1, device=device)


def scale_sparse_tensor(x, factor):
    coords = (x.C[:, 1:] * factor).round().int()
    feats = torch.ones((len(coords), 1), device=coords.device).float()
    coords, feats = ME.utils.sparse_collate([coords], [feats])
    return ME.SparseTensor(features=feats, coordinates=coords, tensor_stride=1, device=x.device)
