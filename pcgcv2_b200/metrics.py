"""D1 (point-to-point) geometry PSNR on the GPU: the numbers the reference gets from the
``pc_error_d`` subprocess (pc_error.py:44-54, coder.py:181-184), computed from the device-resident voxel sets with the
coordinate hash (csrc/metrics.cu).  Exact: squared distances are integers, summed in 64-bit integers."""
from __future__ import annotations

import math

import torch

from . import _lib, ops


def _keys(coords) -> torch.Tensor:
    c = torch.as_tensor(coords)
    if not c.is_cuda:
        c = c.cuda()
    c = c.to(torch.int32)
    if c.shape[1] == 3:
        c = torch.nn.functional.pad(c, (1, 0))
    return torch.unique(ops.pack_keys(c.contiguous(), 1))        # pc_error drops duplicate points (dropDuplicates = 2)


def _direction(q, table, cloud, max_radius):
    acc = torch.empty(3, dtype=torch.int64, device=q.device)
    scratch = torch.empty(max(q.shape[0], 1), dtype=torch.int32, device=q.device)
    _lib.check(_lib.lib().pcgc_d1_sqdist(q.data_ptr(), q.shape[0], table.tkeys.data_ptr(), table.cap, cloud.data_ptr(), cloud.shape[0],
                                         int(max_radius), acc.data_ptr(), scratch.data_ptr(), ops._stream()), "pcgc_d1_sqdist")
    return acc


def d1(a, b, res: int, max_radius: int = 4) -> dict:
    """symmetric D1 metrics of two voxel clouds (int [N,3] or [N,4] arrays / tensors), keyed like pc_error.py's DataFrame."""
    ka, kb = _keys(a), _keys(b)
    ta, tb = ops.HashTable(ka), ops.HashTable(kb)
    acc = torch.stack([_direction(ka, tb, kb, max_radius), _direction(kb, ta, ka, max_radius)]).cpu().tolist()   # one sync
    peak = float(res - 1)
    psnr = lambda mse: float("inf") if mse == 0 else 10.0 * math.log10(3.0 * peak * peak / mse)
    mse1, mse2 = acc[0][0] / max(len(ka), 1), acc[1][0] / max(len(kb), 1)
    h1, h2 = float(acc[0][1]), float(acc[1][1])
    return {"mse1      (p2point)": mse1, "mse1,PSNR (p2point)": psnr(mse1), "h.       1(p2point)": h1, "h.,PSNR  1(p2point)": psnr(h1),
            "mse2      (p2point)": mse2, "mse2,PSNR (p2point)": psnr(mse2), "h.       2(p2point)": h2, "h.,PSNR  2(p2point)": psnr(h2),
            "mseF      (p2point)": max(mse1, mse2), "mseF,PSNR (p2point)": psnr(max(mse1, mse2)),
            "h.        (p2point)": max(h1, h2), "h.,PSNR   (p2point)": psnr(max(h1, h2)),
            "brute_force_queries": int(acc[0][2] + acc[1][2])}


def d1_psnr(a, b, res: int) -> float:
    return d1(a, b, res)["mseF,PSNR (p2point)"]
