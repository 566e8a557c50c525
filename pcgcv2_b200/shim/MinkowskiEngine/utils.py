"""``ME.utils`` subset used by the reference (data_utils.py:107,115; data_loader.py:54)."""
import numpy as np
import torch


def sparse_collate(coords, feats, labels=None, dtype=torch.int32, device=None):
    """Prepend the batch index to every coordinate list and concatenate:
    returns (int32 [N, 1+3] coordinates, features[, labels])."""
    use_label = labels is not None
    bcoords, bfeats, blabels = [], [], []
    for b, (c, f) in enumerate(zip(coords, feats)):
        if isinstance(c, np.ndarray):
            c = torch.from_numpy(c)
        if isinstance(f, np.ndarray):
            f = torch.from_numpy(f)
        if c.shape[0] != f.shape[0]:
            raise ValueError("coordinates and features of one batch item differ in length")
        c = c.to(dtype)
        batch_col = torch.full((c.shape[0], 1), b, dtype=dtype, device=c.device)
        bcoords.append(torch.cat([batch_col, c], dim=1))
        bfeats.append(f)
        if use_label:
            l = labels[b]
            blabels.append(torch.from_numpy(l) if isinstance(l, np.ndarray) else l)
    bcoords = torch.cat(bcoords, 0)
    bfeats = torch.cat(bfeats, 0)
    if device is not None:
        bcoords, bfeats = bcoords.to(device), bfeats.to(device)
    if use_label:
        return bcoords, bfeats, torch.cat(blabels, 0)
    return bcoords, bfeats


def batched_coordinates(coords, dtype=torch.int32, device=None):
    feats = [torch.zeros((len(c), 1)) for c in coords]
    return sparse_collate(coords, feats, dtype=dtype, device=device)[0]
