"""Shared helpers for the tests (fixtures under tests/golden/)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_ckpt(name="r3"):
    """state_dict (name -> float32 torch tensor) from tests/golden/ckpt_<name>.npz."""
    z = np.load(os.path.join(GOLDEN, f"ckpt_{name}.npz"))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def with_batch(pts, b=0):
    pts = np.asarray(pts, dtype=np.int32)
    return np.concatenate([np.full((len(pts), 1), b, np.int32), pts], axis=1)


def canon(coords):
    """rows sorted lexicographically -- the canonical order for set comparison."""
    c = np.asarray(coords)
    return c[np.lexsort(c.T[::-1])]


def octree_pack(coords3):
    """int [N,3] voxel set -> (root int32 [M,3] = occupied 8x8x8 cells in lexicographic (z,y,x)-major canonical order,
    [occ0, occ1, occ2]) with one occupancy byte (bit ix + 2 iy + 4 iz) per node of the levels at stride 8, 4 and 2,
    nodes in breadth-first order (children in ascending bit order).  Exact and ~40x smaller than the coordinates."""
    c = np.unique(np.asarray(coords3, dtype=np.int64), axis=0)
    levels = [c]
    for _ in range(3):
        levels.append(np.unique(levels[-1] >> 1, axis=0))
    key = lambda a: (a[:, 2] << 42) | (a[:, 1] << 21) | a[:, 0]
    root = levels[3][np.argsort(key(levels[3]))]
    occ, nodes = [], root
    for lvl in (2, 1, 0):
        child = levels[lvl]
        pk, ck = key(child >> 1), ((child[:, 0] & 1) | ((child[:, 1] & 1) << 1) | ((child[:, 2] & 1) << 2))
        order = np.argsort(key(nodes))
        idx = order[np.searchsorted(key(nodes)[order], pk)]
        byte = np.zeros(len(nodes), dtype=np.uint8)
        np.bitwise_or.at(byte, idx, (1 << ck).astype(np.uint8))
        occ.append(byte)
        nodes = octree_children(nodes, byte)
    return root.astype(np.int32), occ


def octree_children(nodes, byte):
    bits = (byte[:, None] >> np.arange(8)[None, :]) & 1
    ni, k = np.nonzero(bits)
    return np.stack([2 * nodes[ni, 0] + (k & 1), 2 * nodes[ni, 1] + ((k >> 1) & 1), 2 * nodes[ni, 2] + ((k >> 2) & 1)], axis=1)


def octree_unpack(root, occ):
    nodes = np.asarray(root, dtype=np.int64)
    for byte in occ:
        nodes = octree_children(nodes, np.asarray(byte))
    return nodes.astype(np.int32)
