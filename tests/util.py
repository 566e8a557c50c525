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
