"""Seeded synthetic voxel clouds (stand-ins for the 8iVFB PLYs, which are not
redistributable / not on disk)."""
import numpy as np


def _shell(rng, centre, radii, n):
    u = rng.normal(size=(n, 3))
    u /= np.linalg.norm(u, axis=1, keepdims=True)
    return np.round(np.asarray(centre) + u * np.asarray(radii)).astype(np.int64)


def random_cube(seed=0, size=32, p=0.1):
    """Config 1: random-occupancy cube (3 339 voxels for seed 0, 32^3, p=0.1)."""
    rng = np.random.default_rng(seed)
    occ = rng.random((size, size, size)) < p
    return np.argwhere(occ).astype(np.int32)


def ellipsoid_vox8(seed=0, n=3_000_000):
    """Appendix E.7 cloud: N0 = 91 568 for seed 0 (res 256)."""
    rng = np.random.default_rng(seed)
    pts = _shell(rng, [128, 128, 128], [70, 50, 100], n)
    return np.unique(np.clip(pts, 0, 255), axis=0).astype(np.int32)


def synthetic_vox10(seed=0, scale=1.0, jitter=0.0):
    """Config 2 stand-in (Appendix E.5): two ellipsoid shells on a 1024^3 grid,
    N0 = 795 124 for seed 0.  ``scale`` shrinks/grows the grid (0.5 -> vox9,
    2 -> vox11) with the sample count scaled by scale^2; ``jitter`` perturbs
    the radii by +-jitter (config 3 frames)."""
    rng = np.random.default_rng(seed)
    res = int(round(1024 * scale))
    j = 1.0 + (rng.uniform(-jitter, jitter, size=2) if jitter else np.zeros(2))
    n1 = int(6_000_000 * scale * scale)
    n2 = int(2_500_000 * scale * scale)
    a = _shell(rng, np.array([512, 512, 300]) * scale, np.array([170, 120, 280]) * scale * j[0], n1)
    b = _shell(rng, np.array([512, 512, 760]) * scale, np.array([110, 110, 130]) * scale * j[1], n2)
    pts = np.concatenate([a, b])
    return np.unique(np.clip(pts, 0, res - 1), axis=0).astype(np.int32)
