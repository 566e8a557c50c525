"""D1 (point-to-point) PSNR as the bundled ``pc_error_d`` 0.13.4 reports it.

TEST INFRASTRUCTURE.  Restates what ``pc_error.py:44-54`` asks the binary for
(``-a A -b B --hausdorff=1 --resolution=res-1``) and what ``coder.py:181-184``
reads back (``mseF,PSNR (p2point)``): symmetric nearest-neighbour mean squared
error, ``mseF = max(mse1, mse2)``, ``PSNR = 10*log10(3 * peak^2 / mseF)`` with
``peak = res - 1``.  Cross-checked against the binary itself in the build
container (``tests/test_oracle.py::test_d1_matches_pc_error_binary``).
"""
import numpy as np
from scipy.spatial import cKDTree


def d1_mse(a: np.ndarray, b: np.ndarray):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    d_ab, _ = cKDTree(b).query(a, k=1)
    d_ba, _ = cKDTree(a).query(b, k=1)
    return float(np.mean(d_ab ** 2)), float(np.mean(d_ba ** 2))


def d1_psnr(a: np.ndarray, b: np.ndarray, res: int) -> float:
    mse1, mse2 = d1_mse(a, b)
    mse = max(mse1, mse2)
    peak = float(res - 1)
    if mse == 0:
        return float("inf")
    return 10.0 * np.log10(3.0 * peak * peak / mse)
