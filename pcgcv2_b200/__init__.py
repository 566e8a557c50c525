"""pcgcv2_b200 -- B200-native (sm_100a) hot path of the PCGCv2 point-cloud geometry codec.

Layout: ``csrc/`` hand-written CUDA + the C ABI (``include/pcgc.h`` -> ``lib/libpcgc.so``),
``ops.py`` the ctypes host layer, ``shim/`` drop-in ``MinkowskiEngine`` / ``torchac`` /
``data_utils`` modules mirroring the reference's operator surface, ``codec.py`` the
encode/decode pipeline.  There is no CPU fallback anywhere in this package.
"""
import os
import sys

__version__ = "0.1.0"
SHIM_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shim")


def install_shims():
    """Make ``import MinkowskiEngine``, ``import torchac`` and ``import data_utils`` resolve to
    the drop-in modules (what a user of the reference does once, before importing its files)."""
    if SHIM_DIR not in sys.path:
        sys.path.insert(0, SHIM_DIR)
