"""Host-side pieces of the path that need no GPU: the coordinate coders (row f1) and the native PLY I/O (row f2)."""
import os

import numpy as np
import pytest
import torch

from oracle import refbin
from pcgcv2_b200 import ops, synth
from pcgcv2_b200.coords_coder import OctreeCoordinateCoder, Tmc3CoordinateCoder
from util import GOLDEN, canon


def test_octree_coder_round_trips_and_stays_near_tmc3():
    g = np.load(os.path.join(GOLDEN, "oracle_vox10_seed0.npz"))
    c3 = g["C_coords"].astype(np.int32)
    oc = OctreeCoordinateCoder()
    blob = oc.encode(c3)
    assert (canon(oc.decode(blob)) == canon(c3)).all() and len(blob) < 1.6 * int(g["C_bytes"])
    for n in (0, 1, 2, 7):                                                  # empty and tiny sets
        assert (canon(oc.decode(oc.encode(c3[:n]))) == canon(c3[:n])).all()
    assert oc.decode(oc.encode(np.zeros((1, 3), np.int32))).tolist() == [[0, 0, 0]]
    rng = np.random.default_rng(0)
    p = rng.integers(0, 1 << 20, size=(4000, 3)).astype(np.int32)          # deep, scattered, with duplicates
    p = np.concatenate([p, p[:100]])
    assert (canon(oc.decode(oc.encode(p))) == canon(np.unique(p, axis=0))).all()
    with pytest.raises(Exception):
        oc.encode(np.array([[-1, 0, 0]], np.int32))
    with pytest.raises(Exception):
        oc.decode(b"nonsense-bytes")
    bad = bytearray(blob)
    bad[20] ^= 0x55                                                          # corruption is detected or yields another set, never a crash
    try:
        oc.decode(bytes(bad))
    except Exception:
        pass


@pytest.mark.skipif(not refbin.available(), reason="reference binaries not installed (oracle/_ref)")
def test_tmc3_coder_is_the_reference_command_line():
    g = np.load(os.path.join(GOLDEN, "oracle_vox10_seed0.npz"))
    c3 = g["C_coords"].astype(np.int32)
    tc = Tmc3CoordinateCoder(refbin.TMC3)
    blob = tc.encode(c3)
    assert len(blob) == int(g["C_bytes"]) and (canon(tc.decode(blob)) == canon(c3)).all()
    with pytest.raises(FileNotFoundError):
        Tmc3CoordinateCoder("/nonexistent/tmc3")


def test_ply_io_native_matches_reference_semantics(tmp_path):
    """f2: native ASCII PLY writer / reader (csrc/ply.cpp) against the reference's line semantics (data_utils.py:19-48)."""
    pts = synth.ellipsoid_vox8()
    f = str(tmp_path / "a.ply")
    ops.ply_write_ascii(f, pts)
    assert (refbin.read_ply(f) == pts).all()                               # what the reference's reader sees
    pin = torch.cuda.is_available()                                         # pinned staging needs a CUDA context
    back = ops.ply_read_ascii(f, pinned=pin)
    assert back.is_pinned() == pin and back.dtype == torch.int32 and (back.numpy() == pts).all()
    refbin.write_ply(f, pts[:1000])                                        # the reference writer's output
    assert (ops.ply_read_ascii(f).numpy() == pts[:1000]).all()
    odd = str(tmp_path / "b.ply")
    with open(odd, "w") as fh:
        fh.write("ply\nformat ascii 1.0\ncomment 1 2 3\nelement vertex 4\nproperty float x\nend_header\n"
                 "1 2 3\n4.7 -5.2 6e0 9 9\n7  8 9\n-1 -2 -3 \n10 11 12")
    assert ops.ply_read_ascii(odd).numpy().tolist() == [[1, 2, 3], [4, -5, 6], [-1, -2, -3], [10, 11, 12]]
