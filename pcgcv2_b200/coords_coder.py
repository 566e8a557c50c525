"""Coordinate side channel of the bitstream (``_C.bin``): the ~14 k stride-8 bottleneck coordinates, lossless
(reference ``CoordinateCoder``, coder.py:17-36 -> gpcc.py:6-36 -> external ``tmc3``).

Two interchangeable coders with the interface ``encode(int32 [n,3]) -> bytes`` / ``decode(bytes) -> int32 [n,3]``:

* ``OctreeCoordinateCoder`` -- in-process (libpcgc host code, csrc/octree_coder.cpp): breadth-first octree occupancy
  with neighbour-conditioned adaptive binary range coding.  No subprocess, no files, ~1 ms per frame; its own
  format (NOT a G-PCC stream).  This is what ``bench.py`` times.
* ``Tmc3CoordinateCoder`` -- the reference's path: MPEG G-PCC ``tmc3`` as an external executable with exactly the
  command lines of gpcc.py:11-21,30-36, ASCII PLY hand-over through a tmpfs directory.  The binary is NOT part of
  this package (the reference bundles it next to gpcc.py; the caller passes its path) -- with it ``_C.bin`` is
  byte-identical to the reference's, so the total bit count is too.

``Codec`` runs the coordinate coder on a side thread, overlapped with the host range coder of the features.
"""
from __future__ import annotations

import os
import subprocess
import tempfile
import threading

import numpy as np

from . import _lib, ops


class OctreeCoordinateCoder:
    name = "octree (in-process, own format)"

    def encode(self, coords3) -> bytes:
        c = np.ascontiguousarray(np.asarray(coords3, dtype=np.int32)).reshape(-1, 3)
        cap = 64 + 8 * c.shape[0]
        out = np.empty(cap, dtype=np.uint8)
        n = _lib.check(_lib.lib().pcgc_octree_encode_host(c.ctypes.data if c.size else None, c.shape[0], out.ctypes.data, cap),
                       "pcgc_octree_encode_host")
        return out[:n].tobytes()

    def decode(self, data: bytes) -> np.ndarray:
        buf = np.frombuffer(data, dtype=np.uint8)
        L = _lib.lib()
        n = _lib.check(L.pcgc_octree_decode_host(buf.ctypes.data, buf.size, None, 0), "pcgc_octree_decode_host")   # count only
        out = np.empty((max(n, 1), 3), dtype=np.int32)
        n = _lib.check(L.pcgc_octree_decode_host(buf.ctypes.data, buf.size, out.ctypes.data, n), "pcgc_octree_decode_host")
        return out[:n]


class Tmc3CoordinateCoder:
    name = "tmc3 (external G-PCC binary, reference command line)"
    _seq = 0
    _lock = threading.Lock()

    def __init__(self, tmc3_path, workdir=None):
        if not (os.path.isfile(tmc3_path) and os.access(tmc3_path, os.X_OK)):
            raise FileNotFoundError(f"tmc3 executable not found at {tmc3_path!r} (the reference bundles it next to gpcc.py)")
        self.tmc3 = tmc3_path
        base = workdir or ("/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None)
        self._tmp = tempfile.TemporaryDirectory(prefix="pcgc_tmc3_", dir=base)

    def _names(self):
        with Tmc3CoordinateCoder._lock:
            Tmc3CoordinateCoder._seq += 1
            k = Tmc3CoordinateCoder._seq
        d = self._tmp.name
        return os.path.join(d, f"{k}.ply"), os.path.join(d, f"{k}.bin")

    def encode(self, coords3) -> bytes:
        ply, out = self._names()
        ops.ply_write_ascii(ply, np.asarray(coords3, dtype=np.int32))
        try:
            subprocess.run([self.tmc3, "--mode=0", "--positionQuantizationScale=1", "--trisoupNodeSizeLog2=0",
                            "--neighbourAvailBoundaryLog2=8", "--intra_pred_max_node_size_log2=6",
                            "--inferredDirectCodingMode=0", "--maxNumQtBtBeforeOt=4", "--uncompressedDataPath=" + ply,
                            "--compressedStreamPath=" + out], check=True, stdout=subprocess.DEVNULL)      # gpcc.py:11-21
            with open(out, "rb") as f:
                return f.read()
        finally:
            for p in (ply, out):
                if os.path.exists(p):
                    os.remove(p)

    def decode(self, data: bytes) -> np.ndarray:
        ply, binf = self._names()
        with open(binf, "wb") as f:
            f.write(data)
        try:
            subprocess.run([self.tmc3, "--mode=1", "--compressedStreamPath=" + binf, "--reconstructedDataPath=" + ply,
                            "--outputBinaryPly=0"], check=True, stdout=subprocess.DEVNULL)                # gpcc.py:30-36
            return ops.ply_read_ascii(ply).numpy().copy()
        finally:
            for p in (ply, binf):
                if os.path.exists(p):
                    os.remove(p)
