"""Full-size oracle fixtures for BASELINE.json configs 2, 3 and 4 (run in the BUILD container: needs
/root/reference for the pc_error_d / tmc3 binaries; ~5 minutes of CPU):

    python tests/golden/make_golden_fullsize.py

For every case the CPU oracle (oracle/codec_ref.py) codes the cloud and the outputs the GPU tests compare against
are frozen in tests/golden/oracle_<case>.npz:

  C_coords      int16 [N3,3]   bottleneck coordinates / 8 in the canonical symbol order (coder.py:84,89)
  y_F           float32 [N3,8] bottleneck features before rounding (activation parity at full size)
  F, H, num_points             the three feature-side files of coder.py:49-55,85-87, byte for byte
  C_bytes       int            size of the tmc3-coded coordinate file (gpcc.py:11-21 flags) -> bits of `_C.bin`
  dec_root/dec_occ*            the oracle's decoded voxel set as an octree (stride-8 cells + one occupancy byte per
                               node of the three levels below): exact, ~200 KB instead of 9.5 MB of coordinates
  d1_psnr       float          `mseF,PSNR (p2point)` printed by the reference's pc_error_d binary for
                               (input cloud, oracle-decoded cloud) at --resolution=res-1 (pc_error.py:44-54)
"""
import os
import subprocess
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import codec_ref, refbin  # noqa: E402
from pcgcv2_b200 import synth  # noqa: E402
from util import load_ckpt, octree_pack, with_batch  # noqa: E402

CASES = {
    # config 2: the headline cloud
    "vox10_seed0": dict(cloud=dict(seed=0), res=1024),
    # config 3: one of the four jittered frames (rank 1's)
    "vox10_seed1_jitter": dict(cloud=dict(seed=1, jitter=0.1), res=1024),
    # config 4 at a size the oracle affords: scaling_factor 0.375, rho 4 (coder.py:149-152,166-167; data_utils.py:112-118)
    "vox10_seed0_scale0375_rho4": dict(cloud=dict(seed=0), res=1024, scaling_factor=0.375, rho=4.0),
}


def run_case(name, sd, cloud, res, scaling_factor=1.0, rho=1.0):
    pts = synth.synthetic_vox10(**cloud)
    x_in = codec_ref.scale_coords(pts, scaling_factor) if scaling_factor != 1.0 else pts
    st = codec_ref.encode(sd, with_batch(x_in))
    dec, _ = codec_ref.decode(sd, st, rho=rho)
    dec = dec[:, 1:]
    if scaling_factor != 1.0:
        dec = codec_ref.scale_coords(dec, 1.0 / scaling_factor)
    with tempfile.TemporaryDirectory() as tmp:
        c_bytes = len(refbin.gpcc_encode_coords(st["C_coords"], tmp))
        d1 = refbin.pc_error_d1(pts, dec, res, tmp)
    root, occ = octree_pack(dec)
    out = os.path.join(HERE, f"oracle_{name}.npz")
    np.savez_compressed(out, C_coords=st["C_coords"].astype(np.int16), y_F=st["y_F"].numpy(),
                        F=np.frombuffer(st["F"], np.uint8), H=np.frombuffer(st["H"], np.uint8),
                        num_points=np.frombuffer(st["num_points"], np.uint8), C_bytes=np.int64(c_bytes),
                        dec_root=root, dec_occ0=occ[0], dec_occ1=occ[1], dec_occ2=occ[2], d1_psnr=np.float64(d1),
                        n_in=np.int64(len(pts)), n_coded=np.int64(len(x_in)), n_dec=np.int64(len(dec)))
    print(name, "N", len(pts), "coded", len(x_in), "N3", len(st["C_coords"]), "F", len(st["F"]), "C", c_bytes, "dec", len(dec),
          "D1", d1, "->", os.path.getsize(out), "bytes", flush=True)


if __name__ == "__main__":
    torch.set_flush_denormal(True)
    sd = load_ckpt("r3")
    for name in (sys.argv[1:] or CASES):
        run_case(name, sd, **CASES[name])
