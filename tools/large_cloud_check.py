"""BASELINE config 4 stand-in: vox11-scale cloud (~3.2 M voxels on a 2048^3 grid) through encode+decode."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from pcgcv2_b200 import synth
from pcgcv2_b200.codec import Codec
from util import load_ckpt
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
pts = synth.synthetic_vox10(0, scale=scale)
print("voxels", len(pts), "grid", int(1024 * scale))
codec = Codec(load_ckpt("r3"))
for it in range(3):
    torch.cuda.synchronize(); t = time.perf_counter()
    st = codec.encode(pts); torch.cuda.synchronize(); t1 = time.perf_counter()
    dec = codec.decode(st); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"enc {1e3*(t1-t):.1f} ms dec {1e3*(t2-t1):.1f} ms  N3 {len(st.coords)}  F bytes {len(st.F)}  bpp {st.bits()/len(pts):.4f}  out {len(dec)}  "
          f"{len(pts)/(t2-t)/1e6:.1f} Mpoints/s  peak mem {torch.cuda.max_memory_allocated()/2**30:.2f} GiB")
cells = set(map(tuple, np.unique(pts // 8, axis=0).tolist()))
assert len(dec) == len(pts) and set(map(tuple, np.unique(dec // 8, axis=0).tolist())) <= cells
inter = len(set(map(tuple, pts.tolist())) & set(map(tuple, dec.tolist())))
print("IoU", inter / (2 * len(pts) - inter))
