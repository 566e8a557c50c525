"""One encode + decode of the vox10 stand-in (for `ncu -k ... -c 1` captures of a production kernel in its production context).
    ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:h2c4_dual -c 1 -o out python tools/one_frame.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from pcgcv2_b200 import synth
from pcgcv2_b200.codec import Codec
from util import load_ckpt

cache = "/tmp/vox10_seed0.npy"
pts = np.load(cache) if os.path.exists(cache) else synth.synthetic_vox10(0)
codec = Codec(load_ckpt("r3"), coord_bits=10)
st = codec.encode(pts)
out = codec.decode(st)
torch.cuda.synchronize()
print(len(pts), len(out), st.bits())
