"""Where the HOST time of one encode+decode goes (cProfile over a few passes; the GPU runs asynchronously, so the
python thread's own time per call is what bounds the launch rate).  Not a bench number."""
import cProfile, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from pcgcv2_b200 import synth
from pcgcv2_b200.codec import Codec
from util import load_ckpt

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
cache = "/tmp/vox10_seed0.npy"
if os.path.exists(cache): pts = np.load(cache)
else:
    pts = synth.synthetic_vox10(0); np.save(cache, pts)
codec = Codec(load_ckpt("r3"))
host = torch.from_numpy(pts).pin_memory()
for _ in range(3):
    st = codec.encode(host); codec.decode(st)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(reps):
    st = codec.encode(host); codec.decode(st)
torch.cuda.synchronize()
print(f"wall {1e3 * (time.perf_counter() - t0) / reps:.3f} ms/pass")
pr = cProfile.Profile()
pr.enable()
for _ in range(reps):
    st = codec.encode(host); codec.decode(st)
torch.cuda.synchronize()
pr.disable()
ps = pstats.Stats(pr)
ps.sort_stats("tottime").print_stats(28)
