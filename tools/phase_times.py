"""Wall-clock per phase of one encode+decode (with a synchronize after each phase) -- where does the step go?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from pcgcv2_b200 import ops, synth
from pcgcv2_b200 import codec as codec_mod
from pcgcv2_b200.codec import Codec
from util import load_ckpt

cache = "/tmp/vox10_seed0.npy"
if os.path.exists(cache): pts = np.load(cache)
else:
    pts = synth.synthetic_vox10(0); np.save(cache, pts)
codec = Codec(load_ckpt("r3"))
dev_coords = torch.from_numpy(pts).cuda()
for _ in range(3):
    st = codec.encode(dev_coords); codec.decode(st, to_host=False)
torch.cuda.synchronize()

class T:
    def __init__(self): self.t = {}; self.last = None
    def mark(self, name):
        torch.cuda.synchronize(); now = time.perf_counter()
        if self.last is not None: self.t[name] = self.t.get(name, 0) + (now - self.last) * 1e3
        self.last = now
reps = 5
tm = T()
for _ in range(reps):
    tm.mark("_")
    coords = torch.cat([torch.zeros((len(dev_coords), 1), dtype=torch.int32, device="cuda"), dev_coords], 1)
    level0, _ = codec._sorted_input(coords, False); tm.mark("enc: pack+sort input")
    y, level3, num_points = codec.analysis(level0); tm.mark("enc: analysis network (incl. maps)")
    c3 = ops.unpack_keys(level3.keys, 1)[:, 1:]; order = codec._canonical_order(c3); y = y[order].contiguous(); c3 = c3[order]
    sym, lo, hi = ops.eb_quantize(y); _, table = ops.eb_cdf_table(codec.eb_params, lo, hi)
    tab_h, sym_h = table.cpu().numpy(), sym.cpu().numpy(); tm.mark("enc: sort bottleneck, quantise, table, D2H")
    f_bytes = ops.rc_encode_u16(tab_h, sym_h); tm.mark("enc: host range encode")
    st = codec.encode(dev_coords); tm.mark("(full encode again)")
    # decode phases
    n3 = len(st.coords)
    _, table = ops.eb_cdf_table(codec.eb_params, lo, hi); tab_h = table.cpu().numpy(); tm.mark("dec: table")
    sym = ops.rc_decode_u16(tab_h, st.F, n3 * 8); tm.mark("dec: host range decode")
    out = codec.decode(st, to_host=False); tm.mark("(full decode)")
for k, v in tm.t.items():
    print(f"{k:45s} {v / reps:8.3f} ms")
# synthesis only
c3 = torch.as_tensor(st.coords).cuda()
yq = torch.from_numpy(sym.reshape(n3, 8).astype(np.float32)).cuda() + float(lo)
keys = ops.pack_keys(torch.cat([torch.zeros((n3, 1), dtype=torch.int32, device="cuda"), c3 * 8], 1), 8)
keys, order = ops.argsort_u64(keys)
nums = np.frombuffer(st.num_points, np.int32).tolist()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(reps):
    codec.synthesis(yq[order.long()].contiguous(), codec_mod._Level(keys, 8), nums)
torch.cuda.synchronize(); print(f"{'dec: synthesis network only':45s} {(time.perf_counter() - t0) / reps * 1e3:8.3f} ms")
