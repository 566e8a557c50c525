"""Warm, in-pipeline kernel times of one encode+decode from CUPTI (torch.profiler sees kernels launched through
ctypes too): per-kernel totals, GPU busy time vs wall time, and the largest idle gaps.  Not a bench number."""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
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
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(reps):
        st = codec.encode(host); torch.cuda.synchronize()
        codec.decode(st); torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
tot = collections.defaultdict(lambda: [0.0, 0])
for e in evs:
    d = e.time_range.end - e.time_range.start
    tot[e.name][0] += d; tot[e.name][1] += 1
busy = sum(v[0] for v in tot.values())
span = evs[-1].time_range.end - evs[0].time_range.start
print(f"{reps} passes: GPU busy {busy / reps / 1e3:.3f} ms/pass, first-to-last span {span / reps / 1e3:.3f} ms/pass, "
      f"{len(evs) / reps:.0f} device activities/pass")
for name, (d, c) in sorted(tot.items(), key=lambda kv: -kv[1][0])[:60]:
    print(f"{d / reps / 1e3:8.3f} ms/pass {100 * d / busy:5.1f}% x{c / reps:5.1f}  {name[:110]}")
gaps = []
for a, b in zip(evs[:-1], evs[1:]):
    g = b.time_range.start - a.time_range.end
    if g > 20: gaps.append((g, a.name[:50], b.name[:50]))
gaps.sort(reverse=True)
print(f"idle gaps > 20 us: {len(gaps) / reps:.0f}/pass, total {sum(g[0] for g in gaps) / reps / 1e3:.3f} ms/pass")
for g in gaps[:25]:
    print(f"  {g[0]:8.1f} us  after {g[1]}  before {g[2]}")
