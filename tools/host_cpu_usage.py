"""How many host cores does one rank of the bench keep busy?  Runs the pipelined device-resident step of bench.py for a few
seconds and prints process CPU time / wall time and the busiest threads (psutil per-thread user + system time).
    [taskset -c 0-3] python tools/host_cpu_usage.py [depth]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, psutil, torch, threading
from pcgcv2_b200 import synth
from pcgcv2_b200.pipeline import FramePipeline
from util import load_ckpt

depth = int(sys.argv[1]) if len(sys.argv) > 1 else 4
torch.set_num_threads(1)
pts = synth.synthetic_vox10(0)
pipe = FramePipeline(load_ckpt("r3"), depth=depth, coord_bits=10)
dev = [torch.from_numpy(pts).cuda()] * depth
for _ in range(3):
    pipe.roundtrip(dev, to_host=False)
torch.cuda.synchronize()
proc = psutil.Process()
names = {t.native_id: t.name for t in threading.enumerate()}
t_before = {t.id: (t.user_time, t.system_time) for t in proc.threads()}
c0, w0 = time.process_time(), time.perf_counter()
steps = 40
for _ in range(steps):
    pipe.roundtrip(dev, to_host=False)
torch.cuda.synchronize()
c1, w1 = time.process_time(), time.perf_counter()
wall, cpu = w1 - w0, c1 - c0
print(f"depth {depth}: {1e3 * wall / steps / depth:.2f} ms per frame, {len(pts) * depth * steps / wall / 1e6:.1f} Mpoints/s; "
      f"process CPU {cpu:.2f} s over {wall:.2f} s wall = {cpu / wall:.2f} cores busy ({len(os.sched_getaffinity(0))} allowed); "
      f"{1e3 * cpu / steps / depth:.2f} ms of CPU per frame")
rows = []
for t in proc.threads():
    u0, s0 = t_before.get(t.id, (0.0, 0.0))
    rows.append((t.user_time - u0, t.system_time - s0, t.id))
for u, s, tid in sorted(rows, key=lambda r: -(r[0] + r[1]))[:12]:
    print(f"  thread {names.get(tid, '?'):<22} user {1e3 * u / steps / depth:6.2f}  system {1e3 * s / steps / depth:6.2f}  ms per frame")
pipe.close()
