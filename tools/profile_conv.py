"""Profiling harness: replay single k=3 convolutions of the finest decoder scale of the vox10 workload
(decoder.conv2: 16->16 on 8*N1 = 1.69 M rows, and the 16->4 / 4->8 / 16->1 shapes on the same map) so that
`ncu --set full -k regex:conv_k3` captures just these launches.

    python tools/profile_conv.py [--shapes 16x16,16x4] [--reps 3]
"""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from pcgcv2_b200 import ops, synth
from pcgcv2_b200.codec import Codec
from util import load_ckpt

ap = argparse.ArgumentParser()
ap.add_argument("--shapes", default="16x16")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--ffma", action="store_true")
ap.add_argument("--dump", default="")
args = ap.parse_args()
cache = "/tmp/vox10_seed0.npy"
pts = np.load(cache) if os.path.exists(cache) else synth.synthetic_vox10(0)
codec = Codec(load_ckpt("r3"), use_tensor_cores=False)     # prep without mma kernels: ncu -k regex:mma sees only the replays
codec.record = {}
st = codec.encode(pts); codec.decode(st)
x, keys, stride = codec.record["decoder.up2"]
codec.record = None
n = keys.shape[0]
nbr, npairs = ops.kernel_map_k3(keys, ops.HashTable(keys), count_pairs=True)
pairs = int(npairs.item())
print(f"rows {n} pairs {pairs} ({pairs / n:.2f} nbrs/row)")
if args.dump:
    with open(args.dump, "wb") as fh:
        fh.write(np.array([n, pairs], dtype=np.int64).tobytes())
        fh.write(nbr.cpu().numpy().tobytes())
g = torch.Generator().manual_seed(0)
for shape in args.shapes.split(","):
    cin, cout = map(int, shape.split("x"))
    f = torch.randn(n, cin, generator=g).cuda() if cin != x.shape[1] else x
    w = (torch.randn(27, cin, cout, generator=g) / np.sqrt(27 * cin)).cuda()
    b = torch.randn(1, cout, generator=g).cuda()
    pw = ops.PackedK3(w)
    use_mma = pw.packed is not None and not args.ffma
    run = (lambda: ops.conv_k3_packed(f, nbr, pw, b, relu=True)) if use_mma else (lambda: ops.conv_k3(f, nbr, w, b, relu=True))
    run(); torch.cuda.synchronize()
    ts = []
    for _ in range(args.reps):
        flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda").fill_(1.0)   # 256 MB: evict L2
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); run(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    alg = 4 * (n * cin + n * cout) + 8 * pairs + 4 * 27 * cin * cout
    ms = float(np.median(ts))
    print(f"{shape:8s} {'mma ' if use_mma else 'ffma'} {ms:.4f} ms  alg {alg / 1e6:.1f} MB  {alg / ms / 1e6:.1f} GB/s  "
          f"{2 * pairs * cin * cout / ms / 1e9:.2f} TFLOP/s")
