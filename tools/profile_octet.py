"""Profiling harness for the full-octet k=3 kernels: replays single convolutions of the finest decoder scale of
the vox10 workload (8*N1 = 1.69 M rows, parent map of the 211 k kept rows) so that
`ncu --set full -k regex:octet` captures just these launches; prints CUDA-event times (L2 flushed in between).

    python tools/profile_octet.py [--shapes 16x16,16x4,4x8] [--reps 5] [--level 2]
"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from pcgcv2_b200 import ops, synth
from pcgcv2_b200.codec import Codec
from util import load_ckpt

ap = argparse.ArgumentParser()
ap.add_argument("--shapes", default="16x16,16x4,16x1,4x8,4x4")
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--level", type=int, default=2, help="decoder scale: 0 (110 k rows), 1 (435 k), 2 (1.69 M)")
ap.add_argument("--child", action="store_true", help="also time the child-map kernels on the same set")
args = ap.parse_args()
cache = "/tmp/vox10_seed0.npy"
pts = np.load(cache) if os.path.exists(cache) else synth.synthetic_vox10(0)
codec = Codec(load_ckpt("r3"), use_octet_kernels=False)
codec.record = {}
st = codec.encode(pts); codec.decode(st)
_, keys, stride = codec.record[f"decoder.up{args.level}"]
codec.record = None
n = keys.shape[0]
pkeys = keys[::8] >> 3                                           # parents (batch 0): Morton field one level up
pnbr = ops.kernel_map_k3(pkeys, ops.HashTable(pkeys))
nbr, npairs = ops.kernel_map_k3(keys, ops.HashTable(keys), count_pairs=True)
assert torch.equal(nbr, ops.kernel_map_k3_from_parent(pnbr, n))
pairs = int(npairs.item())
print(f"rows {n} parents {n // 8} pairs {pairs} ({pairs / n:.2f} nbrs/row)")
g = torch.Generator().manual_seed(0)


def timeit(run):
    run(); torch.cuda.synchronize()
    ts = []
    for _ in range(args.reps):
        flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda").fill_(1.0)   # 256 MB: evict L2
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); run(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return float(np.median(ts))


for shape in args.shapes.split(","):
    cin, cout = map(int, shape.split("x"))
    f = torch.randn(n, cin, generator=g).cuda()
    w = (torch.randn(27, cin, cout, generator=g) / np.sqrt(27 * cin)).cuda()
    b = torch.randn(1, cout, generator=g).cuda()
    po = ops.PackedK3Octet(w)
    alg = 4 * (n * cin + n * cout) + 8 * pairs + 4 * 27 * cin * cout
    ms = timeit(lambda: ops.conv_k3_octet(f, pnbr, po, b, relu=True))
    line = f"{shape:8s} octet {ms:.4f} ms  alg {alg / 1e6:.1f} MB  {alg / ms / 1e6:.1f} GB/s  {2 * pairs * cin * cout / ms / 1e9:.2f} TFLOP/s"
    if args.child:
        pw = ops.PackedK3(w)
        run = (lambda: ops.conv_k3_packed(f, nbr, pw, b, relu=True)) if pw.packed is not None else (lambda: ops.conv_k3(f, nbr, w, b, relu=True))
        line += f"   child-map kernel {timeit(run):.4f} ms"
    print(line)
