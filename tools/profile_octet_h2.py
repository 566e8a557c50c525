"""Timing harness for the full-octet h2 kernels (pre-split f16 features) next to the 3xTF32 full-octet kernels and the
gather h2 kernels on the same decoder level of the vox10 workload; CUDA-event times, L2 flushed in between.

    PCGC_OCTET_H2_VARIANT=0|1|2 python tools/profile_octet_h2.py [--shapes 16x16,16x4,16x1,16x32] [--level 2]
"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from pcgcv2_b200 import ops, synth
from pcgcv2_b200.codec import Codec
from util import load_ckpt

ap = argparse.ArgumentParser()
ap.add_argument("--shapes", default="16x16,16x4,16x1,16x32")
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--level", type=int, default=2)
ap.add_argument("--only", default="", help="run only this kernel family (octet_h2) -- for ncu captures")
ap.add_argument("--h2out", action="store_true", help="write the h2 copy too, as the production path does for decoder.conv2")
args = ap.parse_args()
cache = "/tmp/vox10_seed0.npy"
if os.path.exists(cache): pts = np.load(cache)
else:
    pts = synth.synthetic_vox10(0); np.save(cache, pts)
codec = Codec(load_ckpt("r3"), use_octet_kernels=False, use_h2=False)
codec.record = {}
st = codec.encode(pts); codec.decode(st)
_, keys, stride = codec.record[f"decoder.up{args.level}"]
codec.record = None
n = keys.shape[0]
pkeys = keys[::8] >> 3
pnbr = ops.kernel_map_k3(pkeys, ops.HashTable(pkeys))
nbr, npairs = ops.kernel_map_k3(keys, ops.HashTable(keys), count_pairs=True)
pairs = int(npairs.item())
print(f"variant {os.environ.get('PCGC_OCTET_H2_VARIANT', '0')}: rows {n} parents {n // 8} pairs {pairs} ({pairs / n:.2f} nbrs/row)")
g = torch.Generator().manual_seed(0)


def timeit(run):
    run(); torch.cuda.synchronize()
    ts = []
    for _ in range(args.reps):
        flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda").fill_(1.0)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); run(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return float(np.median(ts))


for shape in args.shapes.split(","):
    cin, cout = map(int, shape.split("x"))
    f = torch.randn(n, cin, generator=g).cuda()
    w = (torch.randn(27, cin, cout, generator=g) / np.sqrt(27 * cin)).cuda()
    b = torch.randn(1, cout, generator=g).cuda()
    ph = ops.PackedK3H2(w)
    xh = ops.split_h2(f)
    alg = 4 * (n * cin + n * cout) + 8 * pairs + 4 * 27 * cin * cout
    h2o = args.h2out and cout % 4 == 0
    ms = timeit(lambda: ops.conv_k3_octet_h2(xh, pnbr, ph, b, relu=True, want_h2=h2o))
    line = f"{shape:8s} octet-h2 {ms:.4f} ms  alg {alg / 1e6:.1f} MB  {alg / ms / 1e6:.1f} GB/s  {2 * pairs * cin * cout / ms / 1e9:.2f} TFLOP/s"
    if ops.PackedK3OctetTc05.supported(cin, cout):
        pt = ops.PackedK3OctetTc05(w)
        ms_t = timeit(lambda: ops.conv_k3_octet_tc05(xh, pnbr, pt, b, relu=True, want_h2=h2o))
        ref_t = ops.conv_k3_octet_h2(xh, pnbr, ph, b, relu=True)[0]
        got_t = ops.conv_k3_octet_tc05(xh, pnbr, pt, b, relu=True)[0]
        line += (f"   octet-tcgen05 {ms_t:.4f} ms ({ms / ms_t:.2f}x, {alg / ms_t / 1e6:.0f} GB/s alg, diff "
                 f"{float((got_t - ref_t).abs().max() / ref_t.abs().max()):.1e})")
    if not args.only:
        if cout % 4 == 0:
            line += f"  (+h2 out {timeit(lambda: ops.conv_k3_octet_h2(xh, pnbr, ph, b, relu=True, want_h2=True)):.4f})"
        line += f"   gather-h2 {timeit(lambda: ops.conv_k3_h2(xh, nbr, ph, b, relu=True)):.4f} ms"
        po = ops.PackedK3Octet(w)
        if po.packed is not None:
            ref = ops.conv_k3_octet(f, pnbr, po, b, relu=True)
            line += f"   octet-tf32 {timeit(lambda: ops.conv_k3_octet(f, pnbr, po, b, relu=True)):.4f} ms"
            got = ops.conv_k3_octet_h2(xh, pnbr, ph, b, relu=True)[0]
            line += f"   maxdiff/max {float((got - ref).abs().max() / ref.abs().max()):.1e}"
        line += f"   split {timeit(lambda: ops.split_h2(f, out=xh)):.4f} ms"
    print(line)
