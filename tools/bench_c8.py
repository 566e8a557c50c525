"""cin = 8 k=3 layers: the 3xTF32 kernels (conv_pipe / conv_octet) against the h2 kernels with [hi | lo] in one k-step
(conv_h2.cuh / conv_octet_h2.cuh, CIN = 8), on the coordinate sets of the vox10 workload; CUDA events, L2 flushed.
    python tools/bench_c8.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from pcgcv2_b200 import ops, synth


def timeit(fn, reps=10):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda").fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


dev = torch.device("cuda")
pts = synth.synthetic_vox10(0)
k = ops.argsort_u64(ops.pack_keys(torch.nn.functional.pad(torch.from_numpy(pts).to(dev), (1, 0)), 1))[0]
sets = {}
for name in ("N1 211k", "N2 54k", "N3 13.8k"):
    k = ops.stride_down(k, keys_are_sorted=True)[0].contiguous()
    sets[name] = k
g = torch.Generator().manual_seed(0)
for name, parent in (("N1 211k (gather)", None), ("N3 13.8k (gather)", None), ("8N2 435k (full octets)", "N2 54k")):
    if parent is None:
        keys = sets[name.split(" (")[0]]
        pnbr = None
    else:
        pkeys = sets[parent]
        pnbr = ops.kernel_map_k3(pkeys, ops.HashTable(pkeys))
        keys = ops.upsample_keys(pkeys)
    n = keys.shape[0]
    nbr, npairs = ops.kernel_map_k3(keys, ops.HashTable(keys), count_pairs=True)
    pairs = int(npairs.item())
    print(f"== {name}: rows {n} pairs {pairs}")
    for cin, cout in ((8, 16), (8, 8)):
        x = (torch.randn(n, cin, generator=g) * 2).to(dev)
        w = (torch.randn(27, cin, cout, generator=g) / np.sqrt(27 * cin)).to(dev)
        b = torch.randn(1, cout, generator=g).to(dev)
        xh = ops.split_h2(x)
        ph, pk = ops.PackedK3H2(w), ops.PackedK3(w)
        alg = 4 * (n * cin + n * cout) + 8 * pairs + 4 * 27 * cin * cout
        if pnbr is None:
            t_tf = timeit(lambda: ops.conv_k3_packed(x, nbr, pk, b, relu=True))
            t_h2 = timeit(lambda: ops.conv_k3_h2(xh, nbr, ph, b, relu=True, want_h2=True))
            ref, got = ops.conv_k3_packed(x, nbr, pk, b, relu=True), ops.conv_k3_h2(xh, nbr, ph, b, relu=True)[0]
        else:
            po = ops.PackedK3Octet(w)
            t_tf = timeit(lambda: ops.conv_k3_octet(x, pnbr, po, b, relu=True))
            t_h2 = timeit(lambda: ops.conv_k3_octet_h2(xh, pnbr, ph, b, relu=True, want_h2=True))
            ref, got = ops.conv_k3_octet(x, pnbr, po, b, relu=True), ops.conv_k3_octet_h2(xh, pnbr, ph, b, relu=True)[0]
        err = float((ref - got).abs().max() / ref.abs().max())
        print(f"  {cin}->{cout:<2}  3xTF32 {t_tf:.4f} ms   h2 (fp32 + h2 out) {t_h2:.4f} ms   {t_tf / t_h2:.2f}x   {alg / t_h2 / 1e6:6.0f} GB/s alg   diff {err:.1e}")
