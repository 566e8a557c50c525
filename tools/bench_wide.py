"""Time the tcgen05 / TMA wide-layer kernel (csrc/conv_wide.cuh) against the mma.sync h2 kernels it replaces, on the
coordinate sets of synthetic_vox10(0) (Morton order, as the pipeline keeps them): pruned sets of 795 k / 211 k / 54 k /
13.8 k rows and the 8-child expansions of 110 k / 435 k / 1.69 M rows.  Prints one line per (set, shape):
    rows pairs | h2 (mma.sync) ms | wide (tcgen05) ms | speed-up | algorithmic GB/s | err of wide vs h2 result
Run on the GPU box:  python tools/bench_wide.py [--reps 20]"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pcgcv2_b200 import ops, synth  # noqa: E402

SHAPES = {  # set name -> shapes run there in the network (autoencoder.py channel plan)
    "N1 211k": [(32, 32), (32, 8), (8, 16), (8, 8)],
    "N2 54k": [(64, 64), (64, 16), (16, 32), (16, 16)],
    "N3 13.8k": [(32, 8), (8, 16), (8, 8)],
    "8N3 110k": [(64, 64), (64, 16), (16, 32), (16, 16), (64, 1)],
    "8N2 435k": [(32, 32), (32, 8), (8, 16), (8, 8), (32, 1)],
    "8N1 1.69M": [(16, 16), (16, 4), (16, 1)],
}


def timeit(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--only", default="", help='e.g. "8N3 110k:64x64" -- one set and shape, tcgen05 kernel only (for ncu captures)')
    args = ap.parse_args()
    only_set, only_shape = (args.only.split(":") + [""])[:2] if args.only else ("", "")
    dev = torch.device("cuda")
    pts = synth.synthetic_vox10(0)
    keys0, _ = ops.argsort_u64(ops.pack_keys(torch.nn.functional.pad(torch.from_numpy(pts).to(dev), (1, 0)), 1))
    sets = {"N0 795k": keys0}
    k = keys0
    for name in ("N1 211k", "N2 54k", "N3 13.8k"):
        k = ops.stride_down(k, keys_are_sorted=True)[0].contiguous()
        sets[name] = k
    sets["8N3 110k"] = ops.upsample_keys(sets["N3 13.8k"])
    sets["8N2 435k"] = ops.upsample_keys(sets["N2 54k"])
    sets["8N1 1.69M"] = ops.upsample_keys(sets["N1 211k"])
    g = torch.Generator().manual_seed(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for name, shapes in SHAPES.items():
        if only_set and name != only_set:
            continue
        keys = sets[name]
        n = keys.shape[0]
        nbr, npairs = ops.kernel_map_k3(keys, ops.HashTable(keys), count_pairs=True)
        pairs = int(npairs.item())
        print(f"== {name}: rows {n} pairs {pairs}", flush=True)
        for cin, cout in shapes:
            if only_shape and f"{cin}x{cout}" != only_shape:
                continue
            x = (torch.randn(n, cin, generator=g) * 2).to(dev)
            w = (torch.randn(27, cin, cout, generator=g) / np.sqrt(27 * cin)).to(dev)
            b = torch.randn(1, cout, generator=g).to(dev)
            xh = ops.split_h2(x)
            pw = ops.PackedK3Wide(w)
            want_h = cout % 4 == 0
            t_wide = timeit(lambda: ops.conv_k3_wide(xh, nbr, pw, b, relu=True, want_h2=want_h), args.reps)
            if args.only:
                print(f"  {cin}->{cout} tcgen05 {t_wide:.4f} ms", flush=True)
                continue
            y_wide = ops.conv_k3_wide(xh, nbr, pw, b, relu=True)[0]
            t_h2, y_h2 = None, None
            if cin % 16 == 0 and ops.PackedK3H2.supported(cin, cout):
                ph = ops.PackedK3H2(w)
                t_h2 = timeit(lambda: ops.conv_k3_h2(xh, nbr, ph, b, relu=True, want_h2=want_h), args.reps)
                y_h2 = ops.conv_k3_h2(xh, nbr, ph, b, relu=True)[0]
            elif cin % 16 == 0 and cout % 16 == 0 and ops.PackedK3H2.supported(cin, 16):      # cout = 64: four 16-wide slices
                sl = [(ops.PackedK3H2(w[:, :, j:j + 16].contiguous()), b[:, j:j + 16].contiguous()) for j in range(0, cout, 16)]
                out = torch.empty((n, cout), device=dev)
                outh = torch.empty((n, cout), dtype=torch.int32, device=dev)

                def run():
                    for j, (ps, bb) in enumerate(sl):
                        ops.conv_k3_h2(xh, nbr, ps, bb, relu=True, out=out[:, 16 * j:16 * j + 16], out_h2=outh[:, 16 * j:16 * j + 16])
                t_h2 = timeit(run, args.reps)
                run()
                y_h2 = out
            else:                                                                               # cin = 8: 3xTF32 mma.sync
                pk = ops.PackedK3(w)
                if pk.packed is not None:
                    t_h2 = timeit(lambda: ops.conv_k3_packed(x, nbr, pk, b, relu=True), args.reps)
                    y_h2 = ops.conv_k3_packed(x, nbr, pk, b, relu=True)
            alg = 4 * (n * cin + n * cout) + 8 * pairs + 4 * 27 * cin * cout
            err = float((y_wide - y_h2).abs().max() / y_h2.abs().max()) if y_h2 is not None else float("nan")
            ref_ms = f"{t_h2:.4f}" if t_h2 else "   n/a"
            sp = f"{t_h2 / t_wide:.2f}x" if t_h2 else "  n/a"
            print(f"  {cin:>2}->{cout:<2}  mma.sync {ref_ms} ms   tcgen05 {t_wide:.4f} ms   {sp}   {alg / t_wide / 1e6:7.0f} GB/s alg   "
                  f"{2 * pairs * cin * cout / t_wide / 1e9:6.1f} TFLOP/s   diff {err:.1e}", flush=True)
        del nbr
    del flush


if __name__ == "__main__":
    main()
