"""CPU emulation of the pre-split half-precision (h2) arithmetic of csrc/conv_h2.cuh over the WHOLE network, on the
oracle: every k=3 layer with >= 8 input channels computes hi*Whi + lo*Whi + hi*Wlo with x = f16 hi + f16 lo and the
weights split after a power-of-two scale (the lo*Wlo term is dropped, as in the kernels).  Prints max|delta|/max|ref|
per layer against the fp32 oracle and whether the bitstream / decoded set change.  No GPU needed.

    python tools/h2_emulation.py [r3|r7] [vox8|cube32]
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from oracle import codec_ref, sparse_ref as S
from pcgcv2_b200 import synth
from util import load_ckpt


def split16(x, scale=1.0):
    xs = x * scale
    hi = xs.clamp(-65504, 65504).half()
    lo = (xs - hi.float()).clamp(-65504, 65504).half()
    return hi.float(), lo.float()


def run(ckpt="r3", cloud="vox8", verbose=True):
    torch.set_flush_denormal(True)
    orig = S.conv_from_map
    mode = {"on": False}

    def conv_split(feats, nbr, weight, bias, n_out=None):
        if not mode["on"] or weight.shape[1] < 8:
            return orig(feats, nbr, weight, bias, n_out)
        wmax = float(weight.abs().max())
        sw = 2.0 ** int(np.floor(np.log2(16384.0 / wmax))) if wmax > 0 else 1.0
        xh, xl = split16(feats)
        wh, wl = split16(weight, sw)
        out = (orig(xh, nbr, wh, None, n_out) + orig(xl, nbr, wh, None, n_out) + orig(xh, nbr, wl, None, n_out)) / sw
        return out if bias is None else out + bias.reshape(1, -1)

    S.conv_from_map = conv_split
    try:
        sd = load_ckpt(ckpt)
        pts = synth.ellipsoid_vox8(0) if cloud == "vox8" else synth.random_cube(0, 32, 0.1)
        coords = np.concatenate([np.zeros((len(pts), 1), np.int32), pts.astype(np.int32)], 1)
        res = {}
        for on in (False, True):
            mode["on"] = on
            rec = {}
            st = codec_ref.encode(sd, coords, rec)
            out, _ = codec_ref.decode(sd, st, 1.0, rec)
            res[on] = (rec, st, out)
    finally:
        S.conv_from_map = orig
    (ref, st0, out0), (new, st1, out1) = res[False], res[True]
    worst, peak = 0.0, 0.0
    for k in ref:
        if k.endswith(".C") or ref[k].shape != new[k].shape:
            continue
        m = float(ref[k].abs().max()) or 1e-30
        err = float((ref[k] - new[k]).abs().max()) / m
        worst, peak = max(worst, err), max(peak, m)
        if verbose:
            print(f"{k:32s} max|ref| {m:10.4f}  rel err {err:9.2e}")
    same_stream = st0["F"] == st1["F"]
    same_set = out0.shape == out1.shape and bool(np.array_equal(out0, out1))
    if verbose:
        print(f"worst relative error {worst:.2e}; largest activation {peak:.1f}; bitstream identical {same_stream}; "
              f"decoded set identical {same_set}")
    return worst, peak, same_stream, same_set


if __name__ == "__main__":
    run(*(sys.argv[1:3]))
