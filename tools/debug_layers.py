"""debug helper: per-layer error of the CUDA codec vs the oracle on the 32^3 cube."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from oracle import codec_ref
from pcgcv2_b200 import ops, synth
from pcgcv2_b200.codec import Codec
from util import load_ckpt, with_batch
torch.set_flush_denormal(True)
sd = load_ckpt("r3")
coords = with_batch(synth.random_cube(0, 32, 0.1))
rec_ref = {}
st_ref = codec_ref.encode(sd, coords, rec_ref)
codec_ref.decode(sd, st_ref, record=rec_ref)
for tc in (False, True):
    codec = Codec(sd, use_tensor_cores=tc)
    codec.record = {}
    st = codec.encode(coords[:, 1:]); codec.decode(st)
    print("tensor cores", tc)
    for name, (t, keys, stride) in codec.record.items():
        if name not in rec_ref: continue
        ref = rec_ref[name]
        relu_names = [f"{a}{i}" for a in ("encoder.down", "decoder.up", "encoder.conv", "decoder.conv") for i in range(3)]
        if name.endswith((".conv0_0", ".conv1_1")) or name in relu_names:
            ref = torch.relu(ref)
        rc = rec_ref[name + ".C"]
        o1 = np.lexsort(np.asarray(rc).T[::-1]); gc = ops.unpack_keys(keys, stride).cpu().numpy(); o2 = np.lexsort(gc.T[::-1])
        d = (t.cpu()[torch.from_numpy(o2)] - ref[torch.from_numpy(o1)]).abs().max().item()
        m = ref.abs().max().item()
        if d / max(m, 1e-30) > 5e-6:
            print(f"  {name:34s} max|ref| {m:.3e}  max|d| {d:.3e}  rel {d/max(m,1e-30):.2e}")
