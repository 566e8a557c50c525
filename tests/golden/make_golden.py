"""Generate the committed fixtures under tests/golden/.

Run in the BUILD container only (needs /root/reference):
    python tests/golden/make_golden.py

Writes
  ckpt_r3.npz, ckpt_r7.npz     -- the reference's shipped checkpoints
                                  (ckpts.zip: r3_0.10bpp.pth / r7_0.4bpp.pth), float32
                                  arrays keyed by state_dict name (fixture data, MIT).
  entropy_r3.npz, entropy_r7.npz -- outputs of the REFERENCE's own
                                  entropy_model.py (imported from /root/reference with a
                                  stub ``torchac``) on those weights: this is what pins
                                  oracle/entropy_ref.py and the CUDA entropy kernels.
  oracle_cube32_r3.npz         -- oracle outputs for config 1 (32^3 random cube, r3).
  oracle_vox8_kat.json         -- checkpoint-behaviour known answers.
"""
import io
import json
import os
import sys
import types
import zipfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from oracle import codec_ref, entropy_ref, metrics_ref, rangecoder_ref  # noqa: E402
from pcgcv2_b200 import synth  # noqa: E402


def load_ckpt(name):
    with zipfile.ZipFile(os.path.join(REF, "ckpts.zip")) as z:
        return torch.load(io.BytesIO(z.read(f"ckpts/{name}")), map_location="cpu")["model"]


def save_ckpt(sd, out):
    arrs = {k: v.detach().float().numpy() for k, v in sd.items()
            if not k.startswith("entropy_bottleneck.") or "._" in k}       # drop the 3 alias entries
    np.savez_compressed(out, **arrs)


def reference_entropy_model(sd):
    """Import the reference's entropy_model.py unchanged (stub torchac)."""
    sys.modules.setdefault("torchac", types.ModuleType("torchac"))
    sys.path.insert(0, REF)
    from entropy_model import EntropyBottleneck
    sys.path.remove(REF)
    eb = EntropyBottleneck(8)
    eb.load_state_dict({k[len("entropy_bottleneck."):]: v for k, v in sd.items()
                        if k.startswith("entropy_bottleneck.")})
    return eb.eval()


def entropy_golden(sd, out):
    eb = reference_entropy_model(sd)
    rng = np.random.default_rng(1)
    vals = torch.from_numpy(np.concatenate([
        rng.normal(scale=4.0, size=(512, 8)), np.round(rng.normal(scale=6.0, size=(512, 8))),
        np.linspace(-30, 30, 8 * 61).reshape(61, 8)]).astype(np.float32))
    with torch.no_grad():
        lik = eb._likelihood(vals).contiguous().numpy()
        tables = {}
        for (lo, hi) in [(-3, 1), (-10, 8), (-12, 12), (0, 0), (-40, 37)]:
            symbols = torch.arange(float(lo), float(hi) + 1).reshape(-1, 1).repeat(1, 8)
            pmf = torch.clamp(eb._likelihood(symbols), min=eb._likelihood_bound).permute(1, 0)
            tables[f"cdf_{lo}_{hi}"] = eb._pmf_to_cdf(pmf).numpy()
    np.savez_compressed(out, values=vals.numpy(), likelihood=lik, **tables)


def cube_golden(sd, out):
    torch.set_flush_denormal(True)
    pts = synth.random_cube(0, 32, 0.1)
    coords = np.concatenate([np.zeros((len(pts), 1), np.int32), pts], 1)
    rec = {}
    st = codec_ref.encode(sd, coords, rec)
    dec, cls_list = codec_ref.decode(sd, st)
    digest = {}
    for k, v in rec.items():
        if isinstance(v, torch.Tensor):
            digest[k] = np.array([v.shape[0], v.shape[1], float(v.double().sum()), float(v.double().abs().sum()),
                                  float(v.abs().max())], dtype=np.float64)
    np.savez_compressed(out, y_C=st["y_C"], y_F=st["y_F"].numpy(), F=np.frombuffer(st["F"], np.uint8),
                        H=np.frombuffer(st["H"], np.uint8), num_points=np.frombuffer(st["num_points"], np.uint8),
                        dec_C=dec, **{"digest/" + k: v for k, v in digest.items()})


def vox8_kat(sds, out):
    torch.set_flush_denormal(True)
    pts = synth.ellipsoid_vox8()
    coords = np.concatenate([np.zeros((len(pts), 1), np.int32), pts], 1)
    kat = {"generator": "pcgcv2_b200.synth.ellipsoid_vox8(seed=0)", "N0": int(len(pts))}
    for name, sd in sds.items():
        st = codec_ref.encode(sd, coords)
        dec, _ = codec_ref.decode(sd, st)
        kat[name] = {"N3": int(len(st["C_coords"])), "ideal_bits": round(st["ideal_bits"], 1),
                     "F_bytes": len(st["F"]), "N_out": int(len(dec)),
                     "D1_psnr": round(metrics_ref.d1_psnr(pts, dec[:, 1:], 256), 4),
                     "sym_min": float(np.frombuffer(st["H"][9:13], np.float32)[0]),
                     "sym_max": float(np.frombuffer(st["H"][13:17], np.float32)[0])}
    with open(out, "w") as f:
        json.dump(kat, f, indent=1)
    print(json.dumps(kat, indent=1))


if __name__ == "__main__":
    sds = {"r3": load_ckpt("r3_0.10bpp.pth"), "r7": load_ckpt("r7_0.4bpp.pth")}
    for name, sd in sds.items():
        save_ckpt(sd, os.path.join(HERE, f"ckpt_{name}.npz"))
        entropy_golden(sd, os.path.join(HERE, f"entropy_{name}.npz"))
    cube_golden(sds["r3"], os.path.join(HERE, "oracle_cube32_r3.npz"))
    vox8_kat(sds, os.path.join(HERE, "oracle_vox8_kat.json"))
