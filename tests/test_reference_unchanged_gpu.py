"""The reference's own files -- coder.py, pcc_model.py, autoencoder.py, entropy_model.py, gpcc.py, pc_error.py, byte for
byte as shipped (installed to oracle/_ref by __graft_entry__.build(), never committed) -- running on the GPU over the
drop-in MinkowskiEngine / torchac / data_utils shims: ``Coder.encode`` / ``Coder.decode`` (coder.py:80-112) with real
files, the real tmc3 subprocess and the real pc_error_d metric, i.e. the flow of ``python coder.py`` (coder.py:114-184).
Checked against the CPU oracle (bit-exact files) and against ``Codec`` (same decoded set)."""
import os

import numpy as np
import pytest
import torch

from oracle import codec_ref, refbin
from pcgcv2_b200 import ops, synth
from pcgcv2_b200.codec import Codec
from pcgcv2_b200.coords_coder import Tmc3CoordinateCoder
from util import canon, load_ckpt, with_batch

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not refbin.reference_sources_installed(), reason="reference not installed under oracle/_ref")]


def _run_reference(sd, pts, tmp_path, res, rho=1.0, scaling_factor=1.0):
    coder = refbin.load_reference_coder()
    ply = str(tmp_path / "input_cloud.ply")                                # (CoordinateCoder deletes <prefix>.ply, coder.py:27,34)
    ops.ply_write_ascii(ply, pts)
    x = coder.load_sparse_tensor(ply, coder.device)                        # data_utils.load_sparse_tensor (shim)
    model = coder.PCCModel().to(coder.device)
    model.load_state_dict(refbin.reference_state_dict(sd))                 # strict, coder.py:142
    os.makedirs(str(tmp_path / "output"), exist_ok=True)                   # coder.py:131-134
    prefix = str(tmp_path / "output" / "cloud")
    c = coder.Coder(model=model, filename=prefix)
    x_in = coder.scale_sparse_tensor(x, factor=scaling_factor) if scaling_factor != 1 else x
    c.encode(x_in)
    x_dec = c.decode(rho=rho)
    if scaling_factor != 1:
        x_dec = coder.scale_sparse_tensor(x_dec, factor=1.0 / scaling_factor)
    files = {p: open(prefix + p, "rb").read() for p in ("_C.bin", "_F.bin", "_H.bin", "_num_points.bin")}
    dec = x_dec.C.detach().cpu().numpy()[:, 1:]
    coder.write_ply_ascii_geo(prefix + "_dec.ply", dec)
    metrics = coder.pc_error(ply, prefix + "_dec.ply", res=res, show=False)          # coder.py:181-184
    assert "mseF,PSNR (p2point)" in metrics, (
        f"pc_error printed no D1 line for {len(dec)} decoded points; columns {list(metrics.columns)}; "
        + __import__("subprocess").run([refbin.PC_ERROR, "-a", ply, "-b", prefix + "_dec.ply", "--hausdorff=1", f"--resolution={res - 1}"],
                                       capture_output=True, text=True).stdout[-600:])
    return files, dec, float(metrics["mseF,PSNR (p2point)"][0]), len(x)


def test_unchanged_reference_coder_on_gpu_matches_oracle_and_codec(tmp_path):
    torch.set_flush_denormal(True)
    sd = load_ckpt("r3")
    pts = synth.ellipsoid_vox8()                                           # 91 568 voxels, res 256 (the KAT cloud)
    files, dec, d1, n = _run_reference(sd, pts, tmp_path, res=256)
    assert n == len(pts)
    ref = codec_ref.encode(sd, with_batch(pts))
    ref_dec, _ = codec_ref.decode(sd, ref)
    c_ref = refbin.gpcc_encode_coords(ref["C_coords"], str(tmp_path))
    assert files["_num_points.bin"] == ref["num_points"] and files["_H.bin"] == ref["H"]
    assert files["_F.bin"] == ref["F"], "feature stream written by the reference over the shim != oracle's"
    assert files["_C.bin"] == c_ref, "tmc3 stream differs"
    bits = 8 * sum(len(v) for v in files.values())
    ref_bits = codec_ref.stream_bits(ref) + 8 * len(c_ref)
    assert abs(bits - ref_bits) <= 1e-4 * ref_bits
    assert (canon(dec) == canon(ref_dec[:, 1:])).all(), "decoded set != oracle"
    # the bench's pipeline gives the same stream and the same decoded set
    codec = Codec(sd, coords_coder=Tmc3CoordinateCoder(refbin.TMC3))
    st = codec.encode(pts)
    assert st.F == files["_F.bin"] and st.H == files["_H.bin"] and st.num_points == files["_num_points.bin"]
    assert st.C == files["_C.bin"] and st.bits() == bits
    assert (canon(codec.decode(st)) == canon(dec)).all()
    with __import__("tempfile").TemporaryDirectory() as t:
        assert abs(d1 - refbin.pc_error_d1(pts, ref_dec[:, 1:], 256, t)) < 0.01
    assert abs(d1 - 62.5504) < 0.01                                        # SURVEY Appendix E.7 known answer for r3


def test_unchanged_reference_scaled_path_config4(tmp_path):
    """scaling_factor 0.375 / rho 4 through the reference's own scale_sparse_tensor call sites (coder.py:149-152,166-167)."""
    torch.set_flush_denormal(True)
    sd = load_ckpt("r3")
    pts = synth.ellipsoid_vox8()
    files, dec, d1, _ = _run_reference(sd, pts, tmp_path, res=256, rho=4.0, scaling_factor=0.375)
    x_in = codec_ref.scale_coords(pts, 0.375)
    ref = codec_ref.encode(sd, with_batch(x_in))
    ref_dec, _ = codec_ref.decode(sd, ref, rho=4.0)
    ref_dec = codec_ref.scale_coords(ref_dec[:, 1:], 1.0 / 0.375)
    assert files["_F.bin"] == ref["F"] and files["_H.bin"] == ref["H"] and files["_num_points.bin"] == ref["num_points"]
    assert (canon(dec) == canon(ref_dec)).all()
    codec = Codec(sd)
    mine = codec.scale(codec.decode(codec.encode(codec.scale(pts, 0.375)), rho=4.0, to_host=False), 1.0 / 0.375).cpu().numpy()
    assert (canon(mine) == canon(dec)).all()
