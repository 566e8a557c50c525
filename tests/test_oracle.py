"""Pins for the CPU oracle (runs without a GPU)."""
import json
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch

from oracle import codec_ref, entropy_ref, metrics_ref, rangecoder_ref as rc, sparse_ref as S
from pcgcv2_b200 import synth
from util import GOLDEN, canon, load_ckpt, with_batch


@pytest.mark.parametrize("name", ["r3", "r7"])
def test_entropy_ref_matches_reference_module(name):
    """oracle/entropy_ref.py vs outputs frozen from the reference's own entropy_model.py."""
    g = np.load(os.path.join(GOLDEN, f"entropy_{name}.npz"))
    params = entropy_ref.params_from_state_dict(load_ckpt(name))
    lik = entropy_ref.likelihood(params, torch.from_numpy(g["values"])).numpy()
    np.testing.assert_allclose(lik, g["likelihood"], rtol=1e-6, atol=1e-9)
    for key in g.files:
        if key.startswith("cdf_"):
            _, lo, hi = key.split("_")
            cdf = entropy_ref.cdf_table(params, float(lo), float(hi), 8).numpy()
            np.testing.assert_allclose(cdf, g[key], rtol=0, atol=1e-7)
            assert (rc.cdf_float_to_u16(cdf) == rc.cdf_float_to_u16(g[key])).all()


def _random_stream(seed, n, C, L):
    rng = np.random.default_rng(seed)
    pmf = rng.random((C, L)).astype(np.float32) ** 3 + 1e-4
    pmf /= pmf.sum(1, keepdims=True)
    cdf = np.concatenate([np.zeros((C, 1), np.float32), np.cumsum(pmf, 1)], 1).clip(max=1).astype(np.float32)
    sym = np.stack([rng.choice(L, size=n, p=pmf[c] / pmf[c].sum()) for c in range(C)], 1).astype(np.int16)
    return cdf, sym, pmf


@pytest.mark.parametrize("n,C,L", [(1, 1, 1), (1, 8, 5), (300, 8, 5), (257, 3, 19), (64, 8, 1), (2000, 8, 78)])
def test_rangecoder_roundtrip_and_py_vs_c(n, C, L):
    cdf, sym, pmf = _random_stream(n * 7 + L, n, C, L)
    tiled = np.broadcast_to(cdf, (n, C, L + 1)).copy()
    data = rc.encode_float_cdf(tiled, sym, check_input_bounds=True)
    assert (rc.decode_float_cdf(tiled, data) == sym).all()
    table = rc.cdf_float_to_u16(cdf)
    rows = np.tile(np.arange(C, dtype=np.int32), n)
    if n <= 300:
        assert rc.py_encode(table, rows, sym.reshape(-1)) == data
        assert (rc.py_decode(table, rows, data).reshape(n, C) == sym).all()
    # table quantisation (16 bit + the +arange normalisation) bounds the excess over the ideal length
    q = np.diff(np.concatenate([table[:, :-1].astype(np.int64), np.full((C, 1), 65536)], 1), axis=1) / 65536.0
    ideal_q = -np.log2(q[np.arange(C)[None, :].repeat(n, 0), sym]).sum()
    assert 8 * len(data) <= ideal_q + 16 + 8


def test_rangecoder_empty_and_per_symbol_rows():
    cdf, sym, _ = _random_stream(3, 40, 4, 6)
    per_sym = np.broadcast_to(cdf, (40, 4, 7)).copy()
    per_sym[::2] = np.linspace(0, 1, 7, dtype=np.float32)             # genuinely different rows
    data = rc.encode_float_cdf(per_sym, sym)
    assert (rc.decode_float_cdf(per_sym, data) == sym).all()
    with pytest.raises(ValueError):
        rc.encode_float_cdf(per_sym, sym + 6, check_input_bounds=True)


def test_u16_table_shape_matches_survey_example():
    """Appendix B.1 / E.4: r3 channel 0 over symbols -12..12 is [0,1,..,12,65524,..,65535,(65536->0)]."""
    params = entropy_ref.params_from_state_dict(load_ckpt("r3"))
    t = rc.cdf_float_to_u16(entropy_ref.cdf_table(params, -12, 12, 8).numpy())
    assert t[0].tolist() == list(range(0, 13)) + list(range(65524, 65536)) + [0]


def test_kernel_map_bruteforce():
    rng = np.random.default_rng(5)
    for stride in (1, 2, 4):
        pts = np.unique(rng.integers(0, 9, size=(300, 3)), axis=0) * stride
        coords = with_batch(pts)
        coords[::3, 0] = 1                                           # two batch items
        nbr = S.kernel_map_k3(coords, stride)
        d = {tuple(c): i for i, c in enumerate(coords.tolist())}
        for u, c in enumerate(coords.tolist()):
            for k in range(27):
                off = ((k % 3) - 1, ((k // 3) % 3) - 1, (k // 9) - 1)
                q = (c[0], c[1] + off[0] * stride, c[2] + off[1] * stride, c[3] + off[2] * stride)
                assert nbr[u, k] == d.get(q, -1)


def test_stride_down_and_transpose_conventions():
    coords = with_batch([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [3, 2, 5], [-1, 0, 0]])
    out, parent, kidx = S.stride_down(coords, 1)
    assert kidx.tolist() == [0, 1, 2, 4, 1 + 0 + 4, 1]
    assert out[parent].tolist() == [[0, 0, 0, 0]] * 4 + [[0, 2, 2, 4], [0, -2, 0, 0]]
    f = torch.arange(2 * 3, dtype=torch.float32).reshape(2, 3)
    w = torch.randn(8, 3, 2)
    o, oc = S.convT_k2s2(f, with_batch([[0, 0, 0], [8, 0, 4]]), 4, w, None)
    assert oc[8 + 5].tolist() == [0, 8 + 2, 0, 4 + 2]               # k=5 -> (+x, 0, +z)
    torch.testing.assert_close(o[8 + 5], f[1] @ w[5])


def test_conv_k3_matches_dense_conv3d():
    """sparse k3 conv == dense cross-correlation on the occupied voxels (Appendix A.3)."""
    rng = np.random.default_rng(2)
    occ = rng.random((6, 7, 5)) < 0.4
    pts = np.argwhere(occ)
    coords = with_batch(pts)
    Cin, Cout = 3, 4
    f = torch.randn(len(pts), Cin)
    w = torch.randn(27, Cin, Cout)
    b = torch.randn(1, Cout)
    out = S.conv_k3(f, coords, 1, w, b)
    dense = torch.zeros(1, Cin, 5, 7, 6)                            # (z, y, x)
    dense[0, :, pts[:, 2], pts[:, 1], pts[:, 0]] = f.t()
    wd = w.reshape(3, 3, 3, Cin, Cout).permute(4, 3, 0, 1, 2)       # k = ix + 3iy + 9iz -> [iz, iy, ix]
    ref = torch.nn.functional.conv3d(dense, wd, bias=b.reshape(-1), padding=1)
    torch.testing.assert_close(out, ref[0, :, pts[:, 2], pts[:, 1], pts[:, 0]].t(), rtol=1e-4, atol=1e-4)


def test_cube32_regression():
    """config 1 (32^3 random cube, r3): oracle output is frozen."""
    torch.set_flush_denormal(True)
    g = np.load(os.path.join(GOLDEN, "oracle_cube32_r3.npz"))
    sd = load_ckpt("r3")
    coords = with_batch(synth.random_cube(0, 32, 0.1))
    assert len(coords) == 3339
    rec = {}
    st = codec_ref.encode(sd, coords, rec)
    assert (st["y_C"] == g["y_C"]).all()
    np.testing.assert_allclose(st["y_F"].numpy(), g["y_F"], rtol=1e-4, atol=1e-5)
    assert st["F"] == g["F"].tobytes() and st["H"] == g["H"].tobytes()
    dec, _ = codec_ref.decode(sd, st)
    assert (canon(dec) == canon(g["dec_C"])).all()
    for k in g.files:
        if k.startswith("digest/"):
            v = rec[k[len("digest/"):]]
            d = g[k]
            assert tuple(v.shape) == (int(d[0]), int(d[1]))
            np.testing.assert_allclose(float(v.double().abs().sum()), d[3], rtol=1e-4)


@pytest.mark.slow
def test_vox8_checkpoint_behaviour_kat():
    """SURVEY.md Appendix E.7/E.8 operating point (r3): the conventions in sparse_ref are the
    ones the shipped weights were trained with (any deviation costs 4-19 dB)."""
    torch.set_flush_denormal(True)
    kat = json.load(open(os.path.join(GOLDEN, "oracle_vox8_kat.json")))
    pts = synth.ellipsoid_vox8()
    assert len(pts) == kat["N0"] == 91568
    sd = load_ckpt("r3")
    st = codec_ref.encode(sd, with_batch(pts))
    assert len(st["C_coords"]) == kat["r3"]["N3"] == 1521
    assert abs(st["ideal_bits"] - 4623) < 2 and len(st["F"]) == kat["r3"]["F_bytes"] == 575
    dec, _ = codec_ref.decode(sd, st)
    assert abs(metrics_ref.d1_psnr(pts, dec[:, 1:], 256) - 62.5504) < 0.01


@pytest.mark.skipif(not os.path.exists("/root/reference/pc_error_d"), reason="reference binaries not present")
def test_d1_matches_pc_error_binary(tmp_path):
    exe = str(tmp_path / "pc_error_d")
    shutil.copy("/root/reference/pc_error_d", exe)
    os.chmod(exe, 0o755)
    rng = np.random.default_rng(0)
    a = np.unique(rng.integers(0, 64, size=(4000, 3)), axis=0)
    b = np.unique(np.clip(a + rng.integers(-1, 2, size=a.shape) * (rng.random(a.shape) < 0.3), 0, 63), axis=0)

    def write(path, pts):
        with open(path, "w") as f:
            f.write("ply\nformat ascii 1.0\nelement vertex %d\nproperty float x\nproperty float y\n"
                    "property float z\nend_header\n" % len(pts))
            np.savetxt(f, pts, fmt="%d")
    write(tmp_path / "a.ply", a)
    write(tmp_path / "b.ply", b)
    out = subprocess.run([exe, "-a", str(tmp_path / "a.ply"), "-b", str(tmp_path / "b.ply"), "--hausdorff=1",
                          "--resolution=63"], capture_output=True, text=True).stdout
    line = [l for l in out.splitlines() if "mseF,PSNR (p2point)" in l][0]
    assert abs(float(line.split(":")[-1]) - metrics_ref.d1_psnr(a, b, 64)) < 1e-3
