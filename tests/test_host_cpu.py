"""Host-side pieces of the path that need no GPU: the coordinate coders (row f1) and the native PLY I/O (row f2)."""
import os

import numpy as np
import pytest
import torch

from oracle import refbin
from pcgcv2_b200 import ops, synth
from pcgcv2_b200.coords_coder import OctreeCoordinateCoder, Tmc3CoordinateCoder
from util import GOLDEN, canon


def test_octree_coder_round_trips_and_stays_near_tmc3():
    g = np.load(os.path.join(GOLDEN, "oracle_vox10_seed0.npz"))
    c3 = g["C_coords"].astype(np.int32)
    oc = OctreeCoordinateCoder()
    blob = oc.encode(c3)
    assert (canon(oc.decode(blob)) == canon(c3)).all() and len(blob) < 1.6 * int(g["C_bytes"])
    for n in (0, 1, 2, 7):                                                  # empty and tiny sets
        assert (canon(oc.decode(oc.encode(c3[:n]))) == canon(c3[:n])).all()
    assert oc.decode(oc.encode(np.zeros((1, 3), np.int32))).tolist() == [[0, 0, 0]]
    rng = np.random.default_rng(0)
    p = rng.integers(0, 1 << 20, size=(4000, 3)).astype(np.int32)          # deep, scattered, with duplicates
    p = np.concatenate([p, p[:100]])
    assert (canon(oc.decode(oc.encode(p))) == canon(np.unique(p, axis=0))).all()
    with pytest.raises(Exception):
        oc.encode(np.array([[-1, 0, 0]], np.int32))
    with pytest.raises(Exception):
        oc.decode(b"nonsense-bytes")
    bad = bytearray(blob)
    bad[20] ^= 0x55                                                          # corruption is detected or yields another set, never a crash
    try:
        oc.decode(bytes(bad))
    except Exception:
        pass


@pytest.mark.skipif(not refbin.available(), reason="reference binaries not installed (oracle/_ref)")
def test_tmc3_coder_is_the_reference_command_line():
    g = np.load(os.path.join(GOLDEN, "oracle_vox10_seed0.npz"))
    c3 = g["C_coords"].astype(np.int32)
    tc = Tmc3CoordinateCoder(refbin.TMC3)
    blob = tc.encode(c3)
    assert len(blob) == int(g["C_bytes"]) and (canon(tc.decode(blob)) == canon(c3)).all()
    with pytest.raises(FileNotFoundError):
        Tmc3CoordinateCoder("/nonexistent/tmc3")


def test_ply_io_native_matches_reference_semantics(tmp_path):
    """f2: native ASCII PLY writer / reader (csrc/ply.cpp) against the reference's line semantics (data_utils.py:19-48)."""
    pts = synth.ellipsoid_vox8()
    f = str(tmp_path / "a.ply")
    ops.ply_write_ascii(f, pts)
    assert (refbin.read_ply(f) == pts).all()                               # what the reference's reader sees
    pin = torch.cuda.is_available()                                         # pinned staging needs a CUDA context
    back = ops.ply_read_ascii(f, pinned=pin)
    assert back.is_pinned() == pin and back.dtype == torch.int32 and (back.numpy() == pts).all()
    refbin.write_ply(f, pts[:1000])                                        # the reference writer's output
    assert (ops.ply_read_ascii(f).numpy() == pts[:1000]).all()
    odd = str(tmp_path / "b.ply")
    with open(odd, "w") as fh:
        fh.write("ply\nformat ascii 1.0\ncomment 1 2 3\nelement vertex 4\nproperty float x\nend_header\n"
                 "1 2 3\n4.7 -5.2 6e0 9 9\n7  8 9\n-1 -2 -3 \n10 11 12")
    assert ops.ply_read_ascii(odd).numpy().tolist() == [[1, 2, 3], [4, -5, 6], [-1, -2, -3], [10, 11, 12]]


@pytest.mark.parametrize("name", ["r3", "r7"])
def test_host_cdf_table_is_exactly_the_references(name):
    """the table Codec codes with (entropy_host.HostTable: the reference's float32 CPU operator sequence) == the table the
    reference's own entropy_model.py produced (tests/golden/make_golden.py), bit for bit in float32 and in torchac's uint16
    conversion -- the precondition for _F.bin being decodable on either side."""
    from oracle import entropy_ref, rangecoder_ref
    from pcgcv2_b200.entropy_host import HostTable
    from util import load_ckpt
    gold = np.load(os.path.join(GOLDEN, f"entropy_{name}.npz"))
    p = entropy_ref.params_from_state_dict(load_ckpt(name))
    host = HostTable(p["matrices"], p["biases"], p["factors"])
    checked = 0
    for key in gold.files:
        if not key.startswith("cdf_"):
            continue
        _, lo, hi = key.split("_")
        cdf = host.cdf_float(int(lo), int(hi))
        assert cdf.dtype == np.float32 and np.array_equal(cdf, gold[key]), key
        lp = cdf.shape[1]
        u16 = (np.round(cdf * np.float32(65536 - (lp - 1))).astype(np.int64) + np.arange(lp)).astype(np.uint16)
        assert np.array_equal(u16, rangecoder_ref.cdf_float_to_u16(gold[key]).astype(np.uint16)), key
        checked += 1
    assert checked >= 1


def test_range_coder_from_symbol_intervals_is_the_same_stream():
    """pcgc_rc_encode_ranges_host (intervals as pcgc_symbol_ranges writes them) == the table-walking coder == the oracle."""
    from oracle import rangecoder_ref
    rng = np.random.default_rng(5)
    C, L = 8, 19
    pmf = rng.random((C, L)) ** 4 + 1e-9
    pmf /= pmf.sum(1, keepdims=True)
    cdf = np.concatenate([np.zeros((C, 1)), np.cumsum(pmf, 1)], 1).clip(max=1).astype(np.float32)
    tab = rangecoder_ref.cdf_float_to_u16(cdf).astype(np.uint16)
    sym = rng.integers(0, L, size=(3000, C)).astype(np.int16)
    sym[:40] = L - 1                                                       # the last symbol: c_high = 0x10000
    flat, rows = sym.reshape(-1), np.arange(sym.size) % C
    lo = tab[rows, flat].astype(np.uint32)
    hi = np.where(flat == L - 1, 65536, tab[rows, np.minimum(flat + 1, L)]).astype(np.uint32)
    got = ops.rc_encode_ranges(lo | ((hi - 1) << 16))
    assert got == ops.rc_encode_u16(tab, sym) == rangecoder_ref.encode_u16(tab, rows.astype(np.int32), flat)
    assert ops.rc_encode_ranges(np.zeros(0, np.uint32)) == ops.rc_encode_u16(tab, np.zeros((0, C), np.int16))


def test_bench_algorithmic_bytes_match_the_survey():
    """bench.py's section-8(d) byte model on SURVEY Appendix C's level sizes gives the survey's totals: 8.57 GB over the 106
    convolutions of one encode + decode, 474.4 MB for decoder.conv2 (the roofline kernel)."""
    import importlib.util
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_for_test", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        spec.loader.exec_module(bench)
    finally:
        sys.argv = argv
    n = {"L0": 795124, "L1": 211346, "L2": 54369, "L3": 13784, "U2": 110272, "U1": 434952, "U0": 1690768}
    p = {"L0": 10403828, "L1": 2910160, "L2": 765843, "L3": 196102, "U2": 2137666, "U1": 8388378, "U0": 32243344}
    convs, everything = bench.pass_algorithmic_bytes(n, p)
    assert abs(convs - 8.569e9) < 0.01 * 8.569e9 and everything > convs
    assert bench.k3_algorithmic_bytes(1690768, 32248086, 16, 16) == 474430640
