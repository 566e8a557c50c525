"""End-to-end parity of the CUDA codec (pcgcv2_b200.Codec and the MinkowskiEngine shim) against
the CPU oracle with the reference's shipped weights (r3 / r7 fixtures).

Bars (BASELINE.json north_star): occupancy coordinates bit-exact after canonical sort, activations
within 1e-4 relative (max|d| / max|ref| per layer), bits within 1e-4 relative, D1 PSNR within 0.01 dB."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import codec_ref, metrics_ref, sparse_ref as S
from pcgcv2_b200 import ops, synth
from pcgcv2_b200.codec import Codec
from util import GOLDEN, canon, load_ckpt, with_batch

pytestmark = pytest.mark.gpu
ACT_TOL = 1e-4


def _sorted_rows(feats, coords):
    order = np.lexsort(np.asarray(coords).T[::-1])
    return feats[torch.from_numpy(order)], np.asarray(coords)[order]


def _check_layers(codec_record, oracle_record, relu_names=()):
    checked = 0
    for name, (t, keys, stride) in codec_record.items():
        if name not in oracle_record or keys is None:
            continue
        ref = oracle_record[name]
        if name in relu_names or name.endswith((".conv0_0", ".conv1_1")):
            ref = torch.relu(ref)                      # the codec fuses these ReLUs into the conv epilogue
        ref, ref_c = _sorted_rows(ref, oracle_record[name + ".C"])
        got, got_c = _sorted_rows(t.cpu(), ops.unpack_keys(keys, stride).cpu().numpy())
        assert (got_c == ref_c).all(), f"{name}: coordinate sets differ"
        err = float((got - ref).abs().max() / ref.abs().max().clamp_min(1e-30))
        assert err < ACT_TOL, f"{name}: relative error {err:.2e}"
        checked += 1
    return checked


@pytest.fixture(scope="module")
def r3():
    torch.set_flush_denormal(True)
    return load_ckpt("r3")


def test_cube32_layers_bitstream_and_decode_match_oracle(r3):
    """config 1: 32^3 random-occupancy cube, r3 checkpoint."""
    coords = with_batch(synth.random_cube(0, 32, 0.1))
    rec_ref = {}
    st_ref = codec_ref.encode(r3, coords, rec_ref)
    dec_ref, _ = codec_ref.decode(r3, st_ref, record=rec_ref)

    codec = Codec(r3)
    codec.record = {}
    st = codec.encode(coords[:, 1:])
    dec = codec.decode(st)
    relu_names = [f"{a}{i}" for a in ("encoder.down", "decoder.up", "encoder.conv", "decoder.conv") for i in range(3)]
    n = _check_layers(codec.record, rec_ref, relu_names=relu_names)
    assert n >= 30
    assert (st.coords == st_ref["C_coords"]).all()                       # canonical order, bit-exact
    assert st.H == st_ref["H"] and st.num_points == st_ref["num_points"]
    assert abs(len(st.F) - len(st_ref["F"])) * 8 <= max(8, 1e-4 * 8 * len(st_ref["F"]))
    assert (canon(dec) == canon(dec_ref[:, 1:])).all()                   # decoded occupancy, bit-exact
    gold = np.load(os.path.join(GOLDEN, "oracle_cube32_r3.npz"))
    assert (canon(dec) == canon(gold["dec_C"][:, 1:])).all()


@pytest.mark.parametrize("name", ["r3", "r7"])
def test_vox8_known_answer(name):
    """checkpoint-behaviour KAT: ellipsoid shell, 91 568 voxels."""
    kat = json.load(open(os.path.join(GOLDEN, "oracle_vox8_kat.json")))
    pts = synth.ellipsoid_vox8()
    codec = Codec(load_ckpt(name))
    st = codec.encode(pts)
    dec = codec.decode(st)
    assert len(st.coords) == kat[name]["N3"] and len(dec) == kat[name]["N_out"]
    assert abs(len(st.F) - kat[name]["F_bytes"]) * 8 <= max(8, 1e-4 * 8 * kat[name]["F_bytes"])
    assert abs(metrics_ref.d1_psnr(pts, dec, 256) - kat[name]["D1_psnr"]) < 0.01


def test_empty_tiny_and_duplicate_inputs(r3):
    codec = Codec(r3)
    one = codec.decode(codec.encode(np.array([[5, 6, 7]], dtype=np.int32)))
    assert one.shape == (1, 3)
    pts = synth.random_cube(1, 16, 0.2)
    dup = np.concatenate([pts, pts[::3]])
    a, b = codec.encode(pts), codec.encode(dup)
    assert a.F == b.F and (a.coords == b.coords).all() and a.num_points == b.num_points


def test_rho_changes_output_size(r3):
    pts = synth.random_cube(2, 32, 0.1)
    codec = Codec(r3)
    st = codec.encode(pts)
    assert len(codec.decode(st, rho=1.0)) == len(pts)
    assert len(codec.decode(st, rho=2.0)) == 2 * len(pts)


def test_shim_model_matches_codec_and_oracle(r3):
    """the same network driven through the drop-in MinkowskiEngine operator surface."""
    from pcgcv2_b200.model import load_model
    import MinkowskiEngine as ME
    from data_utils import sort_spare_tensor
    pts = synth.ellipsoid_vox8(n=150_000)
    coords = with_batch(pts)
    rng = np.random.default_rng(0)
    coords = coords[rng.permutation(len(coords))]                        # user order is arbitrary
    model = load_model(r3)
    with torch.no_grad():
        c, f = ME.utils.sparse_collate([torch.from_numpy(coords[:, 1:])], [torch.ones(len(coords), 1)])
        x = ME.SparseTensor(features=f, coordinates=c, tensor_stride=1, device="cuda")
        assert (x.C.cpu().numpy() == coords).all()                       # row order preserved (Appendix A.2)
        y_list = model.encoder(x)
        y = sort_spare_tensor(y_list[0])
    st_ref = codec_ref.encode(r3, coords)
    assert (y.C.cpu().numpy() == st_ref["y_C"]).all()
    err = float((y.F.cpu() - st_ref["y_F"]).abs().max() / st_ref["y_F"].abs().max())
    assert err < ACT_TOL
    assert [len(t) for t in y_list] == [len(st_ref["y_C"])] + np.frombuffer(st_ref["num_points"], np.int32).tolist()[:2]
    # decoder through the shim, fed with the oracle's quantised bottleneck
    nums = np.frombuffer(st_ref["num_points"], np.int32).tolist()
    yq = ME.SparseTensor(features=st_ref["y_F"].round().cuda(), coordinates=torch.from_numpy(st_ref["y_C"]).cuda(),
                         tensor_stride=8, device="cuda")
    with torch.no_grad():
        _, out = model.decoder(yq, [[n] for n in nums], [None] * 3, training=False)
    dec_ref, _ = codec_ref.decode(r3, st_ref)
    assert (canon(out.C.cpu().numpy()) == canon(dec_ref)).all()


def test_shim_error_behaviour(r3):
    import pcgcv2_b200
    pcgcv2_b200.install_shims()
    import MinkowskiEngine as ME
    f = torch.ones(4, 1)
    c = torch.tensor([[0, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]], dtype=torch.int32)
    with pytest.raises(ValueError):
        ME.SparseTensor(features=f, coordinates=c.long(), device="cuda")            # coords must be int32
    with pytest.raises(ValueError):
        ME.SparseTensor(features=f, coordinates=c, device="cpu")                    # no CPU backend
    a = ME.SparseTensor(features=f, coordinates=c, device="cuda")
    b = ME.SparseTensor(features=f, coordinates=c, device="cuda")
    with pytest.raises(ValueError):
        ME.cat(a, b)                                                                # different coordinate maps
    with pytest.raises(ValueError):
        a + b
    d = ME.SparseTensor(features=torch.ones(6, 1), coordinates=torch.cat([c, c[:2]]), device="cuda")
    assert len(d) == 4                                                              # duplicates collapse (A.2)


@pytest.mark.slow
def test_vox10_roundtrip_properties(r3):
    """BASELINE config 2 size (synthetic vox10, 795 124 voxels): size-independent properties."""
    pts = synth.synthetic_vox10(0)
    assert len(pts) == 795124
    codec = Codec(r3)
    st = codec.encode(pts)
    n2, n1, n0 = np.frombuffer(st.num_points, np.int32).tolist()
    assert n0 == len(pts) and len(st.coords) < n2 < n1 < n0
    # the bottleneck coordinates are exactly the occupied 8x8x8 cells of the input
    cells = np.unique(pts // 8, axis=0)
    assert (canon(st.coords) == canon(cells)).all()
    dec = codec.decode(st).copy()                                                     # decode() returns a reused pinned buffer
    assert len(dec) == n0 and len(np.unique(dec, axis=0)) == n0
    dec_cells = set(map(tuple, np.unique(dec // 8, axis=0).tolist()))
    assert dec_cells <= set(map(tuple, cells.tolist()))                              # decoded voxels stay inside coded cells
    assert len(dec_cells) > 0.99 * len(cells)
    assert codec.decode(st).tolist() == dec.tolist()                                # deterministic
    bpp = st.bits() / n0
    assert 0.03 < bpp < 0.12                                                         # r3 operating point (~0.07 bpp features)
    assert metrics_ref.d1_psnr(pts, dec, 1024) > 68.0


def test_h2_path_equals_tf32_path_and_falls_back_on_overflow(r3):
    """the pre-split half-precision kernels give the 3xTF32 pipeline's bitstream and occupancy; activations beyond
    the f16 range make the codec repeat the pass on the 3xTF32 kernels instead of returning a wrong result."""
    pts = synth.ellipsoid_vox8()
    fast, slow = Codec(r3), Codec(r3, use_h2=False)
    assert fast.packed_h2 and not slow.packed_h2
    a, b = fast.encode(pts), slow.encode(pts)
    assert a.F == b.F and a.H == b.H and (a.coords == b.coords).all()
    assert (canon(fast.decode(a)) == canon(slow.decode(b))).all() and fast.h2_fallbacks == 0
    fast._overflow.fill_(1)                                       # encoder side: a raised flag repeats the pass without h2
    a2 = fast.encode(pts)
    assert fast.h2_fallbacks == 1 and a2.F == b.F and (a2.coords == b.coords).all()
    big = dict(r3)                                                # same bitstream, synthesis activations scaled beyond f16
    big["decoder.up1.kernel"] = r3["decoder.up1.kernel"] * 32768.0
    big["decoder.up1.bias"] = r3["decoder.up1.bias"] * 32768.0
    fast, slow = Codec(big), Codec(big, use_h2=False)
    got, want = fast.decode(a), slow.decode(b)
    assert fast.h2_fallbacks >= 1 and (canon(got) == canon(want)).all()


def test_frame_pipeline_matches_single_frame_path(r3):
    """several frames in flight (one host thread + CUDA stream each) == the same frames one at a time."""
    from pcgcv2_b200.pipeline import FramePipeline
    frames = [synth.ellipsoid_vox8(), synth.random_cube(1, 32, 0.1), synth.random_cube(2, 48, 0.05), synth.ellipsoid_vox8(),
              synth.random_cube(3, 32, 0.2)]
    single = Codec(r3)
    want = []
    for f in frames:
        st = single.encode(f)
        want.append((st, single.decode(st).copy()))
    with FramePipeline(r3, depth=2) as pipe:
        for _ in range(2):                                        # twice: staging buffers are reused across calls
            got = pipe.roundtrip(frames)
            assert len(got) == len(frames)
            for (st, dec), (st_w, dec_w) in zip(got, want):
                assert st.F == st_w.F and st.H == st_w.H and st.num_points == st_w.num_points and (st.coords == st_w.coords).all()
                assert (canon(dec) == canon(dec_w)).all()
        dev = pipe.roundtrip([torch.from_numpy(frames[0]).cuda()], to_host=False)[0][1]
        assert dev.is_cuda and (canon(dev.cpu().numpy()) == canon(want[0][1])).all()
        with pytest.raises(Exception):
            pipe.roundtrip([np.zeros((4, 2), dtype=np.int32)])    # a worker's error reaches the caller
        assert len(pipe.roundtrip(frames[:1])) == 1               # and the pipeline stays usable


@pytest.mark.parametrize("use_h2", [True, False])
def test_irn_block_per_c_call_is_the_same_computation(r3, use_h2):
    """pcgc_irn_fwd (one C call per InceptionResNet block) issues the kernels the layer-by-layer path issues:
    bit-identical bitstream, bottleneck and decoded set; with and without the h2 kernels; also without octet kernels."""
    pts = synth.ellipsoid_vox8()
    for octet in (True, False):
        fused = Codec(r3, use_h2=use_h2, use_octet_kernels=octet, fuse_irn=True, merge_first=False, fuse_tail=False)
        plain = Codec(r3, use_h2=use_h2, use_octet_kernels=octet, fuse_irn=False)
        a, b = fused.encode(pts), plain.encode(pts)
        assert a.F == b.F and a.H == b.H and (a.coords == b.coords).all() and fused._irn_plans and not plain._irn_plans
        da, db = fused.decode(a, to_host=False), plain.decode(b, to_host=False)
        assert torch.equal(da, db)                                # same rows in the same order


def test_tcgen05_routes_give_the_same_stream_and_occupancy(r3):
    """every k=3 layer that has a tcgen05 / TMA kernel routed to it (wide_shapes="all") == none routed: same bitstream,
    same decoded set on the KAT cloud; the default routing is a subset of "all" and goes through pcgc_irn_fwd too."""
    pts = synth.ellipsoid_vox8()
    every, none, default = Codec(r3, wide_shapes="all"), Codec(r3, wide_shapes="none"), Codec(r3)
    assert len(every.packed_wide) >= 40 and not none.packed_wide and default.packed_wide
    a, b, c = every.encode(pts), none.encode(pts), default.encode(pts)
    assert a.F == b.F == c.F and a.H == b.H and (a.coords == b.coords).all()
    da, db, dc = every.decode(a), none.decode(b), default.decode(c)
    assert (canon(da) == canon(db)).all() and (canon(dc) == canon(db)).all()
    assert every.h2_fallbacks == 0 and default.h2_fallbacks == 0
    plain = Codec(r3, wide_shapes="all", fuse_irn=False)                 # layer-by-layer == one C call per block
    assert plain.encode(pts).F == a.F and (canon(plain.decode(a)) == canon(da)).all()


@pytest.mark.parametrize("merge_first,fuse_tail,dual", [(True, False, False), (False, True, False), (True, True, False), (True, True, True)])
def test_merged_first_layers_of_the_16_channel_blocks(r3, merge_first, fuse_tail, dual):
    """conv0_0 (k=3) + conv1_0 (k=1) of the finest decoder blocks as ONE k=3 convolution 16 -> 8 (conv1_0's weights at the
    centre offset, PCGC_IRN_MERGED_FIRST) and conv1_1 (k=3) + ReLU + conv1_2 (k=1) as one kernel (PCGC_IRN_FUSED_TAIL): the
    block output stays within the h2 tolerance of the layer-by-layer block, the stream is untouched (the analysis network has
    no 16-channel block) and the decoded set is the same on the KAT cloud."""
    pts = synth.ellipsoid_vox8()
    merged = Codec(r3, merge_first=merge_first, fuse_tail=fuse_tail, dual_second=dual)
    plain = Codec(r3, merge_first=False, fuse_tail=False)
    a, b = merged.encode(pts), plain.encode(pts)
    assert a.F == b.F and a.H == b.H and (a.coords == b.coords).all()
    da, db = merged.decode(a), plain.decode(b)
    assert any(p["args"].reserved == (1 if merge_first else 0) + (2 if fuse_tail else 0) + (4 if dual else 0) for p in merged._irn_plans.values())
    assert not any(p["args"].reserved for p in plain._irn_plans.values())
    assert (canon(da) == canon(db)).all()
    # one block in isolation on random features over the finest decoder set of the KAT cloud
    plain.record = {}
    plain.decode(b)
    x, keys, stride = plain.record["decoder.conv2"]
    plain.record = None
    from pcgcv2_b200.codec import _F, _Level
    parent = _Level((keys[::8] >> 3).contiguous(), 2 * stride)
    level = _Level(keys, stride, parent=parent)
    g = torch.Generator().manual_seed(1)
    xr = (torch.randn(x.shape, generator=g) * 2).cuda()
    with torch.no_grad(), ops.stream_scope():
        ym = merged._irn_fused("decoder.block2.0", _F(xr.clone()), level).f
        yp = plain._irn_fused("decoder.block2.0", _F(xr.clone()), level).f
    err = float((ym - yp).abs().max() / yp.abs().max())
    assert err < 3e-6, err


def test_coord_bits_hint_is_the_same_codec_and_falls_back(r3):
    """Codec(coord_bits=b): radix sorts over 3 b key bits -- identical stream and decoded set; a cloud that breaks the promise
    is detected by the device flag and coded again at full key width (coord_bits_fallbacks), with the right result."""
    pts = synth.ellipsoid_vox8()                                          # coordinates < 256
    full, hinted = Codec(r3), Codec(r3, coord_bits=8)
    a, b = full.encode(pts), hinted.encode(pts)
    assert a.F == b.F and a.H == b.H and a.C == b.C and a.num_points == b.num_points and (a.coords == b.coords).all()
    assert (canon(hinted.decode(b)) == canon(full.decode(a))).all() and hinted.coord_bits_fallbacks == 0
    narrow = Codec(r3, coord_bits=7)                                      # 2^7 = 128 < the cloud's extent
    c = narrow.encode(pts)
    assert narrow.coord_bits_fallbacks == 1 and narrow.coord_bits is None
    assert c.F == a.F and (c.coords == a.coords).all() and (canon(narrow.decode(c)) == canon(full.decode(a))).all()
    shifted = pts + np.array([[4096, 0, 0]], dtype=np.int32)              # decode side: key width comes from the stream's coordinates
    d = hinted.encode(shifted)
    assert hinted.coord_bits_fallbacks == 1
    assert (canon(hinted.decode(d)) == canon(full.decode(full.encode(shifted)))).all()


def test_fused_first_layer_gives_the_same_stream(r3):
    """Codec(fuse_conv0=True) (encoder.conv0 from the parent's map, no finest-level kernel map) == the generic first layer:
    same bitstream and decoded set on the KAT cloud and on a cube with duplicates removed."""
    for pts in (synth.ellipsoid_vox8(), synth.random_cube(0, 32, 0.1)):
        a, b = Codec(r3, fuse_conv0=True), Codec(r3, fuse_conv0=False)
        sa, sb = a.encode(pts), b.encode(pts)
        assert sa.F == sb.F and sa.H == sb.H and (sa.coords == sb.coords).all() and sa.num_points == sb.num_points
        assert (canon(a.decode(sa)) == canon(b.decode(sb))).all()
