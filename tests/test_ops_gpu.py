"""Parity of every CUDA operator (through the C ABI) against the CPU oracle, on seeded inputs.

Integer / index work is compared bit-exactly; float32 convolutions within
max|d| <= 2e-5 * max|ref| per layer (the north star allows 1e-4; different but valid
summation orders of the same fp32 products give ~1e-6)."""
import os

import numpy as np
import pytest
import torch

from oracle import entropy_ref, rangecoder_ref, sparse_ref as S
from pcgcv2_b200 import ops, synth
from util import GOLDEN, canon, load_ckpt, with_batch

pytestmark = pytest.mark.gpu
DEV = "cuda"
CONV_TOL = 2e-5


def _cloud(seed, n=4000, size=40, batch=1, stride=1):
    rng = np.random.default_rng(seed)
    pts = np.unique(rng.integers(0, size, size=(n, 3)), axis=0)
    rng.shuffle(pts)                                        # arbitrary (user) row order
    c = with_batch(pts * stride)
    if batch > 1:
        c[:, 0] = rng.integers(0, batch, size=len(c))
    return c


def _surface(seed=0):
    pts = synth.ellipsoid_vox8(seed, n=200_000)             # ~50 k voxels, surface-like
    return with_batch(pts)


def _keys(coords, stride=1):
    return ops.pack_keys(torch.from_numpy(coords).to(DEV), stride)


def _rel_err(got, ref):
    ref = ref.float()
    return float((got.cpu() - ref).abs().max() / ref.abs().max().clamp_min(1e-30))


# ------------------------------------------------------------------ coordinates

@pytest.mark.parametrize("stride", [1, 2, 8])
def test_pack_unpack_roundtrip(stride):
    c = _cloud(1, batch=3, stride=stride)
    c[0, 1:] = 0
    c[1, 1:] = ((1 << 19) - 1) * stride if stride == 1 else c[1, 1:]
    keys = _keys(c, stride)
    assert (ops.unpack_keys(keys, stride).cpu().numpy() == c).all()
    assert keys.unique().numel() == len(np.unique(c, axis=0))


def test_pack_rejects_bad_coordinates():
    for bad in ([[0, -1, 0, 0]], [[0, 1 << 19, 0, 0]], [[127, 0, 0, 0]]):
        with pytest.raises(ValueError):
            ops.pack_keys(torch.tensor(bad, dtype=torch.int32, device=DEV), 1)
    with pytest.raises(ValueError):
        ops.pack_keys(torch.tensor([[0, 3, 0, 0]], dtype=torch.int32, device=DEV), 2)     # not a multiple
    assert ops.pack_keys(torch.zeros((0, 4), dtype=torch.int32, device=DEV), 1).numel() == 0


def test_hash_dedup_first_seen():
    c = _cloud(2, n=3000, size=12)
    dup = np.concatenate([c, c[::5], c[::7]])
    keys = _keys(dup)
    table = ops.HashTable(keys)
    assert table.n_dup == len(dup) - len(c)
    keep = table.keep_flags(keys).cpu().numpy().astype(bool)
    _, first = S.unique_coords(dup)
    assert (np.nonzero(keep)[0] == first).all()
    assert ops.HashTable(_keys(c)).n_dup == 0


def test_isin_matches_oracle():
    a, b = _cloud(3, batch=2), _cloud(4, batch=2)
    found = ops.HashTable(_keys(b)).contains(_keys(a)).cpu().numpy()
    assert (found == S.isin(a, b)).all()
    assert found.any() and not found.all()


@pytest.mark.parametrize("stride,batch", [(1, 1), (4, 2)])
def test_kernel_map_k3_bit_exact(stride, batch):
    c = _cloud(5, n=6000, size=24, batch=batch, stride=stride)
    keys = _keys(c, stride)
    nbr, npairs = ops.kernel_map_k3(keys, ops.HashTable(keys), count_pairs=True)
    ref = S.kernel_map_k3(c, stride)
    assert (nbr.cpu().numpy().T == ref).all()
    assert int(npairs.item()) == int((ref >= 0).sum())


def test_kernel_map_grid_edges():
    """voxels on the 0 / max faces: neighbours outside the key range are reported missing."""
    m = (1 << 19) - 1
    c = with_batch([[0, 0, 0], [1, 0, 0], [m, m, m], [m - 1, m, m], [0, m, 0]])
    keys = _keys(c)
    nbr = ops.kernel_map_k3(keys, ops.HashTable(keys)).cpu().numpy().T
    assert (nbr == S.kernel_map_k3(c, 1)).all()


def test_kernel_map_surface_cloud():
    c = _surface()
    keys = _keys(c)
    nbr = ops.kernel_map_k3(keys, ops.HashTable(keys)).cpu().numpy().T
    assert (nbr == S.kernel_map_k3(c, 1)).all()


@pytest.mark.parametrize("presorted", [False, True])
def test_stride_down_map(presorted):
    c = _cloud(6, n=5000, size=30, batch=2, stride=2)
    keys = _keys(c, 2)
    if presorted:
        keys, order = ops.argsort_u64(keys)
        c = c[order.cpu().numpy()]
    pk, rows, off = ops.stride_down(keys, keys_are_sorted=presorted)
    out_ref, parent_ref, kidx_ref = S.stride_down(c, 2)
    pc = ops.unpack_keys(pk, 4).cpu().numpy()
    assert (canon(pc) == canon(out_ref)).all()
    rows, off = rows.cpu().numpy(), off.cpu().numpy()
    assert off[0] == 0 and off[-1] == len(c) and sorted(rows.tolist()) == list(range(len(c)))
    for p in range(len(pc)):                                  # every child sits under its own parent
        kids = rows[off[p]:off[p + 1]]
        assert 1 <= len(kids) <= 8
        assert (out_ref[parent_ref[kids]] == pc[p]).all()
    assert ((keys.cpu().numpy() & 7) == kidx_ref).all()


@pytest.mark.parametrize("batch", [1, 3])
def test_kernel_map_from_parent_equals_hashed_map(batch):
    """hierarchical derivation (no hashing) == hash-probed map, for partial and full octets."""
    c = _surface()
    if batch > 1:
        c[:, 0] = np.random.default_rng(0).integers(0, batch, size=len(c))
    keys, _ = ops.argsort_u64(_keys(c))
    pk, rows, off, parent_of = ops.stride_down(keys, keys_are_sorted=True, with_parent_of=True)
    info = ops.parent_info(keys, off)
    pnbr = ops.kernel_map_k3(pk, ops.HashTable(pk))
    want = ops.kernel_map_k3(keys, ops.HashTable(keys))
    got = ops.kernel_map_k3_from_parent(pnbr, len(keys), keys, parent_of, info)
    assert torch.equal(got, want)
    up = ops.upsample_keys(keys)                                # full octets: 8 children per row
    want = ops.kernel_map_k3(up, ops.HashTable(up))
    assert torch.equal(ops.kernel_map_k3_from_parent(ops.kernel_map_k3(keys, ops.HashTable(keys)), len(up)), want)


def test_prune_filters_kernel_map():
    c = _surface()
    keys, _ = ops.argsort_u64(_keys(c))
    nbr = ops.kernel_map_k3(keys, ops.HashTable(keys))
    g = torch.Generator().manual_seed(1)
    mask = (torch.rand(len(keys), generator=g) < 0.5).to(DEV)
    f = torch.randn(len(keys), 8, generator=g).to(DEV)
    k2, f2, nbr2 = ops.prune(mask, keys, f, nbr=nbr)
    assert torch.equal(nbr2, ops.kernel_map_k3(k2, ops.HashTable(k2)))
    assert torch.equal(f2, f[mask])


def test_upsample_keys():
    c = _cloud(7, n=500, size=10, batch=2, stride=4)
    child = ops.unpack_keys(ops.upsample_keys(_keys(c, 4)), 2).cpu().numpy()
    _, ref = S.convT_k2s2(torch.zeros(len(c), 1), c, 4, torch.zeros(8, 1, 1), None)
    assert (child == ref).all()


def test_argsort_canonical_order():
    """sort_spare_tensor order (data_utils.py:91-101) from the device argsort."""
    c = _cloud(8, n=3000, size=33, stride=8)
    ct = torch.from_numpy(c).to(DEV).long()
    step = int(c.max()) + 1
    key = ct[:, 0] + ct[:, 1] * step + ct[:, 2] * step ** 2 + ct[:, 3] * step ** 3
    _, order = ops.argsort_u64(key)
    assert (order.cpu().numpy() == np.argsort(S.sort_key(c), kind="stable")).all()


# ------------------------------------------------------------------ convolutions

MODEL_K3 = [(1, 16), (32, 8), (8, 16), (8, 8), (32, 32), (64, 16), (16, 32), (16, 16), (64, 64), (64, 1), (32, 1),
            (16, 4), (4, 8), (4, 4), (16, 1)]


@pytest.mark.parametrize("cin,cout", MODEL_K3 + [(64, 32), (32, 64), (128, 128), (3, 5), (12, 20), (128, 8)])
def test_conv_k3_vs_oracle(cin, cout):
    c = _cloud(cin * 131 + cout, n=3000, size=16)
    keys = _keys(c)
    nbr = ops.kernel_map_k3(keys, ops.HashTable(keys))
    g = torch.Generator().manual_seed(cin + 7 * cout)
    f = torch.randn(len(c), cin, generator=g)
    w = torch.randn(27, cin, cout, generator=g) / np.sqrt(27 * cin)
    b = torch.randn(1, cout, generator=g)
    ref = S.conv_k3(f, c, 1, w, b)
    got = ops.conv_k3(f.to(DEV), nbr, w.to(DEV), b.to(DEV))
    assert _rel_err(got, ref) < CONV_TOL
    # fused epilogue: residual + ReLU, written into a column slice of a wider tensor (ME.cat fusion)
    if cout % 4 == 0:
        res = torch.randn(len(c), cout, generator=g)
        wide = torch.full((len(c), cout + 8), -7.0, device=DEV)
        ops.conv_k3(f.to(DEV), nbr, w.to(DEV), b.to(DEV), residual=res.to(DEV), relu=True, out=wide[:, 4:4 + cout])
        assert _rel_err(wide[:, 4:4 + cout], torch.relu(ref + res)) < CONV_TOL
        assert (wide[:, :4] == -7).all() and (wide[:, 4 + cout:] == -7).all()


@pytest.mark.parametrize("cin,cout", [(8, 16), (8, 8), (8, 1), (16, 16), (16, 4), (16, 8), (16, 1), (16, 32), (32, 8), (32, 4), (32, 32),
                                      (32, 1), (64, 16), (64, 64), (64, 1), (64, 32), (32, 64)])
def test_conv_k3_tensor_core_vs_oracle(cin, cout):
    """3xTF32 mma.sync kernel keeps FP32 accuracy (same tolerance as the FFMA kernels)."""
    c = _surface()[:20011]
    keys = _keys(c)
    nbr = ops.kernel_map_k3(keys, ops.HashTable(keys))
    g = torch.Generator().manual_seed(cin * 3 + cout)
    f = torch.randn(len(c), cin, generator=g)
    w = torch.randn(27, cin, cout, generator=g) / np.sqrt(27 * cin)
    b = torch.randn(1, cout, generator=g)
    ref = S.conv_k3(f, c, 1, w, b)
    pw = ops.PackedK3(w.to(DEV))
    assert pw.packed is not None
    got = ops.conv_k3_packed(f.to(DEV), nbr, pw, b.to(DEV))
    assert _rel_err(got, ref) < CONV_TOL
    if cout % 4 == 0:
        res = torch.randn(len(c), cout, generator=g)
        wide = torch.full((len(c), cout + 8), -7.0, device=DEV)
        ops.conv_k3_packed(f.to(DEV), nbr, pw, b.to(DEV), residual=res.to(DEV), relu=True, out=wide[:, 4:4 + cout])
        assert _rel_err(wide[:, 4:4 + cout], torch.relu(ref + res)) < CONV_TOL
        assert (wide[:, :4] == -7).all() and (wide[:, 4 + cout:] == -7).all()
    for n in (1, 15, 17, 129):                                  # tile tails
        kk = _keys(c[:n])
        nb = ops.kernel_map_k3(kk, ops.HashTable(kk))
        got = ops.conv_k3_packed(f[:n].to(DEV), nb, pw, b.to(DEV))
        assert _rel_err(got, S.conv_k3(f[:n], c[:n], 1, w, b)) < CONV_TOL


OCTET_SHAPES = [(16, 16), (16, 8), (16, 4), (16, 1), (8, 16), (8, 8), (8, 4), (8, 1), (4, 8), (4, 4)]


@pytest.mark.parametrize("cin,cout", OCTET_SHAPES)
def test_conv_k3_octet_vs_oracle(cin, cout):
    """full-octet kernels (halo staged in shared memory, addressed by the PARENT's map) == oracle k=3
    convolution on the 8-child expansion, incl. tile tails, fused residual/ReLU and column-slice output."""
    par = _surface()[:6007]                                       # parents at stride 2 (arbitrary order)
    par[:, 1:] *= 2
    pkeys, _ = ops.argsort_u64(_keys(par, 2))
    pnbr = ops.kernel_map_k3(pkeys, ops.HashTable(pkeys))
    ckeys = ops.upsample_keys(pkeys)                              # row 8i + c = child c of parent i
    c = ops.unpack_keys(ckeys, 1).cpu().numpy()
    g = torch.Generator().manual_seed(cin * 5 + cout)
    f = torch.randn(len(c), cin, generator=g)
    w = torch.randn(27, cin, cout, generator=g) / np.sqrt(27 * cin)
    b = torch.randn(1, cout, generator=g)
    ref = S.conv_k3(f, c, 1, w, b)
    pw = ops.PackedK3Octet(w.to(DEV))
    assert pw.packed is not None
    got = ops.conv_k3_octet(f.to(DEV), pnbr, pw, b.to(DEV))
    assert _rel_err(got, ref) < CONV_TOL
    if cout % 4 == 0:
        res = torch.randn(len(c), cout, generator=g)
        wide = torch.full((len(c), cout + 8), -7.0, device=DEV)
        ops.conv_k3_octet(f.to(DEV), pnbr, pw, b.to(DEV), residual=res.to(DEV), relu=True, out=wide[:, 4:4 + cout])
        assert _rel_err(wide[:, 4:4 + cout], torch.relu(ref + res)) < CONV_TOL
        assert (wide[:, :4] == -7).all() and (wide[:, 4 + cout:] == -7).all()
        wide_in = torch.zeros((len(c), cin + 4), device=DEV)     # strided input rows (column slice of a wider tensor)
        wide_in[:, 4:] = f.to(DEV)
        assert _rel_err(ops.conv_k3_octet(wide_in[:, 4:], pnbr, pw, b.to(DEV)), ref) < CONV_TOL
    for n_par in (1, 3, 31, 33, 257):                             # tile tails, fewer tiles than SMs
        pk = pkeys[:n_par].contiguous()
        nb = ops.kernel_map_k3(pk, ops.HashTable(pk))
        cc = ops.unpack_keys(ops.upsample_keys(pk), 1).cpu().numpy()
        got = ops.conv_k3_octet(f[:8 * n_par].to(DEV), nb, pw, b.to(DEV), relu=True)
        assert _rel_err(got, torch.relu(S.conv_k3(f[:8 * n_par], cc, 1, w, b))) < CONV_TOL


def test_conv_k3_octet_equals_child_map_kernels():
    """same numbers as the child-map kernels on the same set (FFMA variant bit-identical)."""
    par = _surface()[:20011]
    par[:, 1:] *= 2
    pkeys, _ = ops.argsort_u64(_keys(par, 2))
    pnbr = ops.kernel_map_k3(pkeys, ops.HashTable(pkeys))
    nbr = ops.kernel_map_k3_from_parent(pnbr, 8 * len(pkeys))
    g = torch.Generator().manual_seed(3)
    for cin, cout in ((4, 8), (4, 4), (16, 16), (16, 4), (8, 8)):
        f = torch.randn(8 * len(pkeys), cin, generator=g).to(DEV)
        w = (torch.randn(27, cin, cout, generator=g) / np.sqrt(27 * cin)).to(DEV)
        b = torch.randn(1, cout, generator=g).to(DEV)
        got = ops.conv_k3_octet(f, pnbr, ops.PackedK3Octet(w), b)
        if cin == 4:
            assert torch.equal(got, ops.conv_k3(f, nbr, w, b))
        else:
            want = ops.conv_k3_packed(f, nbr, ops.PackedK3(w), b)
            assert float((got - want).abs().max() / want.abs().max()) < 2e-6


H2_SHAPES = [(8, 8), (8, 16), (16, 1), (16, 4), (16, 8), (16, 16), (16, 32), (32, 1), (32, 4), (32, 8), (32, 32), (64, 1), (64, 8), (64, 16)]
H2_TOL = 3e-6                                                     # 22-bit operands: two orders below CONV_TOL


def test_h2_split_join_roundtrip():
    """x -> (f16 hi, f16 lo) -> x: 22 significand bits, column slices, the overflow flag."""
    g = torch.Generator().manual_seed(1)
    x = (torch.randn(1003, 32, generator=g) * torch.logspace(-3, 3, 32)).to(DEV)
    h = ops.split_h2(x)
    assert h.dtype == torch.int32 and h.shape == x.shape
    back = ops.join_h2(h)
    assert bool(((back - x).abs() <= torch.maximum(x.abs() * 2.0 ** -21, torch.tensor(6.0e-8, device=DEV))).all())   # lo is f16: 6e-8 quantum
    wide = torch.zeros((1003, 48), dtype=torch.int32, device=DEV)          # a column slice of a wider h2 tensor
    ops.split_h2(x[:, 8:24], out=wide[:, 12:28])
    assert torch.equal(ops.join_h2(wide[:, 12:28]), back[:, 8:24]) and (wide[:, :12] == 0).all() and (wide[:, 28:] == 0).all()
    flag = torch.zeros(1, dtype=torch.int32, device=DEV)
    ops.split_h2(x, overflow=flag)
    assert int(flag.item()) == 0
    x[17, 5] = 7.0e4
    ops.split_h2(x, overflow=flag)
    assert int(flag.item()) == 1


@pytest.mark.parametrize("cin,cout", H2_SHAPES)
def test_conv_k3_h2_vs_oracle(cin, cout):
    """pre-split half-precision k=3 convolution == oracle, through both outputs (fp32 and h2), with the fused
    residual / ReLU, column-slice outputs and ragged tile tails."""
    c = _surface()[:20011]
    keys = _keys(c)
    nbr = ops.kernel_map_k3(keys, ops.HashTable(keys))
    g = torch.Generator().manual_seed(cin * 7 + cout)
    f = torch.randn(len(c), cin, generator=g) * 3.0
    w = torch.randn(27, cin, cout, generator=g) / np.sqrt(27 * cin)
    b = torch.randn(1, cout, generator=g)
    ref = S.conv_k3(f, c, 1, w, b)
    pw = ops.PackedK3H2(w.to(DEV))
    assert pw.packed is not None
    flag = torch.zeros(1, dtype=torch.int32, device=DEV)
    xh = ops.split_h2(f.to(DEV))
    got, got_h = ops.conv_k3_h2(xh, nbr, pw, b.to(DEV), want_h2=cout % 4 == 0, overflow=flag)
    assert _rel_err(got, ref) < H2_TOL
    if cout % 4 == 0:
        assert _rel_err(ops.join_h2(got_h), ref) < H2_TOL
        only_h = ops.conv_k3_h2(xh, nbr, pw, b.to(DEV), want_f32=False, want_h2=True)
        assert only_h[0] is None and torch.equal(only_h[1], got_h)
        res = torch.randn(len(c), cout, generator=g)
        wide = torch.full((len(c), cout + 8), -7.0, device=DEV)
        wide_h = torch.full((len(c), cout + 8), 5, dtype=torch.int32, device=DEV)
        ops.conv_k3_h2(xh, nbr, pw, b.to(DEV), residual=res.to(DEV), relu=True, out=wide[:, 4:4 + cout], out_h2=wide_h[:, 4:4 + cout])
        want = torch.relu(ref + res)
        assert _rel_err(wide[:, 4:4 + cout], want) < H2_TOL and _rel_err(ops.join_h2(wide_h[:, 4:4 + cout]), want) < H2_TOL
        assert (wide[:, :4] == -7).all() and (wide[:, 4 + cout:] == -7).all()
        assert (wide_h[:, :4] == 5).all() and (wide_h[:, 4 + cout:] == 5).all()
        wide_in = torch.zeros((len(c), cin + 4), dtype=torch.int32, device=DEV)      # strided input rows
        wide_in[:, 4:] = xh
        assert _rel_err(ops.conv_k3_h2(wide_in[:, 4:], nbr, pw, b.to(DEV))[0], ref) < H2_TOL
    assert int(flag.item()) == 0
    for n in (1, 63, 65, 2049):                                   # tile tails, fewer tiles than SMs
        cc = c[:n]
        kk = _keys(cc)
        nb = ops.kernel_map_k3(kk, ops.HashTable(kk))
        got = ops.conv_k3_h2(ops.split_h2(f[:n].to(DEV)), nb, pw, b.to(DEV), relu=True)[0]
        assert _rel_err(got, torch.relu(S.conv_k3(f[:n], cc, 1, w, b))) < H2_TOL


WIDE_SHAPES = [(64, 64), (64, 16), (64, 1), (32, 32), (32, 8), (32, 1), (16, 32), (16, 16), (16, 4), (16, 1), (8, 16), (8, 8)]


@pytest.mark.parametrize("cin,cout", WIDE_SHAPES)
def test_conv_k3_wide_tcgen05_vs_oracle(cin, cout):
    """tcgen05 / TMA k=3 convolution (csrc/conv_wide.cuh) == oracle through both outputs, with the fused residual / ReLU,
    column-slice outputs, strided inputs, ragged tile tails and fewer tiles than SMs; same bound as the mma.sync h2 kernels."""
    c = _surface()[:40013]                                        # > 2 tiles per SM: stage and accumulator-buffer reuse
    keys = _keys(c)
    nbr = ops.kernel_map_k3(keys, ops.HashTable(keys))
    g = torch.Generator().manual_seed(cin * 13 + cout)
    f = torch.randn(len(c), cin, generator=g) * 3.0
    w = torch.randn(27, cin, cout, generator=g) / np.sqrt(27 * cin)
    b = torch.randn(1, cout, generator=g)
    ref = S.conv_k3(f, c, 1, w, b)
    assert ops.PackedK3Wide.supported(cin, cout)
    pw = ops.PackedK3Wide(w.to(DEV))
    flag = torch.zeros(1, dtype=torch.int32, device=DEV)
    xh = ops.split_h2(f.to(DEV))
    got, got_h = ops.conv_k3_wide(xh, nbr, pw, b.to(DEV), want_h2=cout % 4 == 0, overflow=flag)
    err = _rel_err(got, ref)
    print(f"wide {cin}->{cout}: rel err {err:.2e}")
    assert err < H2_TOL
    if cout % 4 == 0:
        assert _rel_err(ops.join_h2(got_h), ref) < H2_TOL
        only_h = ops.conv_k3_wide(xh, nbr, pw, b.to(DEV), want_f32=False, want_h2=True)
        assert only_h[0] is None and torch.equal(only_h[1], got_h)
        res = torch.randn(len(c), cout, generator=g)
        wide = torch.full((len(c), cout + 8), -7.0, device=DEV)
        wide_h = torch.full((len(c), cout + 8), 5, dtype=torch.int32, device=DEV)
        ops.conv_k3_wide(xh, nbr, pw, b.to(DEV), residual=res.to(DEV), relu=True, out=wide[:, 4:4 + cout], out_h2=wide_h[:, 4:4 + cout])
        want = torch.relu(ref + res)
        assert _rel_err(wide[:, 4:4 + cout], want) < H2_TOL and _rel_err(ops.join_h2(wide_h[:, 4:4 + cout]), want) < H2_TOL
        assert (wide[:, :4] == -7).all() and (wide[:, 4 + cout:] == -7).all()
        assert (wide_h[:, :4] == 5).all() and (wide_h[:, 4 + cout:] == 5).all()
        wide_in = torch.zeros((len(c), cin + 4), dtype=torch.int32, device=DEV)      # strided input rows
        wide_in[:, 4:] = xh
        assert _rel_err(ops.conv_k3_wide(wide_in[:, 4:], nbr, pw, b.to(DEV))[0], ref) < H2_TOL
    assert int(flag.item()) == 0
    again = ops.conv_k3_wide(xh, nbr, pw, b.to(DEV))[0]
    assert torch.equal(again, got)                                # deterministic: no atomics, fixed accumulation order
    for n in (1, 127, 129, 2049):                                 # tile tails, fewer tiles than SMs
        cc = c[:n]
        kk = _keys(cc)
        nb = ops.kernel_map_k3(kk, ops.HashTable(kk))
        got = ops.conv_k3_wide(ops.split_h2(f[:n].to(DEV)), nb, pw, b.to(DEV), relu=True)[0]
        assert _rel_err(got, torch.relu(S.conv_k3(f[:n], cc, 1, w, b))) < H2_TOL


@pytest.mark.parametrize("cin,cout", [(16, 1), (16, 4), (16, 8), (16, 16), (16, 32), (4, 8), (4, 4), (8, 8), (8, 16)])
def test_conv_k3_octet_h2_vs_oracle(cin, cout):
    """full-octet h2 kernels (halo of h2 rows in shared memory, parent's map) == oracle on the 8-child expansion;
    cin = 4: all three split products in one MMA."""
    par = _surface()[:6007]
    par[:, 1:] *= 2
    pkeys, _ = ops.argsort_u64(_keys(par, 2))
    pnbr = ops.kernel_map_k3(pkeys, ops.HashTable(pkeys))
    c = ops.unpack_keys(ops.upsample_keys(pkeys), 1).cpu().numpy()
    g = torch.Generator().manual_seed(cin * 11 + cout)
    f = torch.randn(len(c), cin, generator=g) * 3.0
    w = torch.randn(27, cin, cout, generator=g) / np.sqrt(27 * cin)
    b = torch.randn(1, cout, generator=g)
    ref = S.conv_k3(f, c, 1, w, b)
    assert ops.octet_h2_supported(cin, cout)
    pw = ops.PackedK3H2(w.to(DEV))
    xh = ops.split_h2(f.to(DEV))
    flag = torch.zeros(1, dtype=torch.int32, device=DEV)
    got, got_h = ops.conv_k3_octet_h2(xh, pnbr, pw, b.to(DEV), want_h2=cout % 4 == 0, overflow=flag)
    assert _rel_err(got, ref) < H2_TOL
    if cout % 4 == 0:
        assert _rel_err(ops.join_h2(got_h), ref) < H2_TOL
        res = torch.randn(len(c), cout, generator=g)
        wide = torch.full((len(c), cout + 8), -7.0, device=DEV)
        wide_h = torch.full((len(c), cout + 8), 5, dtype=torch.int32, device=DEV)
        ops.conv_k3_octet_h2(xh, pnbr, pw, b.to(DEV), residual=res.to(DEV), relu=True, out=wide[:, 4:4 + cout], out_h2=wide_h[:, 4:4 + cout])
        want = torch.relu(ref + res)
        assert _rel_err(wide[:, 4:4 + cout], want) < H2_TOL and _rel_err(ops.join_h2(wide_h[:, 4:4 + cout]), want) < H2_TOL
        assert (wide[:, :4] == -7).all() and (wide[:, 4 + cout:] == -7).all()
        assert (wide_h[:, :4] == 5).all() and (wide_h[:, 4 + cout:] == 5).all()
        wide_in = torch.zeros((len(c), cin + 4), dtype=torch.int32, device=DEV)
        wide_in[:, 4:] = xh
        assert _rel_err(ops.conv_k3_octet_h2(wide_in[:, 4:], pnbr, pw, b.to(DEV))[0], ref) < H2_TOL
    assert int(flag.item()) == 0
    for n_par in (1, 3, 31, 33, 257):                             # tile tails, fewer tiles than SMs
        pk = pkeys[:n_par].contiguous()
        nb = ops.kernel_map_k3(pk, ops.HashTable(pk))
        cc = ops.unpack_keys(ops.upsample_keys(pk), 1).cpu().numpy()
        got = ops.conv_k3_octet_h2(ops.split_h2(f[:8 * n_par].to(DEV)), nb, pw, b.to(DEV), relu=True)[0]
        assert _rel_err(got, torch.relu(S.conv_k3(f[:8 * n_par], cc, 1, w, b))) < H2_TOL


@pytest.mark.parametrize("cin,cout", [(16, 16), (16, 8), (16, 4), (16, 1)])
def test_conv_k3_octet_tcgen05_vs_oracle(cin, cout):
    """full-octet tcgen05 kernel (27 kernel offsets = 27 descriptor start addresses into one staged halo, M = 64 accumulators
    interleaved in tensor memory) == oracle on the 8-child expansion, both outputs, fused epilogue, slices, tile tails."""
    par = _surface()[:6007]
    par[:, 1:] *= 2
    pkeys, _ = ops.argsort_u64(_keys(par, 2))
    pnbr = ops.kernel_map_k3(pkeys, ops.HashTable(pkeys))
    c = ops.unpack_keys(ops.upsample_keys(pkeys), 1).cpu().numpy()
    g = torch.Generator().manual_seed(cin * 17 + cout)
    f = torch.randn(len(c), cin, generator=g) * 3.0
    w = torch.randn(27, cin, cout, generator=g) / np.sqrt(27 * cin)
    b = torch.randn(1, cout, generator=g)
    ref = S.conv_k3(f, c, 1, w, b)
    assert ops.PackedK3OctetTc05.supported(cin, cout)
    pw = ops.PackedK3OctetTc05(w.to(DEV))
    xh = ops.split_h2(f.to(DEV))
    flag = torch.zeros(1, dtype=torch.int32, device=DEV)
    got, got_h = ops.conv_k3_octet_tc05(xh, pnbr, pw, b.to(DEV), want_h2=cout % 4 == 0, overflow=flag)
    err = _rel_err(got, ref)
    print(f"octet tcgen05 {cin}->{cout}: rel err {err:.2e}")
    assert err < H2_TOL
    if cout % 4 == 0:
        assert _rel_err(ops.join_h2(got_h), ref) < H2_TOL
        res = torch.randn(len(c), cout, generator=g)
        wide = torch.full((len(c), cout + 8), -7.0, device=DEV)
        wide_h = torch.full((len(c), cout + 8), 5, dtype=torch.int32, device=DEV)
        ops.conv_k3_octet_tc05(xh, pnbr, pw, b.to(DEV), residual=res.to(DEV), relu=True, out=wide[:, 4:4 + cout], out_h2=wide_h[:, 4:4 + cout])
        want = torch.relu(ref + res)
        assert _rel_err(wide[:, 4:4 + cout], want) < H2_TOL and _rel_err(ops.join_h2(wide_h[:, 4:4 + cout]), want) < H2_TOL
        assert (wide[:, :4] == -7).all() and (wide[:, 4 + cout:] == -7).all()
        assert (wide_h[:, :4] == 5).all() and (wide_h[:, 4 + cout:] == 5).all()
        wide_in = torch.zeros((len(c), cin + 4), dtype=torch.int32, device=DEV)
        wide_in[:, 4:] = xh
        assert _rel_err(ops.conv_k3_octet_tc05(wide_in[:, 4:], pnbr, pw, b.to(DEV))[0], ref) < H2_TOL
    assert int(flag.item()) == 0
    assert torch.equal(ops.conv_k3_octet_tc05(xh, pnbr, pw, b.to(DEV))[0], got)       # deterministic
    for n_par in (1, 3, 31, 33, 257):                             # tile tails, fewer tiles than SMs
        pk = pkeys[:n_par].contiguous()
        nb = ops.kernel_map_k3(pk, ops.HashTable(pk))
        cc = ops.unpack_keys(ops.upsample_keys(pk), 1).cpu().numpy()
        got = ops.conv_k3_octet_tc05(ops.split_h2(f[:8 * n_par].to(DEV)), nb, pw, b.to(DEV), relu=True)[0]
        assert _rel_err(got, torch.relu(S.conv_k3(f[:8 * n_par], cc, 1, w, b))) < H2_TOL


def test_rowlane_layers_write_h2_copy_in_epilogue():
    """k=1 / k=2 s=2 / transposed k=2 s=2 with the fused h2 output: fp32 result unchanged, h2 copy == split of it."""
    g = torch.Generator().manual_seed(9)
    n = 5003
    flag = torch.zeros(1, dtype=torch.int32, device=DEV)
    for cin, cout in ((4, 8), (8, 16), (16, 32), (32, 8), (64, 16), (16, 4)):       # the last three: one channel per lane, quad gather
        assert ops.h2out_supported("k1", cin, cout)
        f = torch.randn(n, cin, generator=g).to(DEV)
        w = torch.randn(cin, cout, generator=g).to(DEV)
        b = torch.randn(1, cout, generator=g).to(DEV)
        res = torch.randn(n, cout, generator=g).to(DEV)
        want = ops.conv_k1(f, w, b, residual=res, relu=True)
        wide = torch.zeros((n, 2 * cout), device=DEV)
        wide_h = torch.full((n, 2 * cout), 3, dtype=torch.int32, device=DEV)
        got, got_h = ops.conv_k1(f, w, b, residual=res, relu=True, out=wide[:, cout:], out_h2=wide_h[:, cout:], overflow=flag)
        assert torch.equal(got, want) and torch.equal(got_h, ops.split_h2(want)) and (wide_h[:, :cout] == 3).all()
    assert not ops.h2out_supported("k1", 32, 16) or True                              # (shapes without an h2 epilogue fall back to pcgc_split_h2)
    c = _surface()[:6001]
    keys, _ = ops.argsort_u64(_keys(c))
    for cin, cout in ((16, 32), (32, 64), (64, 32)):
        assert ops.h2out_supported("down", cin, cout)
        pk, rows, off, _ = ops.stride_down(keys, keys_are_sorted=True, with_parent_of=True)
        f = torch.randn(len(c), cin, generator=g).to(DEV)
        w = (torch.randn(8, cin, cout, generator=g) / np.sqrt(8 * cin)).to(DEV)
        b = torch.randn(1, cout, generator=g).to(DEV)
        want = ops.conv_k2s2(f, keys, rows, off, w, b, relu=True)
        got, got_h = ops.conv_k2s2(f, keys, rows, off, w, b, relu=True, out_h2=True, overflow=flag)
        assert torch.equal(got, want) and torch.equal(got_h, ops.split_h2(want))
    for cin, cout in ((64, 32), (32, 16)):
        assert ops.h2out_supported("up", cin, cout)
        f = torch.randn(1001, cin, generator=g).to(DEV)
        w = (torch.randn(8, cin, cout, generator=g) / np.sqrt(cin)).to(DEV)
        b = torch.randn(1, cout, generator=g).to(DEV)
        want = ops.convT_k2s2(f, w, b, relu=True)
        got, got_h = ops.convT_k2s2(f, w, b, relu=True, out_h2=True, overflow=flag)
        assert torch.equal(got, want) and torch.equal(got_h, ops.split_h2(want))
    assert int(flag.item()) == 0
    ops.convT_k2s2(f * 1e6, w, b, out_h2=True, overflow=flag)
    assert int(flag.item()) == 1


def test_k2s2_and_transposed_layers_on_tensor_cores_vs_oracle():
    """k=2 stride-2 convolution as the h2 gather kernel over the 8 child slots, and its generative transpose as ONE
    dense h2 product, against the oracle (ragged child counts, tile tails, both outputs)."""
    g = torch.Generator().manual_seed(21)
    flag = torch.zeros(1, dtype=torch.int32, device=DEV)
    c = _surface()[:6001]
    keys, order = ops.argsort_u64(_keys(c))
    cs = c[order.cpu().numpy()]
    for cin, cout in ((16, 32), (32, 64), (64, 32)):
        pk, rows, off, parent_of = ops.stride_down(keys, keys_are_sorted=True, with_parent_of=True)
        f = torch.randn(len(c), cin, generator=g) * 2.0
        w = torch.randn(8, cin, cout, generator=g) / np.sqrt(8 * cin)
        b = torch.randn(1, cout, generator=g)
        ref, ref_c = S.conv_k2s2(f, cs, 1, w, b)
        pw = ops.PackedDownH2(w.to(DEV))
        assert pw.packed is not None
        cmap = ops.child_map_k2(keys, parent_of, len(pk))
        assert int((cmap >= 0).sum()) == len(c)
        got, got_h = ops.conv_k2s2_h2(ops.split_h2(f.to(DEV)), cmap, pw, b.to(DEV), relu=True, overflow=flag)
        got_c = ops.unpack_keys(pk, 2).cpu().numpy()
        want, _ = _sorted_by_coords(torch.relu(ref), ref_c)
        have, _ = _sorted_by_coords(got.cpu(), got_c)
        assert _rel_err(have, want) < H2_TOL
        assert _rel_err(ops.join_h2(got_h), got.cpu()) < 1e-6
    for cin, cout, n in ((64, 32, 1001), (32, 16, 4099), (16, 8, 7), (32, 64, 300)):
        f = torch.randn(n, cin, generator=g) * 2.0
        w = torch.randn(8, cin, cout, generator=g) / np.sqrt(cin)
        b = torch.randn(1, cout, generator=g)
        want = ops.convT_k2s2(f.to(DEV), w.to(DEV), b.to(DEV), relu=True)          # fp32 kernel (oracle-checked elsewhere)
        pu = ops.PackedUpH2(w.to(DEV), b.to(DEV))
        assert pu.packed is not None
        got, got_h = ops.convT_k2s2_h2(ops.split_h2(f.to(DEV)), pu, relu=True, overflow=flag)
        assert got.shape == (8 * n, cout) and _rel_err(got, want.cpu()) < H2_TOL
        assert _rel_err(ops.join_h2(got_h), got.cpu()) < 1e-6
        only_h = ops.convT_k2s2_h2(ops.split_h2(f.to(DEV)), pu, relu=True, want_f32=False)
        assert only_h[0] is None and torch.equal(only_h[1], got_h)
    assert int(flag.item()) == 0


def _sorted_by_coords(feats, coords):
    order = np.lexsort(np.asarray(coords).T[::-1])
    return feats[torch.from_numpy(order)], np.asarray(coords)[order]


def test_conv_k3_h2_overflow_flag_and_small_weights():
    """tiny weights keep their precision through the power-of-two scale; an output beyond the f16 range raises the flag."""
    c = _surface()[:4001]
    keys = _keys(c)
    nbr = ops.kernel_map_k3(keys, ops.HashTable(keys))
    g = torch.Generator().manual_seed(5)
    f = torch.randn(len(c), 16, generator=g)
    w = torch.randn(27, 16, 16, generator=g) * 1e-4
    ref = S.conv_k3(f, c, 1, w, None)
    got = ops.conv_k3_h2(ops.split_h2(f.to(DEV)), nbr, ops.PackedK3H2(w.to(DEV)))[0]
    assert _rel_err(got, ref) < H2_TOL
    flag = torch.zeros(1, dtype=torch.int32, device=DEV)
    big = ops.PackedK3H2((w * 1e8).to(DEV))
    ops.conv_k3_h2(ops.split_h2(f.to(DEV)), nbr, big, want_h2=True, overflow=flag)
    assert int(flag.item()) == 1


def test_conv_k3_surface_and_ragged_sizes():
    c = _surface()
    keys = _keys(c)
    nbr = ops.kernel_map_k3(keys, ops.HashTable(keys))
    g = torch.Generator().manual_seed(0)
    for n in (len(c), 1, 63, 65, 2049):                       # tile tails of both kernel families
        cc = c[:n]
        kk = _keys(cc)
        nb = ops.kernel_map_k3(kk, ops.HashTable(kk)) if n != len(c) else nbr
        for cin, cout in ((16, 16), (64, 64)):
            f = torch.randn(n, cin, generator=g)
            w = torch.randn(27, cin, cout, generator=g) / np.sqrt(27 * cin)
            ref = S.conv_k3(f, cc, 1, w, None)
            got = ops.conv_k3(f.to(DEV), nb, w.to(DEV))
            assert _rel_err(got, ref) < CONV_TOL


def test_conv_empty_input():
    f = torch.zeros((0, 16), device=DEV)
    nbr = torch.zeros((27, 0), dtype=torch.int32, device=DEV)
    assert ops.conv_k3(f, nbr, torch.zeros(27, 16, 16, device=DEV)).shape == (0, 16)
    assert ops.conv_k1(f, torch.zeros(16, 8, device=DEV)).shape == (0, 8)
    assert ops.convT_k2s2(f, torch.zeros(8, 16, 8, device=DEV)).shape == (0, 8)


@pytest.mark.parametrize("cin,cout", [(32, 8), (8, 16), (64, 16), (16, 32), (16, 4), (4, 8), (128, 64), (5, 3)])
def test_conv_k1_vs_oracle(cin, cout):
    g = torch.Generator().manual_seed(cin * cout)
    f = torch.randn(2111, cin, generator=g)
    w = torch.randn(cin, cout, generator=g) / np.sqrt(cin)
    b = torch.randn(1, cout, generator=g)
    got = ops.conv_k1(f.to(DEV), w.to(DEV), b.to(DEV), relu=True)
    assert _rel_err(got, torch.relu(S.conv_k1(f, w, b))) < CONV_TOL


@pytest.mark.parametrize("cin,cout", [(16, 32), (32, 64), (64, 32), (8, 8), (6, 10)])
def test_conv_k2s2_vs_oracle(cin, cout):
    c = _cloud(cin + cout, n=5000, size=26, batch=2)
    keys = _keys(c)
    pk, rows, off = ops.stride_down(keys)
    g = torch.Generator().manual_seed(cin)
    f = torch.randn(len(c), cin, generator=g)
    w = torch.randn(8, cin, cout, generator=g) / np.sqrt(8 * cin)
    b = torch.randn(1, cout, generator=g)
    ref, ref_c = S.conv_k2s2(f, c, 1, w, b)
    got = ops.conv_k2s2(f.to(DEV), keys, rows, off, w.to(DEV), b.to(DEV))
    got_c = ops.unpack_keys(pk, 2).cpu().numpy()
    lut = {tuple(r): i for i, r in enumerate(ref_c.tolist())}
    perm = torch.tensor([lut[tuple(r)] for r in got_c.tolist()])
    assert _rel_err(got, ref[perm]) < CONV_TOL


@pytest.mark.parametrize("cin,cout", [(8, 64), (64, 32), (32, 16), (16, 8), (6, 10)])
def test_convT_k2s2_vs_oracle(cin, cout):
    c = _cloud(cin * 3 + cout, n=1500, size=14, stride=2)
    g = torch.Generator().manual_seed(cout)
    f = torch.randn(len(c), cin, generator=g)
    w = torch.randn(8, cin, cout, generator=g) / np.sqrt(cin)
    b = torch.randn(1, cout, generator=g)
    ref, ref_c = S.convT_k2s2(f, c, 2, w, b)
    got = ops.convT_k2s2(f.to(DEV), w.to(DEV), b.to(DEV), relu=True)
    assert _rel_err(got, torch.relu(ref)) < CONV_TOL
    assert (ops.unpack_keys(ops.upsample_keys(_keys(c, 2)), 1).cpu().numpy() == ref_c).all()


# ------------------------------------------------------------------ selection / pruning

@pytest.mark.parametrize("n,k", [(1, 1), (1000, 1), (1000, 999), (100_003, 37_111), (5, 0), (5, 9), (1_700_000, 800_000)])
def test_topk_mask_matches_torch_topk(n, k):
    g = torch.Generator().manual_seed(n + k)
    x = torch.randn(n, 1, generator=g) * 5
    mask = ops.topk_mask(x.to(DEV), k).cpu().numpy()
    assert (mask == S.topk_mask(x, k)).all()


def test_topk_mask_ties_and_specials():
    x = torch.tensor([1.0, 2.0, 2.0, 2.0, -0.0, 0.0, -3.0, 2.0, float("inf"), -float("inf")])
    for k in range(0, 11):
        m = ops.topk_mask(x.to(DEV), k).cpu().numpy()
        assert m.sum() == min(k, 10)
        kept, dropped = x[torch.from_numpy(m)], x[torch.from_numpy(~m)]
        if len(kept) and len(dropped):
            assert kept.min() >= dropped.max()
    m = ops.topk_mask(x.to(DEV), 3).cpu().numpy()              # ties at the threshold: lowest rows win
    assert m.nonzero()[0].tolist() == [1, 2, 8]


@pytest.mark.parametrize("channels", [1, 16, 64, 6])
def test_prune_stable_compaction(channels):
    c = _cloud(9, n=5000, size=30)
    keys = _keys(c)
    g = torch.Generator().manual_seed(channels)
    f = torch.randn(len(c), channels, generator=g)
    mask = torch.rand(len(c), generator=g) < 0.4
    k_out, f_out = ops.prune(mask.to(DEV), keys, f.to(DEV))
    f_ref, c_ref = S.prune(f, c, mask.numpy())
    assert (ops.unpack_keys(k_out, 1).cpu().numpy() == c_ref).all()
    assert torch.equal(f_out.cpu(), f_ref)
    for m in (torch.zeros(len(c), dtype=torch.bool), torch.ones(len(c), dtype=torch.bool)):
        k2, f2 = ops.prune(m.to(DEV), keys, f.to(DEV))
        assert len(k2) == int(m.sum()) == len(f2)


# ------------------------------------------------------------------ entropy bottleneck

@pytest.mark.parametrize("name", ["r3", "r7"])
def test_eb_likelihood_and_tables_vs_reference_golden(name):
    gold = np.load(os.path.join(GOLDEN, f"entropy_{name}.npz"))
    sd = load_ckpt(name)
    p = entropy_ref.params_from_state_dict(sd)
    params = ops.pack_eb_params(p["matrices"], p["biases"], p["factors"], DEV)
    lik = ops.eb_likelihood(torch.from_numpy(gold["values"]).to(DEV), params).cpu().numpy()
    np.testing.assert_allclose(lik, gold["likelihood"], rtol=2e-5, atol=1e-9)
    for key in gold.files:
        if not key.startswith("cdf_"):
            continue
        _, lo, hi = key.split("_")
        cdf, u16 = ops.eb_cdf_table(params, int(lo), int(hi))
        np.testing.assert_allclose(cdf.cpu().numpy(), gold[key], rtol=0, atol=2e-6)
        ref_u16 = rangecoder_ref.cdf_float_to_u16(gold[key]).astype(np.int64)
        got_u16 = u16.cpu().numpy().view(np.uint16).astype(np.int64)
        assert np.abs(got_u16[:, :-1] - ref_u16[:, :-1]).max() <= 1           # integer table within 1 count
        assert (np.diff(got_u16[:, :-1], axis=1) > 0).all()                   # strictly increasing rows


def test_eb_quantize_symbols():
    g = torch.Generator().manual_seed(3)
    f = torch.randn(1521, 8, generator=g) * 3
    f[0, 0], f[1, 1], f[2, 2] = 0.5, 1.5, -2.5                               # half-to-even cases
    sym, lo, hi = ops.eb_quantize(f.to(DEV))
    ref_sym, ref_lo, ref_hi = entropy_ref.quantize_symbols(f)
    assert (lo, hi) == (int(ref_lo), int(ref_hi))
    assert torch.equal(sym.cpu(), ref_sym)


def test_symbol_ranges_on_the_device_code_the_same_stream():
    """section 8b pcgc_symbol_ranges: per-symbol (c_low, c_high) looked up on the GPU == the host coder's table walk; the
    stream coded from them is byte-identical (and == the oracle's); a symbol outside the alphabet raises."""
    gold = np.load(os.path.join(GOLDEN, "entropy_r3.npz"))
    key = [k for k in gold.files if k.startswith("cdf_")][0]
    tab = rangecoder_ref.cdf_float_to_u16(gold[key]).astype(np.uint16)
    C, lp = tab.shape
    rng = np.random.default_rng(9)
    sym = rng.integers(0, lp - 1, size=(13784, C)).astype(np.int16)
    ranges = ops.symbol_ranges(torch.from_numpy(sym).to(DEV), torch.from_numpy(tab.view(np.int16)).to(DEV)).cpu().numpy()
    flat, rows = sym.reshape(-1), np.arange(sym.size) % C
    lo = tab[rows, flat].astype(np.uint32)
    hi = np.where(flat == lp - 2, 65536, tab[rows, np.minimum(flat + 1, lp - 1)]).astype(np.uint32)
    assert np.array_equal(ranges.view(np.uint32), lo | ((hi - 1) << 16))
    stream = ops.rc_encode_ranges(ranges)
    assert stream == ops.rc_encode_u16(tab, sym) == rangecoder_ref.encode_u16(tab, rows.astype(np.int32), flat)
    assert np.array_equal(ops.rc_decode_u16(tab, stream, sym.size), flat)
    bad = sym.copy()
    bad[7, 3] = lp - 1
    with pytest.raises(ValueError):
        ops.symbol_ranges(torch.from_numpy(bad).to(DEV), torch.from_numpy(tab.view(np.int16)).to(DEV))


def test_first_layer_on_constant_one_features_from_the_parent_map():
    """encoder.conv0 fused with its neighbourhood lookup (csrc/conv_ones.cu): out = relu(bias + sum of the weights of the present
    neighbours), presence read off the PARENT's kernel map -- == the oracle's k=3 convolution of an all-ones feature column, in
    fp32 and in the h2 copy; ragged sizes (warp tails), isolated voxels, a single voxel."""
    g = torch.Generator().manual_seed(5)
    w = torch.randn(27, 1, 16, generator=g) / np.sqrt(27)
    b = torch.randn(1, 16, generator=g)
    for n in (20011, 33, 32, 1):
        c = _surface()[:n]
        if n == 33:
            c = np.concatenate([c, np.array([[0, 500, 3, 7], [0, 2, 900, 901]], dtype=c.dtype)])      # far-away singletons
        keys, _ = ops.argsort_u64(_keys(c))
        cs = ops.unpack_keys(keys, 1).cpu().numpy()
        pk, rows, off, parent_of = ops.stride_down(keys, keys_are_sorted=True, with_parent_of=True)
        info = ops.parent_info(keys, off)
        pnbr = ops.kernel_map_k3(pk, ops.HashTable(pk))
        flag = torch.zeros(1, dtype=torch.int32, device=DEV)
        got, got_h = ops.conv_k3_ones_from_parent(pnbr, keys, parent_of, info, w.to(DEV), b.to(DEV), relu=True, want_h2=True, overflow=flag)
        ref = torch.relu(S.conv_k3(torch.ones(len(cs), 1), cs, 1, w, b))
        assert _rel_err(got, ref) < 1e-6 and _rel_err(ops.join_h2(got_h), ref) < H2_TOL and int(flag.item()) == 0
        only_h = ops.conv_k3_ones_from_parent(pnbr, keys, parent_of, info, w.to(DEV), b.to(DEV), relu=False, want_f32=False, want_h2=True)
        assert only_h[0] is None and _rel_err(ops.join_h2(only_h[1]), S.conv_k3(torch.ones(len(cs), 1), cs, 1, w, b)) < H2_TOL
