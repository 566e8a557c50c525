"""Backward passes of the drop-in operators against torch autograd through the CPU
oracle (whose operators are differentiable torch ops).  Tolerance: max|d| <= 5e-5 * max|ref|."""
import numpy as np
import pytest
import torch

import pcgcv2_b200
from oracle import sparse_ref as S
from pcgcv2_b200 import ops, synth
from util import with_batch

pytestmark = pytest.mark.gpu
TOL = 5e-5


def _rel(got, ref):
    return float((got.cpu() - ref).abs().max() / ref.abs().max().clamp_min(1e-30))


def _cloud(seed, n=2500, size=14, stride=1):
    rng = np.random.default_rng(seed)
    pts = np.unique(rng.integers(0, size, size=(n, 3)), axis=0)
    rng.shuffle(pts)
    return with_batch(pts * stride)


@pytest.fixture(scope="module")
def ME():
    pcgcv2_b200.install_shims()
    import MinkowskiEngine
    return MinkowskiEngine


def _sparse(ME, c, f, stride=1):
    return ME.SparseTensor(features=f, coordinates=torch.from_numpy(c), tensor_stride=stride, device="cuda")


@pytest.mark.parametrize("cin,cout,k", [(16, 16, 3), (32, 8, 3), (1, 16, 3), (4, 8, 3), (64, 64, 3), (16, 1, 3),
                                        (32, 8, 1), (8, 16, 1), (6, 10, 3)])
def test_conv_stride1_backward(ME, cin, cout, k):
    c = _cloud(cin + cout)
    g = torch.Generator().manual_seed(cin * cout + k)
    f = torch.randn(len(c), cin, generator=g)
    conv = ME.MinkowskiConvolution(in_channels=cin, out_channels=cout, kernel_size=k, stride=1, bias=True, dimension=3).cuda()
    w, b = conv.kernel.detach().cpu().clone().requires_grad_(), conv.bias.detach().cpu().clone().requires_grad_()
    fr = f.clone().requires_grad_()
    ref = S.conv_k3(fr, c, 1, w, b) if k == 3 else S.conv_k1(fr, w, b)
    probe = torch.randn(ref.shape, generator=g)
    (ref * probe).sum().backward()
    fx = f.cuda().requires_grad_()
    out = conv(_sparse(ME, c, fx))
    (out.F * probe.cuda()).sum().backward()
    assert _rel(out.F.detach(), ref.detach()) < TOL
    assert _rel(fx.grad, fr.grad) < TOL
    assert _rel(conv.kernel.grad, w.grad) < TOL
    assert _rel(conv.bias.grad, b.grad) < TOL


@pytest.mark.parametrize("cin,cout", [(16, 32), (64, 32), (6, 10)])
def test_down_conv_backward(ME, cin, cout):
    c = _cloud(3 + cin, n=4000, size=20)
    g = torch.Generator().manual_seed(cin)
    f = torch.randn(len(c), cin, generator=g)
    conv = ME.MinkowskiConvolution(in_channels=cin, out_channels=cout, kernel_size=2, stride=2, bias=True, dimension=3).cuda()
    w, b = conv.kernel.detach().cpu().clone().requires_grad_(), conv.bias.detach().cpu().clone().requires_grad_()
    fr = f.clone().requires_grad_()
    ref, ref_c = S.conv_k2s2(fr, c, 1, w, b)
    fx = f.cuda().requires_grad_()
    out = conv(_sparse(ME, c, fx))
    lut = {tuple(r): i for i, r in enumerate(ref_c.tolist())}
    perm = torch.tensor([lut[tuple(r)] for r in out.C.cpu().numpy().tolist()])
    probe = torch.randn(ref.shape, generator=g)
    (ref * probe).sum().backward()
    (out.F * probe[perm].cuda()).sum().backward()
    assert _rel(out.F.detach(), ref.detach()[perm]) < TOL
    assert _rel(fx.grad, fr.grad) < TOL and _rel(conv.kernel.grad, w.grad) < TOL and _rel(conv.bias.grad, b.grad) < TOL


@pytest.mark.parametrize("cin,cout", [(8, 64), (32, 16), (6, 10)])
def test_generative_up_conv_backward(ME, cin, cout):
    c = _cloud(9 + cin, n=900, size=10, stride=2)
    g = torch.Generator().manual_seed(cout)
    f = torch.randn(len(c), cin, generator=g)
    conv = ME.MinkowskiGenerativeConvolutionTranspose(in_channels=cin, out_channels=cout, kernel_size=2, stride=2,
                                                      bias=True, dimension=3).cuda()
    w, b = conv.kernel.detach().cpu().clone().requires_grad_(), conv.bias.detach().cpu().clone().requires_grad_()
    fr = f.clone().requires_grad_()
    ref, ref_c = S.convT_k2s2(fr, c, 2, w, b)
    fx = f.cuda().requires_grad_()
    out = conv(_sparse(ME, c, fx, stride=2))
    assert (out.C.cpu().numpy() == ref_c).all()
    probe = torch.randn(ref.shape, generator=g)
    (ref * probe).sum().backward()
    (out.F * probe.cuda()).sum().backward()
    assert _rel(fx.grad, fr.grad) < TOL and _rel(conv.kernel.grad, w.grad) < TOL and _rel(conv.bias.grad, b.grad) < TOL


def test_prune_relu_cat_add_backward(ME):
    c = _cloud(21)
    g = torch.Generator().manual_seed(2)
    f = torch.randn(len(c), 8, generator=g)
    mask = torch.rand(len(c), generator=g) < 0.5
    fx = f.cuda().requires_grad_()
    x = _sparse(ME, c, fx)
    y = ME.cat(ME.MinkowskiReLU()(x), x) + ME.cat(x, x)
    out = ME.MinkowskiPruning()(y, mask.cuda())
    probe = torch.randn(int(mask.sum()), 16, generator=g)
    (out.F * probe.cuda()).sum().backward()
    fr = f.clone().requires_grad_()
    ref = (torch.cat([torch.relu(fr), fr], 1) + torch.cat([fr, fr], 1))[mask]
    (ref * probe).sum().backward()
    assert _rel(out.F.detach(), ref.detach()) < TOL and _rel(fx.grad, fr.grad) < TOL


def test_inception_block_backward_through_the_model(ME):
    """gradients of every parameter of one InceptionResNet block + a down conv, end to end."""
    from pcgcv2_b200.model import InceptionResNet
    from oracle import codec_ref
    c = _cloud(33, n=3000, size=16)
    g = torch.Generator().manual_seed(5)
    f = torch.randn(len(c), 16, generator=g)
    blk = InceptionResNet(16).cuda()
    sd = {"b." + k: v.detach().cpu().clone().requires_grad_() for k, v in blk.named_parameters()}
    fr = f.clone().requires_grad_()
    ref = codec_ref._irn(sd, "b", fr, c, 1, codec_ref._Maps())
    probe = torch.randn(ref.shape, generator=g)
    (ref * probe).sum().backward()
    fx = f.cuda().requires_grad_()
    out = blk(_sparse(ME, c, fx))
    (out.F * probe.cuda()).sum().backward()
    assert _rel(out.F.detach(), ref.detach()) < TOL and _rel(fx.grad, fr.grad) < TOL
    for k, v in blk.named_parameters():
        assert _rel(v.grad, sd["b." + k].grad) < TOL, k


@pytest.mark.parametrize("name", ["r3", "r7"])
def test_entropy_bottleneck_likelihood_backward(name):
    """pcgc_eb_likelihood_bwd vs autograd through the oracle (= the reference's own torch code path)."""
    from oracle import entropy_ref
    from pcgcv2_b200.model import EntropyBottleneck
    from util import load_ckpt
    sd = load_ckpt(name)
    eb = EntropyBottleneck(8)
    eb.load_state_dict({k[len("entropy_bottleneck."):]: v for k, v in sd.items() if k.startswith("entropy_bottleneck.")})
    eb = eb.cuda()
    g = torch.Generator().manual_seed(4)
    vals = torch.randn(3001, 8, generator=g) * 3
    probe = torch.rand(3001, 8, generator=g) + 0.1
    ref_params = {"matrices": [p.detach().cpu().clone().requires_grad_() for p in eb._matrices],
                  "biases": [p.detach().cpu().clone().requires_grad_() for p in eb._biases],
                  "factors": [p.detach().cpu().clone().requires_grad_() for p in eb._factors]}
    vr = vals.clone().requires_grad_()
    (entropy_ref.likelihood(ref_params, vr) * probe).sum().backward()
    vx = vals.cuda().requires_grad_()
    (eb.likelihood(vx) * probe.cuda()).sum().backward()
    assert _rel(vx.grad, vr.grad) < 1e-4
    for kind, plist in (("matrices", eb._matrices), ("biases", eb._biases), ("factors", eb._factors)):
        for i, p in enumerate(plist):
            ref = ref_params[kind][i].grad
            assert float((p.grad.cpu() - ref).abs().max()) <= 1e-4 * float(ref.abs().max()) + 1e-6, (kind, i)


def test_training_step_matches_oracle_and_learns(ME):
    """config-5 flavour: PCCModel.forward(training) + BCE/bits losses + backward on a 64^3 crop."""
    from oracle import codec_ref
    from pcgcv2_b200.model import load_model
    from util import load_ckpt
    rng = np.random.default_rng(3)
    u = rng.normal(size=(30000, 3)); u /= np.linalg.norm(u, axis=1, keepdims=True)
    pts = np.unique(np.round(32 + u * np.array([20, 14, 26])).astype(np.int32), axis=0)       # ellipsoid shell in 64^3
    coords = with_batch(pts)
    sd = load_ckpt("r3")
    model = load_model(sd).train()
    x = ME.SparseTensor(features=torch.ones(len(coords), 1), coordinates=torch.from_numpy(coords), device="cuda")
    out = model(x, training=True, quantize_mode="symbols")
    crit = torch.nn.BCEWithLogitsLoss()
    from data_utils import isin
    bce = 0
    for cls, gt in zip(out["out_cls_list"], out["ground_truth_list"]):
        bce = bce + crit(cls.F.squeeze(), isin(cls.C, gt.C).float()) / np.log(2.0)             # loss.py:7-15, trainer.py:129
    bpp = -torch.log2(out["likelihood"]).sum() / float(len(x))
    loss = bce + bpp
    loss.backward()
    sdr = {k: v.clone().requires_grad_() for k, v in sd.items()}
    ref_loss, ref_bce, ref_bpp = codec_ref.train_forward(sdr, coords)
    ref_loss.backward()
    assert abs(float(bce) - float(ref_bce)) < 1e-3 * float(ref_bce) and abs(float(bpp) - float(ref_bpp)) < 1e-3 * float(ref_bpp) + 1e-6
    checked = 0
    for k, p in model.named_parameters():
        ref = sdr[k].grad
        if ref is None or float(ref.abs().max()) < 1e-6:
            continue
        assert float((p.grad.cpu() - ref).abs().max()) <= 2e-3 * float(ref.abs().max()), k
        checked += 1
    assert checked > 150
    # and one optimiser step lowers the loss on the same batch
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)
    opt.step()
    with torch.no_grad():
        out2 = model(x, training=True, quantize_mode="symbols")
        bce2 = sum(crit(cls.F.squeeze(), isin(cls.C, gt.C).float()) / np.log(2.0)
                   for cls, gt in zip(out2["out_cls_list"], out2["ground_truth_list"]))
        loss2 = bce2 - torch.log2(out2["likelihood"]).sum() / float(len(x))
    assert float(loss2) < float(loss)


def test_fused_bce_isin_matches_reference_loss_and_is_deterministic(ME):
    """row f4: get_bce (one fused pcgc_bce_isin pass) == loss.py:7-15 evaluated with torch (BCEWithLogitsLoss over the
    isin mask, / ln 2, x rows) in value and gradient; batch index is part of the membership test; the reduction is
    bit-reproducible; strided logits; empty candidate set."""
    from pcgcv2_b200 import train
    from data_utils import isin
    rng = np.random.default_rng(11)
    cand = np.unique(rng.integers(0, 24, size=(30000, 4)) % np.array([3, 24, 24, 24]), axis=0).astype(np.int32)
    gt = cand[rng.random(len(cand)) < 0.3].copy()
    gt[:50, 0] = (gt[:50, 0] + 1) % 3                                                      # same xyz, other batch item
    gt = np.unique(gt, axis=0)
    g = torch.Generator().manual_seed(3)
    logits = (torch.randn(len(cand), 1, generator=g) * 6).cuda().requires_grad_()            # saturating values included
    data = ME.SparseTensor(features=logits, coordinates=torch.from_numpy(cand), device="cuda")
    truth = ME.SparseTensor(features=torch.ones(len(gt), 1), coordinates=torch.from_numpy(gt), device="cuda")
    got = train.get_bce(data, truth)
    got.backward()
    mask = isin(data.C, truth.C)
    cset = set(map(tuple, gt.tolist()))
    assert mask.cpu().tolist() == [tuple(r) in cset for r in data.C.cpu().numpy().tolist()]
    ref_logits = logits.detach().double().requires_grad_()
    ref = torch.nn.BCEWithLogitsLoss()(ref_logits.squeeze(), mask.double()) / np.log(2.0) * len(cand)
    ref.backward()
    assert abs(float(got) - float(ref)) <= 2e-6 * float(ref)
    assert float((logits.grad.double() - ref_logits.grad).abs().max()) <= 2e-6 * float(ref_logits.grad.abs().max())
    again = train.get_bce(data, truth)
    assert float(again) == float(got)                                                       # fixed-order reduction
    # the reference's own loss.py over the shims (host isin + torch BCE) gives the same number to fp32 rounding
    crit = torch.nn.BCEWithLogitsLoss()(logits.detach().squeeze(), mask.float()) / np.log(2.0) * len(cand)
    assert abs(float(got) - float(crit)) <= 1e-5 * float(crit)
    wide = torch.randn(len(cand), 4, generator=g).cuda()
    l1, _, t1 = ops.bce_isin(wide[:, 2:3], data.coordinate_manager._get(data.coordinate_map_key).keys,
                             truth.coordinate_manager._get(truth.coordinate_map_key).table)
    l2, _, _ = ops.bce_isin(wide[:, 2].contiguous(), data.coordinate_manager._get(data.coordinate_map_key).keys,
                            truth.coordinate_manager._get(truth.coordinate_map_key).table)
    assert float(l1) == float(l2) and t1.bool().tolist() == mask.tolist()
    assert train.get_cls_metrics(mask, mask) == [1.0, 1.0, 1.0]


def test_weight_and_bias_gradients_are_bit_reproducible(ME):
    """a16: the weight / bias gradient kernels add per-block partials in block order -- two runs give identical bits."""
    c = _cloud(77, n=60000, size=48)
    g = torch.Generator().manual_seed(8)
    f = torch.randn(len(c), 16, generator=g).cuda()
    conv = ME.MinkowskiConvolution(in_channels=16, out_channels=32, kernel_size=3, stride=1, bias=True, dimension=3).cuda()
    probe = torch.randn(len(c), 32, generator=g).cuda()
    grads = []
    for _ in range(3):
        conv.zero_grad(set_to_none=True)
        (conv(_sparse(ME, c, f)).F * probe).sum().backward()
        grads.append((conv.kernel.grad.clone(), conv.bias.grad.clone()))
    assert all(torch.equal(grads[0][0], k) and torch.equal(grads[0][1], b) for k, b in grads[1:])
    ref = S.conv_k3(f.cpu().requires_grad_(False), c, 1, conv.kernel.detach().cpu().requires_grad_(), conv.bias.detach().cpu())
    assert _rel(conv(_sparse(ME, c, f)).F.detach(), ref.detach()) < TOL


def test_data_parallel_train_step_single_rank(ME):
    """train_step (flat gradient bucket, fused losses, Adam) lowers the loss on a config-5 style batch of shells."""
    from pcgcv2_b200 import train
    from pcgcv2_b200.model import PCCModel
    torch.manual_seed(0)
    model = PCCModel().cuda().train()
    coords, feats = train.shell_batch(0, batch=4, size=64, rng_points=6000)
    assert coords[:, 0].max() == 3 and 8000 < len(coords) < 40000
    x = ME.SparseTensor(features=torch.from_numpy(feats), coordinates=torch.from_numpy(coords), device="cuda")
    bucket = train.GradBucket(model.parameters())
    opt = torch.optim.Adam(model.parameters(), lr=train.adam_lr(), betas=(0.9, 0.999))
    first = last = None
    for it in range(8):
        loss, bce, bpp = train.train_step(model, opt, bucket, x)
        first = float(loss) if first is None else first
        last = float(loss)
        assert np.isfinite(last)
    assert last < first, (first, last)
    assert model.encoder.conv0.kernel.grad.data_ptr() == bucket.views[model.encoder.conv0.kernel].data_ptr()
