"""The PCGCv2 network expressed over the drop-in ``MinkowskiEngine`` operator surface.

This is the consumer side of the boundary: the same topology, module names and parameter
names as the reference's ``pcc_model.py:8-16`` / ``autoencoder.py`` so that a reference
checkpoint loads with ``strict=True``.  It exists because the reference's own files cannot
travel to the GPU box; where they are available (the build container) the tests also load
them UNCHANGED on top of the shims (``tests/test_shim_plumbing.py``).

The layer plan is data (``ENCODER_PLAN`` / ``DECODER_PLAN``), the modules are built from it.
"""
from __future__ import annotations

import torch

import pcgcv2_b200
from pcgcv2_b200 import ops

pcgcv2_b200.install_shims()
import MinkowskiEngine as ME  # noqa: E402
from data_utils import isin, istopk  # noqa: E402

ENC_CHANNELS = (1, 16, 32, 64, 32, 8)       # pcc_model.py:11
DEC_CHANNELS = (8, 64, 32, 16)              # pcc_model.py:12
IRN_PER_STAGE = 3


def _conv(cin, cout, k, s=1, transpose=False):
    cls = ME.MinkowskiGenerativeConvolutionTranspose if transpose else ME.MinkowskiConvolution
    return cls(in_channels=cin, out_channels=cout, kernel_size=k, stride=s, bias=True, dimension=3)


class InceptionResNet(torch.nn.Module):
    """two-branch residual block (autoencoder.py:7-57): k3 C->C/4->C/2 and k1-k3-k1 C->C/4->C/4->C/2."""

    def __init__(self, channels):
        super().__init__()
        q, h = channels // 4, channels // 2
        for name, (cin, cout, k) in {"conv0_0": (channels, q, 3), "conv0_1": (q, h, 3), "conv1_0": (channels, q, 1),
                                     "conv1_1": (q, q, 3), "conv1_2": (q, h, 1)}.items():
            setattr(self, name, _conv(cin, cout, k))
        self.relu = ME.MinkowskiReLU(inplace=True)

    def forward(self, x):
        left = self.conv0_1(self.relu(self.conv0_0(x)))
        right = self.conv1_2(self.relu(self.conv1_1(self.relu(self.conv1_0(x)))))
        return ME.cat(left, right) + x


def _stage(channels):
    return torch.nn.Sequential(*[InceptionResNet(channels) for _ in range(IRN_PER_STAGE)])


class Encoder(torch.nn.Module):
    """autoencoder.py:68-147: conv0, then 3 x [k2s2 down, 3 IRN, k3 conv]."""

    def __init__(self, channels=ENC_CHANNELS):
        super().__init__()
        c = channels
        self.conv0 = _conv(c[0], c[1], 3)
        widths = [(c[1], c[2]), (c[2], c[3]), (c[3], c[4])]
        for i, (cin, cout) in enumerate(widths):
            setattr(self, f"down{i}", _conv(cin, cout, 2, 2))
            setattr(self, f"block{i}", _stage(cout))
            setattr(self, f"conv{i + 1}", _conv(cout, cout if i < 2 else c[5], 3))
        self.relu = ME.MinkowskiReLU(inplace=True)

    def forward(self, x):
        outs = []
        x = self.relu(self.conv0(x))
        for i in range(3):
            x = getattr(self, f"block{i}")(self.relu(getattr(self, f"down{i}")(x)))
            outs.append(x)
            x = getattr(self, f"conv{i + 1}")(x)
            if i < 2:
                x = self.relu(x)
        return [x, outs[1], outs[0]]


class Decoder(torch.nn.Module):
    """autoencoder.py:150-273: 3 x [generative k2s2 up, k3 conv, 3 IRN, 1-channel k3 classifier, top-k prune]."""

    def __init__(self, channels=DEC_CHANNELS):
        super().__init__()
        for i in range(3):
            setattr(self, f"up{i}", _conv(channels[i], channels[i + 1], 2, 2, transpose=True))
            setattr(self, f"conv{i}", _conv(channels[i + 1], channels[i + 1], 3))
            setattr(self, f"block{i}", _stage(channels[i + 1]))
            setattr(self, f"conv{i}_cls", _conv(channels[i + 1], 1, 3))
        self.relu = ME.MinkowskiReLU(inplace=True)
        self.pruning = ME.MinkowskiPruning()

    def prune_voxel(self, data, data_cls, nums, ground_truth, training):
        mask = istopk(data_cls, nums)
        if training:
            assert ground_truth is not None
            mask = mask + isin(data_cls.C, ground_truth.C)
        return self.pruning(data, mask.to(data.device))

    def forward(self, x, nums_list, ground_truth_list, training=True):
        out, cls_list = x, []
        for i in range(3):
            out = self.relu(getattr(self, f"conv{i}")(self.relu(getattr(self, f"up{i}")(out))))
            out = getattr(self, f"block{i}")(out)
            cls = getattr(self, f"conv{i}_cls")(out)
            cls_list.append(cls)
            out = self.prune_voxel(out, cls, nums_list[i], ground_truth_list[i], training)
        return cls_list, out


class _LikelihoodFn(torch.autograd.Function):
    """EntropyBottleneck._likelihood (entropy_model.py:112-130) and its gradient on libpcgc kernels."""

    @staticmethod
    def forward(ctx, values, *params):
        packed = ops.pack_eb_params(params[0:4], params[4:8], params[8:12], values.device)
        ctx.save_for_backward(values, packed)
        return ops.eb_likelihood(values, packed)

    @staticmethod
    def backward(ctx, g):
        values, packed = ctx.saved_tensors
        gv, gp = ops.eb_likelihood_bwd(values, packed, g, need_values_grad=ctx.needs_input_grad[0])
        m, b, f = ops.unpack_eb_param_grads(gp)
        return (gv, *m, *b, *f)


class _RoundNoGradient(torch.autograd.Function):          # entropy_model.py:9-17
    @staticmethod
    def forward(ctx, x):
        return x.round()

    @staticmethod
    def backward(ctx, g):
        return g


class _LowBound(torch.autograd.Function):                 # entropy_model.py:20-39
    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return torch.clamp(x, min=1e-9)

    @staticmethod
    def backward(ctx, g):
        x, = ctx.saved_tensors
        return g * ((x >= 1e-9) | (g < 0.0)).to(g.dtype)


class EntropyBottleneck(torch.nn.Module):
    """Factorised-prior bottleneck with the reference's parameter names (entropy_model.py:58-80) so checkpoints
    load strictly; likelihoods, their gradients and the codec tables are evaluated by libpcgc (ops.eb_*)."""

    def __init__(self, channels=8, filters=(3, 3, 3), init_scale=8):
        super().__init__()
        f = (1,) + tuple(filters) + (1,)
        scale = float(init_scale) ** (1 / (len(filters) + 1))
        mats, biases, factors = [], [], []
        for i in range(4):
            init = float(torch.log(torch.expm1(torch.tensor(1.0 / scale / f[i + 1]))))
            mats.append(torch.nn.Parameter(torch.full((channels, f[i + 1], f[i]), init)))
            biases.append(torch.nn.Parameter(torch.rand(channels, f[i + 1], 1) - 0.5))
            factors.append(torch.nn.Parameter(torch.zeros(channels, f[i + 1], 1)))
        self._matrices = torch.nn.ParameterList(mats)
        self._biases = torch.nn.ParameterList(biases)
        self._factors = torch.nn.ParameterList(factors)

    def likelihood(self, values):
        return _LikelihoodFn.apply(values, *self._matrices, *self._biases, *self._factors)

    def forward(self, inputs, quantize_mode="noise"):
        """entropy_model.py:132-140: quantise (uniform noise / straight-through round), likelihood, lower bound."""
        if quantize_mode == "noise":
            outputs = inputs + (torch.rand_like(inputs) - 0.5)
        elif quantize_mode == "symbols":
            outputs = _RoundNoGradient.apply(inputs)
        else:
            outputs = inputs
        return outputs, _LowBound.apply(self.likelihood(outputs))


class PCCModel(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.encoder = Encoder()
        self.decoder = Decoder()
        self.entropy_bottleneck = EntropyBottleneck(ENC_CHANNELS[-1])

    def forward(self, x, training=True, quantize_mode=None):
        """pcc_model.py:26-45: analysis, bottleneck likelihood, synthesis with (top-k | ground truth) pruning."""
        y_list = self.encoder(x)
        y = y_list[0]
        ground_truth_list = y_list[1:] + [x]
        nums_list = [[len(c) for c in gt.decomposed_coordinates] for gt in ground_truth_list]
        mode = quantize_mode or ("noise" if training else "symbols")
        y_f, likelihood = self.entropy_bottleneck(y.F, quantize_mode=mode)
        y_q = ME.SparseTensor(features=y_f, coordinate_map_key=y.coordinate_map_key,
                              coordinate_manager=y.coordinate_manager, device=y.device)
        out_cls_list, out = self.decoder(y_q, nums_list, ground_truth_list, training)
        return {"out": out, "out_cls_list": out_cls_list, "prior": y_q, "likelihood": likelihood,
                "ground_truth_list": ground_truth_list}


def load_model(state_dict, device="cuda") -> PCCModel:
    """reference checkpoint (``ckpt['model']``) -> PCCModel on ``device`` (strict apart from the three
    alias entries ``entropy_bottleneck.{matrix,bias,factor}`` the reference's module duplicates)."""
    sd = {k: v for k, v in state_dict.items() if k not in ("entropy_bottleneck.matrix", "entropy_bottleneck.bias",
                                                             "entropy_bottleneck.factor")}
    model = PCCModel()
    model.load_state_dict(sd, strict=True)
    return model.to(device).eval()
