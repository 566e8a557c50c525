"""CPU-side checks of the drop-in boundary: checkpoints load strictly into the shim-built network; where
the reference sources are present (build container) its UNCHANGED modules import over the shims and its
own EntropyBottleneck.compress/decompress run on our torchac (host range coder, no GPU needed)."""
import io
import os
import sys
import zipfile

import numpy as np
import pytest
import torch

import pcgcv2_b200
from oracle import entropy_ref, rangecoder_ref
from util import load_ckpt

REF = "/root/reference"
has_ref = os.path.exists(os.path.join(REF, "pcc_model.py"))


def test_checkpoint_loads_strictly_into_shim_model():
    from pcgcv2_b200.model import PCCModel, load_model
    sd = load_ckpt("r3")
    model = load_model(sd, device="cpu")
    names = dict(model.named_parameters())
    assert len(names) == 224 and set(names) == set(sd)                     # 227 reference entries minus 3 aliases
    assert names["encoder.block0.0.conv1_0.kernel"].shape == (32, 8)        # k=1 kernels are 2-D
    assert names["decoder.up0.kernel"].shape == (8, 8, 64) and names["encoder.conv0.bias"].shape == (1, 16)
    with pytest.raises(RuntimeError):
        PCCModel().load_state_dict({**sd, "encoder.conv0.extra": torch.zeros(1)}, strict=True)


def test_conv_module_refuses_cpu_and_unsupported_kernels():
    pcgcv2_b200.install_shims()
    import MinkowskiEngine as ME
    with pytest.raises(ValueError):
        ME.SparseTensor(features=torch.ones(2, 1), coordinates=torch.zeros((2, 4), dtype=torch.int32))
    with pytest.raises(NotImplementedError):
        ME.MinkowskiConvolution(in_channels=4, out_channels=4, kernel_size=5, stride=1, bias=True, dimension=3)
    c, f = ME.utils.sparse_collate([torch.zeros((3, 3)), np.ones((2, 3))], [torch.ones(3, 1), np.ones((2, 1))])
    assert c.dtype == torch.int32 and c[:, 0].tolist() == [0, 0, 0, 1, 1] and f.shape == (5, 1)


@pytest.mark.skipif(not has_ref, reason="reference sources not present")
def test_reference_modules_import_unchanged_over_the_shims():
    pcgcv2_b200.install_shims()
    sys.path.append(REF)                                  # shims stay ahead of the reference directory
    try:
        for m in ("pcc_model", "autoencoder", "entropy_model", "data_utils"):
            sys.modules.pop(m, None)
        import pcc_model
        assert pcc_model.ME.__version__.endswith("pcgc.b200")
        model = pcc_model.PCCModel()
        with zipfile.ZipFile(os.path.join(REF, "ckpts.zip")) as z:
            ckpt = torch.load(io.BytesIO(z.read("ckpts/r3_0.10bpp.pth")), map_location="cpu")
        model.load_state_dict(ckpt["model"])               # strict, exactly as coder.py:142
        # the reference's own compress()/decompress() on our torchac: bytes equal the oracle's
        eb = model.entropy_bottleneck
        g = torch.Generator().manual_seed(0)
        feats = torch.randn(1521, 8, generator=g) * 2
        strings, min_v, max_v = eb.compress(feats)
        params = entropy_ref.params_from_state_dict(ckpt["model"])
        sym, lo, hi = entropy_ref.quantize_symbols(feats)
        assert (float(min_v[0]), float(max_v[0])) == (lo, hi)
        table = rangecoder_ref.cdf_float_to_u16(entropy_ref.cdf_table(params, lo, hi, 8).numpy())
        rows = np.tile(np.arange(8, dtype=np.int32), len(feats))
        assert strings == rangecoder_ref.encode_u16(table, rows, sym.numpy().reshape(-1))
        back = eb.decompress(strings, min_v[0], max_v[0], feats.shape, channels=8)
        assert torch.equal(back, feats.round())
    finally:
        sys.path.remove(REF)
        for m in ("pcc_model", "autoencoder", "entropy_model"):
            sys.modules.pop(m, None)
