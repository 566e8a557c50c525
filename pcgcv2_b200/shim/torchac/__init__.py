"""Drop-in ``torchac`` (0.9.3 surface used by the reference, entropy_model.py:174,192) backed
by libpcgc's host range coder (``pcgc_rc_encode_host`` / ``pcgc_rc_decode_host``; SURVEY.md
Appendix B).  Same argument meaning and error behaviour: CPU tensors only, ``sym`` int16 with
``cdf_float.shape[:-1] == sym.shape``, symbols in ``[0, Lp-2]``."""
import numpy as np
import torch

from pcgcv2_b200 import ops as _ops

__version__ = "0.9.3+pcgc.b200"


def _check_cdf(cdf_float, needs_normalization):
    if not isinstance(cdf_float, torch.Tensor) or cdf_float.is_cuda:
        raise ValueError("cdf_float must be a CPU tensor")
    if cdf_float.dim() < 2 or cdf_float.shape[-1] < 2:
        raise ValueError("cdf_float must have shape (..., Lp) with Lp >= 2")
    if not needs_normalization:
        raise NotImplementedError("needs_normalization=False is outside the PCGCv2 hot path")
    return np.ascontiguousarray(cdf_float.detach().to(torch.float32).numpy()).reshape(-1, cdf_float.shape[-1])


def encode_float_cdf(cdf_float, sym, needs_normalization=True, check_input_bounds=False) -> bytes:
    table = _check_cdf(cdf_float, needs_normalization)
    if not isinstance(sym, torch.Tensor) or sym.is_cuda:
        raise ValueError("sym must be a CPU tensor")
    if sym.dtype != torch.int16:
        raise ValueError("sym must be int16")
    if tuple(cdf_float.shape[:-1]) != tuple(sym.shape):
        raise ValueError(f"shape mismatch: cdf {tuple(cdf_float.shape)} vs sym {tuple(sym.shape)}")
    if check_input_bounds:
        if float(cdf_float.min()) < 0 or float(cdf_float.max()) > 1:
            raise ValueError("cdf_float values must be in [0, 1]")
        Lp = cdf_float.shape[-1]
        if sym.numel() and (int(sym.min()) < 0 or int(sym.max()) > Lp - 2):
            raise ValueError("sym values must be in [0, Lp - 2]")
    return _ops.rc_encode_float(table, sym.contiguous().numpy())


def decode_float_cdf(cdf_float, byte_stream, needs_normalization=True) -> torch.Tensor:
    table = _check_cdf(cdf_float, needs_normalization)
    out = _ops.rc_decode_float(table, bytes(byte_stream), table.shape[0])
    return torch.from_numpy(out).reshape(tuple(cdf_float.shape[:-1]))
