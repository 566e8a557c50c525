"""CPU restatement of the factorised-prior entropy model (``entropy_model.py``).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  PINNED: golden outputs of the
reference module itself (imported in the build container with a stub
``torchac``) are frozen in ``tests/golden/entropy_*.npz`` by
``tests/golden/make_golden.py`` and checked in ``tests/test_oracle.py``.

``params`` is a dict with lists ``matrices`` / ``biases`` / ``factors`` of 4
float32 tensors each, shapes ``(C,3,1),(C,3,3),(C,3,3),(C,1,3)`` /
``(C,3,1)x3,(C,1,1)`` / same (``entropy_model.py:58-80``).
"""
from __future__ import annotations

import numpy as np
import torch

LIKELIHOOD_BOUND = 1e-9          # entropy_model.py:53


def params_from_state_dict(sd, prefix="entropy_bottleneck."):
    g = lambda name: [sd[f"{prefix}{name}.{i}"].detach().float().cpu() for i in range(4)]
    return {"matrices": g("_matrices"), "biases": g("_biases"), "factors": g("_factors")}


def logits_cumulative(params, x: torch.Tensor) -> torch.Tensor:
    """``entropy_model.py:82-101``; x is ``[C, 1, n]``."""
    logits = x
    for i in range(4):
        m = torch.nn.functional.softplus(params["matrices"][i])
        logits = torch.matmul(m, logits)
        logits = logits + params["biases"][i]
        f = torch.tanh(params["factors"][i])
        logits = logits + f * torch.tanh(logits)
    return logits


def likelihood(params, values: torch.Tensor) -> torch.Tensor:
    """``entropy_model.py:112-130``; values ``[n, C]`` -> likelihood ``[n, C]``
    (not yet lower-bounded)."""
    x = values.float().permute(1, 0).contiguous()
    shape = x.shape
    x = x.view(shape[0], 1, -1)
    lower = logits_cumulative(params, x - 0.5)
    upper = logits_cumulative(params, x + 0.5)
    sign = -torch.sign(lower + upper)
    lik = torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))
    return lik.view(shape).permute(1, 0)


def cdf_table(params, min_v: float, max_v: float, channels: int) -> torch.Tensor:
    """Float CDF table ``[C, L+1]`` exactly as ``compress``/``decompress`` build
    it before tiling (``entropy_model.py:155-171,181-189``): symbols grid ->
    likelihood -> clamp(min=1e-9) -> cumsum -> prepend 0 -> clamp(max=1)."""
    symbols = torch.arange(float(min_v), float(max_v) + 1)
    symbols = symbols.reshape(-1, 1).repeat(1, channels)
    pmf = likelihood(params, symbols)
    pmf = torch.clamp(pmf, min=LIKELIHOOD_BOUND).permute(1, 0)
    cdf = pmf.cumsum(dim=-1)
    cdf = torch.cat([torch.zeros(pmf.shape[:-1] + (1,), dtype=pmf.dtype), cdf], dim=-1)
    return cdf.clamp(max=1.0)


def quantize_symbols(feats: torch.Tensor):
    """``entropy_model.py:152-163``: round, min/max, int16 symbols."""
    values = feats.float().round()
    min_v = values.min().float()
    max_v = values.max().float()
    sym = (values - min_v).to(torch.int16)
    return sym, float(min_v), float(max_v)


def ideal_bits(params, values: torch.Tensor) -> float:
    """``loss.py:17-20`` on the lower-bounded likelihood."""
    lik = torch.clamp(likelihood(params, values), min=LIKELIHOOD_BOUND)
    return float(-torch.log2(lik).sum())
