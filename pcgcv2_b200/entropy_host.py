"""Host half of the EntropyBottleneck coder: the float CDF table the range coder consumes.

In the reference this table is host data: ``FeatureCoder.__init__`` moves the entropy model to the CPU
(coder.py:44) and ``compress`` / ``decompress`` evaluate the ``[C, L+1]`` table there with float32 torch
operators before handing it to torchac (entropy_model.py:151-171,178-189).  A range-coded stream only decodes
when encoder and decoder hold the IDENTICAL 16-bit table, and the reference's float32 CPU evaluation is not the
correctly rounded value (measured: the float64 evaluation of ``pcgc_eb_cdf_table`` differs from it in ~0.1 % of
the entries over the seven shipped checkpoints), so ``Codec`` builds the table it codes with the way the
reference does -- the same float32 torch CPU operators in the same order on the same ``[C, 1, L]`` operands -- and
its ``_F.bin`` is byte-identical to the reference's for identical symbols.  The table is ``C x (L+1)`` values
(48 at r3, 160 at r7), a pure function of the checkpoint and the symbol range, cached per range; it is not on the
per-voxel path.  The per-element likelihood of the training path (a12) runs on the GPU (csrc/entropy.cu).
"""
from __future__ import annotations

import numpy as np
import torch

LIKELIHOOD_BOUND = 1e-9                      # entropy_model.py:53


class HostTable:
    def __init__(self, matrices, biases, factors):
        cpu = lambda ts: [t.detach().to("cpu", torch.float32) for t in ts]
        self.matrices, self.biases, self.factors = cpu(matrices), cpu(biases), cpu(factors)
        self.channels = int(self.matrices[0].shape[0])

    def _logits_cumulative(self, inputs):    # entropy_model.py:82-101
        logits = inputs
        for m, b, f in zip(self.matrices, self.biases, self.factors):
            logits = torch.matmul(torch.nn.functional.softplus(m), logits)
            logits += b
            logits += torch.tanh(f) * torch.tanh(logits)
        return logits

    @torch.no_grad()
    def cdf_float(self, min_v: int, max_v: int) -> np.ndarray:
        """float32 [C, L+1]: symbols grid -> likelihood -> clamp(min=1e-9) -> cumsum -> prepend 0 -> clamp(max=1)
        (entropy_model.py:155-171 / 181-189, before the per-point tiling, which repeats this table N3 times)."""
        symbols = torch.arange(float(min_v), float(max_v) + 1).reshape(-1, 1).repeat(1, self.channels)
        x = symbols.permute(1, 0).contiguous()
        shape = x.size()
        x = x.view(shape[0], 1, -1)
        lower = self._logits_cumulative(x - 0.5)
        upper = self._logits_cumulative(x + 0.5)
        sign = -torch.sign(torch.add(lower, upper))
        pmf = torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower)).view(shape).permute(1, 0)
        pmf = torch.clamp(pmf, min=LIKELIHOOD_BOUND).permute(1, 0)
        cdf = pmf.cumsum(dim=-1)
        cdf = torch.cat([torch.zeros(pmf.shape[:-1] + (1,), dtype=pmf.dtype), cdf], dim=-1).clamp(max=1.)
        return np.ascontiguousarray(cdf.numpy())
