"""Training path of the codec (BASELINE config 5).

* ``get_bce`` / ``get_bits`` / ``get_metrics`` -- same names, arguments and return values as the reference's ``loss.py:7-28``.
  ``get_bce`` is ONE fused pass (``pcgc_bce_isin``, csrc/loss.cu): the candidate keys probe the ground-truth hash table (the
  reference does the membership test on the host: D2H of both coordinate sets + ``np.isin``, data_utils.py:63-75), the stable
  BCE-with-logits is summed in bits and d loss / d logit is written for the backward pass in the same sweep.
* ``GradBucket`` + ``train_step`` -- data-parallel training step (trainer.py:117-136 has no DDP; config 5 asks for 8 ranks): all
  gradients live in one flat buffer (parameters' ``.grad`` are views), split into buckets in backward order; a bucket's NCCL
  all-reduce is launched from the post-accumulate hook of its last parameter, so the decoder's bucket is on the wire while the
  analysis network's backward kernels still run.  The weight / bias gradients are deterministic (csrc/conv_bwd.cu), so
  every rank applies bit-identical updates and the replicas never drift.
"""
from __future__ import annotations

import math

import torch
import torch.distributed as dist

from pcgcv2_b200 import ops


class _BceIsin(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, cand_keys, gt_table):
        loss, grad_unit, _ = ops.bce_isin(logits, cand_keys, gt_table, need_grad=True)
        ctx.save_for_backward(grad_unit)
        ctx.shape = logits.shape
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        grad_unit, = ctx.saved_tensors
        return (grad_unit * g).reshape(ctx.shape), None, None


def _cmap(t):
    return t.coordinate_manager._get(t.coordinate_map_key)


def get_bce(data, groud_truth):
    """loss.py:7-15: sum over the candidate rows of BCE-with-logits(data.F, isin(data.C, groud_truth.C)) / ln 2.
    Both arguments are sparse tensors of the drop-in ``MinkowskiEngine``; their Morton keys and the ground truth's hash
    table are the ones the convolutions' kernel maps already use (no coordinate leaves the device)."""
    return _BceIsin.apply(data.F, _cmap(data).keys, _cmap(groud_truth).table)


def get_bits(likelihood):
    """loss.py:17-20."""
    return -torch.sum(torch.log2(likelihood))


def get_cls_metrics(pred, real):
    """loss.py:30-41 with the four counts taken on the device (one D2H of four integers)."""
    pred, real = pred.bool(), real.bool().to(pred.device)
    tp, fn, fp = [int(v) for v in torch.stack([(pred & real).sum(), (~pred & real).sum(), (pred & ~real).sum()]).tolist()]
    precision = tp / (tp + fp + 1e-7)
    recall = tp / (tp + fn + 1e-7)
    iou = tp / (tp + fp + fn + 1e-7)
    return [round(precision, 4), round(recall, 4), round(iou, 4)]


def get_metrics(data, groud_truth):
    """loss.py:22-28."""
    from data_utils import istopk
    mask_real = _cmap(groud_truth).table.contains(_cmap(data).keys)
    nums = [len(c) for c in groud_truth.decomposed_coordinates]
    mask_pred = istopk(data, nums, rho=1.0)
    return get_cls_metrics(mask_pred, mask_real)[0]


class GradBucket:
    """Flat gradient storage with bucketed, overlapped all-reduce.

    ``params`` in FORWARD order (``model.parameters()``); buckets are cut in reverse (= backward) order at ``bucket_bytes``.
    ``.grad`` of every parameter is a view into ``self.flat``; ``zero()`` is one memset.  With a process group of more than
    one rank every bucket is all-reduced (SUM) as soon as its last gradient has been accumulated; ``finish()`` waits for the
    reductions and divides by the world size (the mean DistributedDataParallel applies)."""

    def __init__(self, params, bucket_bytes=2 << 20, group=None, overlap=True):
        self.params = [p for p in params if p.requires_grad]
        assert self.params, "no trainable parameters"
        dev, dtype = self.params[0].device, self.params[0].dtype
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, dtype=dtype, device=dev)
        self.views, self.bucket_of, self.buckets = {}, {}, []        # buckets: [start, end, n_params]
        end = total
        cur_end, cur_n = total, 0
        for p in reversed(self.params):                               # backward order: last layers first, at the top of `flat`
            start = end - p.numel()
            self.views[p] = self.flat[start:end].view_as(p)
            self.bucket_of[p] = len(self.buckets)
            cur_n += 1
            end = start
            if (cur_end - end) * self.flat.element_size() >= bucket_bytes or end == 0:
                self.buckets.append((end, cur_end, cur_n))
                cur_end, cur_n = end, 0
        self._pending = [0] * len(self.buckets)
        self._works = []
        self._hooks = []
        self.attach()
        if overlap and self.world > 1:
            for p in self.params:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))

    def attach(self):
        for p in self.params:
            if p.grad is None or p.grad.data_ptr() != self.views[p].data_ptr():
                p.grad = self.views[p]

    def zero(self):
        """``optimizer.zero_grad()`` of the whole model as one memset (and re-attaches views dropped by set_to_none)."""
        self.flat.zero_()
        self.attach()
        self._pending = [n for _, _, n in self.buckets]
        self._works = []

    def _launch(self, b):
        s, e, _ = self.buckets[b]
        self._works.append(dist.all_reduce(self.flat[s:e], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def _on_grad(self, p):
        b = self.bucket_of[p]
        self._pending[b] -= 1
        if self._pending[b] == 0:
            self._launch(b)

    def finish(self):
        """wait for (or, without hooks, launch and wait for) every bucket's all-reduce; gradients become the rank mean."""
        if self.world == 1:
            return
        if not self._hooks:
            for b in range(len(self.buckets)):
                self._launch(b)
        else:
            for b, left in enumerate(self._pending):                  # parameters that received no gradient this step
                if left > 0:
                    self._launch(b)
        for w in self._works:
            w.wait()
        self._works = []
        self.flat.div_(self.world)

    def close(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []


def losses(model_out, n_points, alpha=1.0, beta=1.0):
    """trainer.py:124-131: bce = sum over the three scales of get_bce / candidate count; bpp = bits / input points."""
    bce, bce_list = 0, []
    for out_cls, ground_truth in zip(model_out["out_cls_list"], model_out["ground_truth_list"]):
        curr = get_bce(out_cls, ground_truth) / float(len(out_cls))
        bce = bce + curr
        bce_list.append(curr)
    bpp = get_bits(model_out["likelihood"]) / float(n_points)
    return alpha * bce + beta * bpp, bce, bpp, bce_list


def train_step(model, optimizer, bucket: GradBucket, x, alpha=1.0, beta=1.0):
    """one iteration of trainer.py:117-136 (zero_grad, forward, losses, backward, step) with the gradient all-reduce of
    ``bucket`` overlapped with the backward pass.  Returns the detached (sum_loss, bce, bpp) device scalars."""
    bucket.zero()
    out = model(x, training=True)
    sum_loss, bce, bpp, _ = losses(out, len(x), alpha, beta)
    sum_loss.backward()
    bucket.finish()
    optimizer.step()
    return sum_loss.detach(), bce.detach(), bpp.detach()


def shell_batch(seed, batch=32, size=64, rng_points=12000):
    """config 5 stand-in for a ShapeNet vox64 batch: ``batch`` randomly oriented ellipsoid / box shells in a size^3 grid
     -> (coords int32 [N, 4] with the batch column first, feats float32 [N, 1]); about 4-8 k voxels each."""
    import numpy as np
    rng = np.random.default_rng(seed)
    coords = []
    for b in range(batch):
        q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        radii = rng.uniform(0.22, 0.42, size=3) * size
        if b % 2 == 0:
            u = rng.normal(size=(rng_points, 3))
            u /= np.linalg.norm(u, axis=1, keepdims=True)
        else:                                                         # box shell: points on the faces of a cube
            u = rng.uniform(-1, 1, size=(rng_points, 3))
            ax = rng.integers(0, 3, size=rng_points)
            u[np.arange(rng_points), ax] = np.sign(u[np.arange(rng_points), ax])
        p = np.round((u * radii) @ q.T + size / 2).astype(np.int32)
        p = np.unique(p[((p >= 0) & (p < size)).all(1)], axis=0)
        coords.append(np.concatenate([np.full((len(p), 1), b, np.int32), p], axis=1))
    c = np.concatenate(coords, axis=0)
    return c, np.ones((len(c), 1), np.float32)


def adam_lr():
    return 8e-4                                                       # train.py:21


LN2 = math.log(2.0)
