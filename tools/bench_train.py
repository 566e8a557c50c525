"""Config-5 training line: ``python bench.py --workload train [--gpus N] ...`` lands here.

One step = one iteration of trainer.py:117-136 on a synthetic ShapeNet-vox64-like batch (32 shells per rank, Adam lr 8e-4,
alpha = beta = 1): H2D of the batch coordinates, ME.SparseTensor, PCCModel.forward(training), fused BCE/isin + bits losses,
backward (deterministic weight gradients), bucketed gradient all-reduce over NCCL overlapped with the backward pass,
optimizer step.  Weak scaling: every rank trains on its own batch.  Prints one JSON line (rank 0)."""
import json
import os
import sys
import time

import numpy as np


def run(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import pcgcv2_b200
    from pcgcv2_b200 import _lib, train
    from pcgcv2_b200 import dist as pdist
    pcgcv2_b200.install_shims()
    import MinkowskiEngine as ME
    from pcgcv2_b200.model import PCCModel

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    stdout_fd = None
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        sys.stdout.flush()
        stdout_fd = os.dup(1)                                  # NCCL's version banner goes to fd 1: keep stdout for the JSON line
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)                                       # identical replicas
    model = PCCModel().to(dev).train()
    bucket = train.GradBucket(model.parameters())
    opt = torch.optim.Adam(model.parameters(), lr=train.adam_lr(), betas=(0.9, 0.999))
    batches = [train.shell_batch(1000 * rank + i, batch=args.batch) for i in range(4)]
    host = [(torch.from_numpy(c).pin_memory(), torch.from_numpy(f).pin_memory()) for c, f in batches]
    points = int(np.mean([len(c) for c, _ in batches]))

    def step(i):
        c, f = host[i % len(host)]
        x = ME.SparseTensor(features=f.to(dev, non_blocking=True), coordinates=c.to(dev, non_blocking=True), device=dev)
        loss, bce, bpp = train.train_step(model, opt, bucket, x)
        return float(loss)                                     # D2H of the step's result

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    losses = [step(i) for i in range(max(3, args.warmup))]
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0, t0 = _lib.launch_count(), time.time()
    a.record()
    for i in range(args.steps):
        losses.append(step(i))
    b.record()
    barrier()
    ms = pdist.max_over_ranks(a.elapsed_time(b), dev) / args.steps
    launches = _lib.launch_count() - l0
    counters = pdist.gather_counters(torch.tensor([points, args.batch], dtype=torch.int64, device=dev))
    if rank == 0:
        total_pts, total_samples = int(counters[:, 0].sum()), int(counters[:, 1].sum())
        grad_bytes = bucket.flat.numel() * 4
        line = {"metric": "training samples/sec (config 5: vox64 shells, batch 32 per GPU)", "value": round(total_samples / (ms * 1e-3), 2),
                "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": round(ms, 3),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 forward (emulated, see the codec line) / f32 FFMA backward",
                "data": "synthetic",
                "config": {"workload": f"trainer.py:117-136 step on {args.batch} synthetic 64^3 shells per rank (~{points} voxels per batch), "
                                       "Adam lr 8e-4, alpha=beta=1, random-init PCCModel", "voxels_per_step": total_pts,
                           "Mvoxels_per_s": round(total_pts / (ms * 1e-3) / 1e6, 3),
                           "gradient_bytes": grad_bytes, "gradient_buckets": len(bucket.buckets),
                           "parallelism": f"dp{world}: replicas, bucketed NCCL all-reduce launched from the last gradient of each bucket",
                           "loss_first_last": [round(losses[0], 4), round(losses[-1], 4)]},
                "e2e": {"value": round(total_samples / (ms * 1e-3), 2), "unit": "samples/s",
                        "h2d_bytes_per_step": int(np.mean([c.numel() * 4 + f.numel() * 4 for c, f in host])), "d2h_bytes_per_step": 4,
                        "note": "the timed step is already end to end: pinned host batch in, loss scalar out"},
                "gpu_launches": int(launches)}
        if stdout_fd is not None:
            sys.stdout.flush()
            os.dup2(stdout_fd, 1)
        print(json.dumps(line), flush=True)
    bucket.close()
    if world > 1:
        dist.destroy_process_group()
