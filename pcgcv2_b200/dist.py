"""Multi-GPU plumbing of the codec: frames shard one per rank, nothing is exchanged on the data path.

The only collectives (SURVEY.md section 8 e) are an all-gather of a few int64 counters per rank (points,
bits, decoded points) and a max-reduce of the per-rank device time.  Backend: NCCL over NVLink on the
GPU box, gloo in the CPU tests.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def frames_for_rank(n_frames: int, rank: int, world: int):
    """frame i -> rank i mod world (independent units, no halo)."""
    return list(range(rank, n_frames, world))


def gather_counters(counters: torch.Tensor) -> torch.Tensor:
    """[k] int64 per rank -> [world, k] on every rank (identity when not distributed)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return counters.unsqueeze(0)
    out = [torch.zeros_like(counters) for _ in range(dist.get_world_size())]
    dist.all_gather(out, counters)
    return torch.stack(out)


def max_over_ranks(value: float, device="cpu") -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def aggregate_throughput(points_per_rank: torch.Tensor, ms_max: float) -> float:
    """whole-job Mpoints/s = all points coded by all ranks / slowest rank's time."""
    return float(points_per_rank.sum().item()) / (ms_max * 1e-3) / 1e6
