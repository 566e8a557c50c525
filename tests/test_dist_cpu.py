"""world_size-2 gloo test of the multi-rank host logic (frames shard per rank; counters all-gather)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pcgcv2_b200 import dist as pdist


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    frames = pdist.frames_for_rank(5, rank, world)
    counters = torch.tensor([1000 * (rank + 1), 77 + rank, len(frames)], dtype=torch.int64)
    allc = pdist.gather_counters(counters)
    ms = pdist.max_over_ranks(10.0 + 5.0 * rank)
    q.put((rank, frames, allc.tolist(), ms, pdist.aggregate_throughput(allc[:, 0], ms)))
    dist.destroy_process_group()


def test_two_rank_counters_and_max_time():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == [0, 2, 4] and res[1][1] == [1, 3]                  # frame i -> rank i mod world
    for r in res:
        assert r[2] == [[1000, 77, 3], [2000, 78, 2]]                       # every rank sees every counter
        assert r[3] == 15.0                                                 # max over ranks
        assert abs(r[4] - 3000 / 15e-3 / 1e6) < 1e-12                       # whole-job throughput


def test_single_process_is_identity():
    c = torch.tensor([5, 6, 7])
    assert pdist.gather_counters(c).tolist() == [[5, 6, 7]]
    assert pdist.max_over_ranks(3.5) == 3.5
