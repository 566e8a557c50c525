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
    res = _spawn_two(_worker)
    assert res[0][1] == [0, 2, 4] and res[1][1] == [1, 3]                  # frame i -> rank i mod world
    for r in res:
        assert r[2] == [[1000, 77, 3], [2000, 78, 2]]                       # every rank sees every counter
        assert r[3] == 15.0                                                 # max over ranks
        assert abs(r[4] - 3000 / 15e-3 / 1e6) < 1e-12                       # whole-job throughput


def test_single_process_is_identity():
    c = torch.tensor([5, 6, 7])
    assert pdist.gather_counters(c).tolist() == [[5, 6, 7]]
    assert pdist.max_over_ranks(3.5) == 3.5


# ---- data-parallel training step (row f4): flat gradient buckets, all-reduce launched from the last gradient of a bucket

def _bucket_worker(rank, world, port, q, overlap):
    from pcgcv2_b200.train import GradBucket
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)                                                    # identical replicas
    model = torch.nn.Sequential(torch.nn.Linear(6, 300), torch.nn.Tanh(), torch.nn.Linear(300, 300), torch.nn.Tanh(),
                                torch.nn.Linear(300, 2))
    bucket = GradBucket(model.parameters(), bucket_bytes=100 << 10, overlap=overlap)
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    g = torch.Generator().manual_seed(100 + rank)                           # every rank its own batch
    grads = None
    for step in range(3):
        x, y = torch.randn(16, 6, generator=g), torch.randn(16, 2, generator=g)
        bucket.zero()
        opt.zero_grad(set_to_none=True)                                     # a trainer's zero_grad must not detach the bucket
        bucket.attach()
        loss = ((model(x) - y) ** 2).mean()
        loss.backward()
        local = bucket.flat.clone() if step == 0 else None
        bucket.finish()
        if step == 0:
            grads = (local, bucket.flat.clone())
        opt.step()
    q.put((rank, len(bucket.buckets), grads[0].tolist(), grads[1].tolist(),
           torch.cat([p.detach().flatten() for p in model.parameters()]).tolist()))
    bucket.close()
    dist.destroy_process_group()


def _spawn_two(target, extra=(), attempts=3):
    """two gloo ranks on a free port; a rendezvous that fails (the port was taken between probing and binding) is retried."""
    last = None
    for _ in range(attempts):
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=target, args=(r, 2, port, q) + tuple(extra)) for r in range(2)]
        for p in procs:
            p.start()
        try:
            res = sorted(q.get(timeout=180) for _ in procs)
            for p in procs:
                p.join(timeout=60)
            if all(p.exitcode == 0 for p in procs):
                return res
            last = [p.exitcode for p in procs]
        except Exception as e:                                   # queue.Empty: a rank died before reporting
            last = e
        for p in procs:
            if p.is_alive():
                p.kill()
            p.join(timeout=10)
    raise AssertionError(f"two-rank run failed {attempts} times: {last}")


def _run_bucket(overlap):
    return _spawn_two(_bucket_worker, (overlap,))


def test_two_rank_gradient_buckets_average_and_keep_replicas_identical():
    for overlap in (True, False):
        (_, nb0, local0, avg0, params0), (_, nb1, local1, avg1, params1) = _run_bucket(overlap)
        assert nb0 == nb1 and nb0 >= 2                                      # several buckets, cut in backward order
        mean = (torch.tensor(local0) + torch.tensor(local1)) / 2
        assert torch.equal(torch.tensor(avg0), mean) and torch.equal(torch.tensor(avg1), mean)
        assert params0 == params1                                           # bit-identical replicas after three steps


def test_gradient_bucket_single_process():
    from pcgcv2_b200.train import GradBucket
    lin = torch.nn.Linear(4, 3)
    bucket = GradBucket(lin.parameters())
    bucket.zero()
    lin(torch.ones(2, 4)).sum().backward()
    bucket.finish()
    assert bucket.flat.numel() == 15 and lin.weight.grad.data_ptr() == bucket.views[lin.weight].data_ptr()
    assert torch.equal(lin.bias.grad, torch.full((3,), 2.0)) and float(bucket.flat.abs().sum()) > 0
