"""CPU, world_size 2 over gloo: the N>1 host logic of the data-parallel harness (no GPU needed)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vampire_b200.dp import ADJACENT_PARAM_FLOATS, GradBucket, shard_samples


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        bucket = GradBucket("cpu", world)
        beta_grad = torch.tensor(float(rank + 1))            # ranks hold 1.0 and 2.0
        other = torch.full((5, 3), float(10 * (rank + 1)))
        bucket.allreduce_async([beta_grad, other])
        bucket.wait()
        # a second step must not leak state from the first
        g2 = torch.tensor(float(rank))
        bucket.allreduce([g2])
        q.put((rank, beta_grad.item(), other.mean().item(), g2.item(), bucket.flat.numel()))
    finally:
        dist.destroy_process_group()


def test_bucket_allreduce_mean_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, beta, other, g2, numel in res:
        assert beta == pytest.approx(1.5)
        assert other == pytest.approx(15.0)
        assert g2 == pytest.approx(0.5)
        assert numel == ADJACENT_PARAM_FLOATS == 479_543


def test_shard_samples_partition():
    for world in (1, 2, 4, 8):
        seen = sorted(b for r in range(world) for b in shard_samples(16, world, r))
        assert seen == list(range(16))
    assert shard_samples(8, 4, 1) == [1, 5]


def test_bucket_overflow_and_single_rank():
    b = GradBucket("cpu", 1, numel=4)
    g = torch.tensor([1.0, 2.0])
    b.allreduce([g])
    assert torch.equal(g, torch.tensor([1.0, 2.0]))
    with pytest.raises(ValueError):
        b.pack([torch.zeros(5)])
