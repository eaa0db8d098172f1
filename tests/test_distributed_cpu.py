"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: batch sharding covers the batch exactly,
max-over-ranks timing, and the flat-bucket gradient all-reduce equals the single-process gradient."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from magnet_b200 import dist as md
    r, _, w = md.init_process_group("gloo")
    assert (r, w) == (rank, world)
    lo, hi = md.shard_range(7, rank, world)
    covered = md.sum_over_ranks(hi - lo, device="cpu")
    tmax = md.max_over_ranks(1.0 + rank, device="cpu")
    # flat-bucket gradient averaging == gradient of the mean loss over the full batch
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3, padding=1), torch.nn.ReLU(), torch.nn.Conv2d(4, 2, 1))
    red = md.FlatGradAllReduce(model.parameters())
    red.broadcast_parameters(0)
    x = torch.arange(4 * 3 * 5 * 5, dtype=torch.float32).reshape(4, 3, 5, 5).sin()
    lo4, hi4 = md.shard_range(4, rank, world)
    model(x[lo4:hi4]).square().mean().backward()
    red()
    grads = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    model.zero_grad()
    model(x).square().mean().backward()
    full = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    md.barrier()
    q.put((rank, covered, tmax, float((grads - full).abs().max())))
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_gloo_sharding_and_grad_allreduce():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=100) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    for rank, covered, tmax, gerr in res:
        assert covered == 7 and tmax == 2.0 and gerr < 1e-6
