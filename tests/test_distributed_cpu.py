"""Multi-rank host logic on CPU: ray sharding and the single gradient all-reduce of the flat arena (gloo, world 2).
The data path has no other collective."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from monosdf_b200 import training


def test_shard_range_partitions_the_rays():
    for n in (1, 7, 1024, 65536, 262144 + 3):
        for w in (1, 2, 3, 4, 8):
            spans = [training.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)                        # replicas start from identical parameters
        net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Linear(7, 3))
        groups = [list(net[0].parameters()), list(net[1].parameters())]
        arena = training.FlatArena(groups)
        before = {k: v.clone() for k, v in net.state_dict().items()}
        # parameters are views of the arena, names/shapes untouched
        assert all(p.data_ptr() >= arena.flat.data_ptr() for p in net.parameters())
        assert list(net.state_dict().keys()) == list(before.keys())
        # each rank renders its own shard: gradient = (rank + 1) on every element
        arena.zero_grad()
        x = torch.ones(4, 5) * (rank + 1)
        net(x).sum().backward()
        local = arena.grad.clone()
        w = arena.all_reduce()
        assert w == world
        gathered = [torch.zeros_like(local) for _ in range(world)]
        dist.all_gather(gathered, local)
        assert torch.allclose(arena.grad, sum(gathered))
        # gradients landed in the arena slices (autograd accumulated into the views)
        for p, (off, k) in zip(arena.params, arena.slices):
            assert torch.equal(p.grad.reshape(-1), arena.grad[off:off + k])
        ret[rank] = float(arena.grad.sum())
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_flat_arena_all_reduce_gloo_world2():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert len(ret) == world and ret[0] == ret[1]
