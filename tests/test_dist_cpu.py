"""world_size-2 gloo test (CPU) of the multi-GPU plumbing: scene sharding is a partition, and the flat-bucket
gradient all-reduce averages gradients across ranks."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import pcf_b200  # noqa: F401


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pcf_b200 import sharding
    sizes = [1000, 400, 900, 300, 50]
    mine = sharding.shard_scenes(sizes, world)[rank]
    lin = torch.nn.Linear(4, 3)
    torch.manual_seed(0)
    for p in lin.parameters():
        torch.nn.init.constant_(p, 0.5)
    x = torch.full((2, 4), float(rank + 1))
    lin(x).sum().backward()
    sharding.allreduce_gradients(lin.parameters(), world)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    if rank == 0:
        ret["parts"] = gathered
        ret["grad"] = lin.weight.grad.clone()
    dist.destroy_process_group()


def test_two_rank_sharding_and_allreduce():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    parts = ret["parts"]
    assert sorted(i for p in parts for i in p) == [0, 1, 2, 3, 4]
    # d/dW sum(W x + b) = sum over rows of x: rank0 -> 2*1, rank1 -> 2*2; mean = 3
    assert torch.allclose(ret["grad"], torch.full((3, 4), 3.0))
