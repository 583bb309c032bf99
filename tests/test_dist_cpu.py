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
    # the flat-buffer path the bench uses: same mean gradient, one tensor for the optimizer
    lin2 = torch.nn.Linear(4, 3)
    for p in lin2.parameters():
        torch.nn.init.constant_(p, 0.5)
    flat = sharding.FlatParameters(lin2)
    lin2(x).sum().backward()
    g = flat.gather_grads(world)
    flat_ok = (g.numel() == 15 and torch.allclose(g[:12].view(3, 4), lin.weight.grad) and lin2.weight.grad is None
               and lin2.weight.data_ptr() == flat.flat.data_ptr())
    from pcf_b200 import fused_mlp
    rows_ok = True
    t = torch.full((6,), float(rank + 1))
    fused_mlp.sync_all_reduce(t)                                          # gloo: the dist.all_reduce fallback
    rows_ok = rows_ok and bool(torch.equal(t, torch.full((6,), 3.0)))
    gathered = [None] * world
    dist.all_gather_object(gathered, (mine, 0))
    if rank == 0:
        ret["parts"] = [g[0] for g in gathered]
        ret["grad"] = lin.weight.grad.clone()
        ret["flat_ok"] = bool(flat_ok)
        ret["rows_ok"] = bool(rows_ok)
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
    assert ret["flat_ok"]
    assert ret["rows_ok"]


def test_flat_parameters_match_per_tensor_adamw():
    """One AdamW step on the flat buffer equals the per-tensor optimizer (same hyper-parameters, uniform weight decay)."""
    from pcf_b200 import sharding
    torch.manual_seed(3)
    a = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Tanh(), torch.nn.Linear(7, 2))
    b = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Tanh(), torch.nn.Linear(7, 2))
    b.load_state_dict(a.state_dict())
    flat = sharding.FlatParameters(b)
    oa = torch.optim.AdamW(a.parameters(), lr=1e-2, weight_decay=0.05)
    ob = torch.optim.AdamW([flat.flat], lr=1e-2, weight_decay=0.05)
    x = torch.randn(16, 5)
    for _ in range(3):
        oa.zero_grad(set_to_none=True)
        a(x).square().sum().backward()
        torch.nn.utils.clip_grad_norm_(a.parameters(), 1.0)
        oa.step()
        b(x).square().sum().backward()
        flat.gather_grads()
        torch.nn.utils.clip_grad_norm_([flat.flat], 1.0)
        ob.step()
    for pa, pb in zip(a.parameters(), b.parameters()):
        assert torch.allclose(pa, pb, rtol=1e-4, atol=2e-5)      # norm reduction order differs in the last bit
