"""Two-GPU parity of the SyncBatchNorm path of the fused MLP chain: two ranks with different row counts must produce
the outputs / gradients of ONE process that sees the concatenated batch (torch float64 reference).  Needs >= 2 GPUs
(skipped otherwise; the single-GPU driver box skips it, `scripts/gpu_n2.sh`-style 2-GPU runs execute it)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import pcf_b200  # noqa: F401

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _build(dtype, device, sync):
    torch.manual_seed(5)
    lins = [torch.nn.Linear(12, 16), torch.nn.Linear(16, 16), torch.nn.Linear(16, 32)]
    bns = [torch.nn.BatchNorm1d(16), torch.nn.BatchNorm1d(16), torch.nn.BatchNorm1d(32)]
    for bn in bns:
        torch.nn.init.uniform_(bn.weight, 0.5, 1.5)
        torch.nn.init.uniform_(bn.bias, -0.5, 0.5)
    mods = torch.nn.ModuleList(lins + bns).to(device=device, dtype=dtype)
    if sync:
        mods = torch.nn.SyncBatchNorm.convert_sync_batchnorm(mods)
    return list(mods[:3]), list(mods[3:])


def _worker(rank, world, port, ret, peer):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), PCFB_PEER_REDUCE=peer)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from pcf_b200 import fused_mlp
    g = torch.Generator().manual_seed(11)
    rows = [3000, 1777]
    xs = [torch.randn(n, 12, generator=g) for n in rows]
    gos = [torch.randn(n, 32, generator=g) for n in rows]
    lins, bns = _build(torch.float32, dev, sync=True)
    x = xs[rank].to(dev).requires_grad_(True)
    y = fused_mlp.mlp_chain(x, [(lins[i], bns[i], fused_mlp.ACT_RELU) for i in range(3)], training=True)
    (y * gos[rank].to(dev)).sum().backward()
    out = {"y": y.detach().cpu(), "gx": x.grad.cpu(), "gw0": lins[0].weight.grad.cpu(), "gw2": lins[2].weight.grad.cpu(),
           "gg1": bns[1].weight.grad.cpu(), "gb2": bns[2].bias.grad.cpu()}
    # the wide BatchNorm + ReLU (pcfb_bn_*) under SyncBatchNorm (the global row count travels inside the statistics message)
    torch.manual_seed(7)
    wbn = torch.nn.SyncBatchNorm.convert_sync_batchnorm(torch.nn.BatchNorm1d(96)).to(dev)
    torch.nn.init.uniform_(wbn.weight, 0.5, 1.5)
    xw_all = [torch.randn(1, n, 96, generator=g) * 2 + 0.3 for n in rows]
    gw_all = [torch.randn(1, n, 96, generator=g) for n in rows]
    xw = xw_all[rank].to(dev).requires_grad_(True)
    yw = fused_mlp.bn_act(xw, wbn, fused_mlp.ACT_RELU)
    (yw * gw_all[rank].to(dev)).sum().backward()
    out.update(wy=yw.detach().cpu()[0], wgx=xw.grad.cpu()[0], wgg=wbn.weight.grad.cpu(), wgb=wbn.bias.grad.cpu(),
               wrm=wbn.running_mean.cpu(), wrv=wbn.running_var.cpu())
    gathered = [None] * world
    dist.all_gather_object(gathered, out)
    if rank == 0:
        # single-process float64 reference over the concatenated batch
        rl, rb = _build(torch.float64, dev, sync=False)
        X = torch.cat(xs).to(dev, torch.float64).requires_grad_(True)
        h = X
        for i in range(3):
            h = torch.relu(rb[i](rl[i](h)))
        (h * torch.cat(gos).to(dev, torch.float64)).sum().backward()
        ref = {"y": h.detach().cpu(), "gx": X.grad.cpu(), "gw0": rl[0].weight.grad.cpu(), "gw2": rl[2].weight.grad.cpu(),
               "gg1": rb[1].weight.grad.cpu(), "gb2": rb[2].bias.grad.cpu()}
        errs = {}
        got_y = torch.cat([o["y"] for o in gathered]).double()
        got_gx = torch.cat([o["gx"] for o in gathered]).double()
        errs["y"] = float((got_y - ref["y"]).abs().max() / ref["y"].abs().max())
        errs["gx"] = float((got_gx - ref["gx"]).abs().max() / ref["gx"].abs().max())
        for k in ("gw0", "gw2", "gg1", "gb2"):                     # parameter gradients: sum over ranks = global gradient
            tot = sum(o[k].double() for o in gathered)
            errs[k] = float((tot - ref[k]).abs().max() / ref[k].abs().max())
        torch.manual_seed(7)
        rbn = torch.nn.BatchNorm1d(96).to(dev)
        torch.nn.init.uniform_(rbn.weight, 0.5, 1.5)
        rbn = rbn.double()
        XW = torch.cat([t[0] for t in xw_all]).to(dev, torch.float64).requires_grad_(True)
        hw = torch.relu(rbn(XW))
        (hw * torch.cat([t[0] for t in gw_all]).to(dev, torch.float64)).sum().backward()
        rel = lambda a, b: float((a.double().cpu() - b.cpu()).abs().max() / b.abs().max())
        errs["wy"] = rel(torch.cat([o["wy"] for o in gathered]), hw.detach())
        errs["wgx"] = rel(torch.cat([o["wgx"] for o in gathered]), XW.grad)
        errs["wgg"] = rel(sum(o["wgg"].double() for o in gathered), rbn.weight.grad)
        errs["wgb"] = rel(sum(o["wgb"].double() for o in gathered), rbn.bias.grad)
        errs["wrm"] = rel(gathered[1]["wrm"], rbn.running_mean)
        errs["wrv"] = rel(gathered[1]["wrv"], rbn.running_var)
        ret["errs"] = errs
        ret["peer_active"] = bool(fused_mlp._PEER["state"])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("peer", ["1", "0"])
def test_fused_chain_syncbn_two_ranks_match_single_process(peer):
    """peer "1": statistics exchanged by the one-kernel peer-memory path; "0": the torch.distributed fallback."""
    world, port = 2, _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret, peer), nprocs=world, join=True)
    errs = dict(ret["errs"])
    print(errs, "peer path active:", ret["peer_active"])
    assert ret["peer_active"] == (peer == "1")
    assert errs["y"] < 2e-5 and errs["gx"] < 2e-4, errs
    for k in ("gw0", "gw2", "gg1", "gb2"):
        assert errs[k] < 1e-3, errs                                # fp32 sums behind three train-mode BatchNorms
    assert errs["wy"] < 1e-5 and errs["wgx"] < 1e-4 and errs["wgg"] < 1e-4 and errs["wgb"] < 1e-4, errs
    assert errs["wrm"] < 1e-5 and errs["wrv"] < 1e-5, errs
