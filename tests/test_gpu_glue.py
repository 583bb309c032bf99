"""GPU parity of the layer-glue kernels (csrc/glue.cu) against their torch formulations: the guidance input of the
PointConvFormer layer (/root/reference/layers.py:372-382), clip_grad_norm_ + AdamW on the flat parameter buffer
(train_ScanNet_DDP_WarmUP.py:421-424) and the training loss (train_ScanNet_DDP_WarmUP.py:417)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from gpu_util import cuda

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("use_max", [False, True])
def test_guidance_input_matches_torch(use_max):
    from pcf_b200 import layer_utils, pcf_cuda
    g = torch.Generator().manual_seed(3)
    N, M, K, G, P = 700, (700 if not use_max else 260), 16, 32, 32
    gx = torch.randn(1, N, G, generator=g, dtype=torch.float64)
    pe = torch.randn(1, M, K, P, generator=g, dtype=torch.float64)
    nei = torch.randint(0, N, (1, M, K), generator=g)
    nei[0, 5, 3] = -1                                              # padding entry: gathers zeros, takes no gradient
    go = torch.randn(1, M, K, G + P, generator=g, dtype=torch.float64)
    # torch formulation in float64 (the reference's lines, with index_points' zero rows for padding)
    gx_r, pe_r = gx.clone().requires_grad_(True), pe.clone().requires_grad_(True)
    gathered = gx_r[0][nei[0].clamp(min=0)] * (nei[0] >= 0)[..., None]
    q = torch.cat([gathered[None], pe_r], dim=-1)
    key = q.max(dim=2, keepdim=True)[0] if use_max else q[:, :, :1, :]
    (q - key).backward(go)
    ref = (q - key).detach()
    gx_c, pe_c = cuda(gx.float()).requires_grad_(True), cuda(pe.float()).requires_grad_(True)
    inv = pcf_cuda.compute_knn_inverse(cuda(nei), N)
    out = layer_utils.guidance_input(gx_c, pe_c, cuda(nei), inv, use_max)
    assert float((out.cpu().double() - ref).abs().max()) < 1e-6
    out.backward(cuda(go.float()))
    assert float((gx_c.grad.cpu().double() - gx_r.grad).abs().max()) < 2e-5
    assert float((pe_c.grad.cpu().double() - pe_r.grad).abs().max()) < 1e-5


def test_flat_adamw_matches_torch():
    from pcf_b200 import sharding
    torch.manual_seed(0)
    n = 100003                                                     # not a multiple of 4: the scalar tail runs too
    lin = torch.nn.Linear(n, 1, bias=False).cuda()
    ref_p = torch.nn.Parameter(lin.weight.detach().clone().reshape(-1))
    flat = sharding.FlatParameters(lin)
    ours = sharding.FlatAdamW(flat, lr=2e-3, weight_decay=0.05, max_norm=10.0)
    ref = torch.optim.AdamW([ref_p], lr=2e-3, weight_decay=0.05)
    for it in range(4):
        g = torch.randn(n, device="cuda") * (50.0 if it % 2 == 0 else 0.001)     # one step clips, the next does not
        ref_p.grad = g.clone()
        norm_ref = torch.nn.utils.clip_grad_norm_([ref_p], 10.0)
        ref.step()
        norm = ours.step(g)
        if it == 1:
            for grp in ref.param_groups:
                grp["lr"] = 5e-4
            ours.set_lr(5e-4)                                      # a scheduler step between optimizer steps
        assert abs(float(norm) - float(norm_ref)) <= 1e-5 * float(norm_ref)
        err = float((flat.flat.data - ref_p.data).abs().max())
        assert err < 2e-7, (it, err)
    assert float(ours.step_t) == 4.0


@pytest.mark.parametrize("weighted", [False, True])
def test_cross_entropy_matches_torch(weighted):
    from pcf_b200 import losses
    g = torch.Generator().manual_seed(1)
    N, C = 5000, 20
    logits = (torch.randn(N, C, generator=g) * 3).cuda().requires_grad_(True)
    target = torch.randint(0, C, (N,), generator=g)
    target[::7] = -100
    target = target.cuda()
    w = (0.5 + torch.rand(C, generator=g)).cuda() if weighted else None
    ref_in = logits.detach().double().requires_grad_(True)
    ref = F.cross_entropy(ref_in, target, weight=None if w is None else w.double(), ignore_index=-100, label_smoothing=0.2)
    (ref * 1.7).backward()
    ours = losses.cross_entropy(logits, target, weight=w, ignore_index=-100, label_smoothing=0.2)
    (ours * 1.7).backward()
    assert abs(float(ours) - float(ref)) < 2e-6 * max(1.0, abs(float(ref)))
    assert float((logits.grad.double() - ref_in.grad).abs().max()) < 1e-9 + 2e-6 * float(ref_in.grad.abs().max())
