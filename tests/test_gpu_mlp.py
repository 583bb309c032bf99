"""GPU parity of the fused MLP chains (csrc/mlp.cu via fused_mlp.mlp_chain) against the same Linear -> BatchNorm ->
activation sequence evaluated by torch in float64: outputs, input gradient, every parameter gradient, and the
running-statistics update, in train and eval mode."""
import pytest
import torch
import torch.nn as nn

from gpu_util import max_err_scaled

pytestmark = pytest.mark.gpu

CASES = {  # name: (rows, [dims...], [acts...], bn)
    "weightnet": (50001, [12, 8, 8, 16], [1, 1, 1], True),
    "weightnet_decoder": (20000, [12, 8, 8, 1], [1, 1, 1], True),
    "pe_convs": (33333, [3, 16, 16], [1, 1], True),
    "pe_convs_32": (9000, [3, 32, 32], [1, 1], True),
    "mlp_conv": (20000, [12, 32], [1], True),
    "guidance": (20000, [64, 8, 8], [1, 3], True),
    "unary_leaky": (7000, [64, 16], [2], True),
    "unary_noact": (7000, [32, 64], [0], True),
    "no_bn": (5000, [12, 8, 16], [1, 3], False),
    "tiny": (37, [5, 7], [1], True),
}


def build(dims, bn, seed):
    torch.manual_seed(seed)
    mods = []
    for a, b in zip(dims[:-1], dims[1:]):
        lin = nn.Linear(a, b)
        norm = nn.BatchNorm1d(b, momentum=0.1) if bn else None
        if bn:
            with torch.no_grad():
                norm.weight.copy_(1 + 0.3 * torch.randn(b)); norm.bias.copy_(0.3 * torch.randn(b))
                norm.running_mean.copy_(0.2 * torch.randn(b)); norm.running_var.copy_(0.5 + torch.rand(b))
        mods.append((lin, norm))
    return mods


def act_ref(z, a):
    return [lambda t: t, torch.relu, lambda t: torch.nn.functional.leaky_relu(t, 0.1), torch.sigmoid][a](z)


@pytest.mark.parametrize("training", [True, False])
@pytest.mark.parametrize("name", sorted(CASES))
def test_chain_matches_float64(name, training):
    from pcf_b200 import fused_mlp
    rows, dims, acts, bn = CASES[name]
    if not fused_mlp.supported(list(zip(dims[:-1], dims[1:]))):
        pytest.skip("layer size outside the fused kernels' template table (callers fall back to GEMM + BatchNorm)")
    mods = build(dims, bn, sum(map(ord, name)))
    import copy
    ref = copy.deepcopy(mods)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(rows, dims[0], generator=g) * 0.7 + 0.1
    go = torch.randn(rows, dims[-1], generator=g)
    # float64 reference on CPU
    xr = x.double().requires_grad_(True)
    h = xr
    for (lin, norm), a in zip(ref, acts):
        lin.double(); lin.train(training)
        h = lin(h)
        if norm is not None:
            norm.double().train(training)
            h = norm(h)
        h = act_ref(h, a)
    (h * go.double()).sum().backward()
    # fused CUDA path
    for lin, norm in mods:
        lin.cuda()
        if norm is not None:
            norm.cuda().train(training)
    xc = x.cuda().requires_grad_(True)
    out = fused_mlp.mlp_chain(xc, [(lin, norm, a) for (lin, norm), a in zip(mods, acts)], training)
    (out * go.cuda()).sum().backward()
    assert max_err_scaled(out, h) < 2e-5, ("out", max_err_scaled(out, h))
    assert max_err_scaled(xc.grad, xr.grad) < 2e-4, ("dx", max_err_scaled(xc.grad, xr.grad))
    for li, ((lin, norm), (rl, rn)) in enumerate(zip(mods, ref)):
        assert max_err_scaled(lin.weight.grad, rl.weight.grad) < 2e-4, ("dW", li, max_err_scaled(lin.weight.grad, rl.weight.grad))
        if not (norm is not None and training):           # bias before a train-mode BN has zero gradient (noise only)
            assert max_err_scaled(lin.bias.grad, rl.bias.grad) < 2e-4
        if norm is not None:
            # eval mode too: a frozen BatchNorm's affine parameters keep training (dgamma = sum dz*xhat with the running statistics)
            assert max_err_scaled(norm.weight.grad, rn.weight.grad) < 2e-4
            assert max_err_scaled(norm.bias.grad, rn.bias.grad) < 2e-4
            assert max_err_scaled(norm.running_mean, rn.running_mean) < 1e-5
            assert max_err_scaled(norm.running_var, rn.running_var) < 1e-5


def test_chain_follows_each_batchnorm_mode():
    """torch semantics per BatchNorm module: a chain whose BatchNorms alone are put in eval mode (frozen-BN fine-tuning) must
    use the running statistics and leave them untouched, whatever the parent module's mode; momentum=None is the cumulative
    moving average 1 / num_batches_tracked."""
    import copy
    from pcf_b200 import fused_mlp
    mods = build([12, 8, 16], True, 11)
    mods[1][1].momentum = None
    ref = copy.deepcopy(mods)
    x = torch.randn(5000, 12, generator=torch.Generator().manual_seed(2))
    mods[0][1].eval()                                           # layer 0 frozen, layer 1 training with a cumulative average
    ref[0][1].eval()
    for lin, norm in mods:
        lin.cuda(); norm.cuda()
    for rep in range(2):
        h = x.double()
        for lin, norm in ref:
            h = torch.relu(norm.double()(lin.double()(h)))
        out = fused_mlp.mlp_chain(x.cuda(), [(lin, norm, 1) for lin, norm in mods], True)
        assert max_err_scaled(out, h) < 2e-5
    for (lin, norm), (rl, rn) in zip(mods, ref):
        assert max_err_scaled(norm.running_mean, rn.running_mean) < 1e-5 and max_err_scaled(norm.running_var, rn.running_var) < 1e-5
        assert int(norm.num_batches_tracked) == int(rn.num_batches_tracked)


def test_chain_strided_input_and_determinism():
    """The chain input may be a column slice of a wider tensor (pe_convs reads the last 3 VI channels in place)."""
    from pcf_b200 import fused_mlp
    mods = build([3, 16, 16], True, 3)
    for lin, norm in mods:
        lin.cuda(); norm.cuda()
    vi = torch.randn(40000, 12, device="cuda")
    x = vi[:, 9:12]
    spec = [(lin, norm, 1) for lin, norm in mods]
    a = fused_mlp.mlp_chain(x, spec, True)
    b = fused_mlp.mlp_chain(x.contiguous(), spec, True)
    assert torch.equal(a, b)


@pytest.mark.parametrize("training", [True, False])
@pytest.mark.parametrize("act", [0, 1, 2])
@pytest.mark.parametrize("small", [True, False])
@pytest.mark.parametrize("rows,C", [(102095, 32), (5153, 96), (1026, 128), (184, 384), (7, 64), (40000, 20), (300, 20), (2048, 12), (1500, 64)])
def test_bn_act_matches_float64(rows, C, act, training, small, monkeypatch):
    """pcfb_bn_* (the BatchNorm + activation behind the fused contraction and the wide per-point blocks) vs
    torch BatchNorm1d + activation in float64: output, input gradient, dgamma / dbeta, running statistics.
    small: the one-kernel path for tensors of <= pcfb_bn_small_max_rows() rows (pcfb_bn_small_*) on / off."""
    from pcf_b200 import fused_mlp
    if small and (rows > fused_mlp._small_bn_rows() or not training):
        pytest.skip("the one-kernel path only takes small training-mode tensors")
    monkeypatch.setattr(fused_mlp, "SMALL_BN", small)
    assert fused_mlp.bn_supported(C)
    torch.manual_seed(rows + C)
    bn = nn.BatchNorm1d(C, momentum=0.1)
    with torch.no_grad():
        bn.weight.copy_(1 + 0.3 * torch.randn(C)); bn.bias.copy_(0.3 * torch.randn(C))
        bn.running_mean.copy_(0.2 * torch.randn(C)); bn.running_var.copy_(0.5 + torch.rand(C))
    import copy
    ref = copy.deepcopy(bn).double().train(training)
    pivot = torch.randn(C) * 0.5
    x = torch.randn(1, rows, C) * (0.5 + torch.rand(C)) + pivot + 0.1 * torch.randn(C)      # mean near the pivot, like y = xW^T + b
    go = torch.randn(1, rows, C)
    xr = x.double().requires_grad_(True)
    h = act_ref(ref(xr[0]), act)
    (h * go[0].double()).sum().backward()
    bn.cuda().train(training)
    xc = x.cuda().requires_grad_(True)
    out = fused_mlp.bn_act(xc, bn, act, pivot=pivot.cuda() if C != 20 else None)
    assert out.shape == xc.shape
    (out * go.cuda()).sum().backward()
    assert max_err_scaled(out[0], h) < 1e-5, ("out", max_err_scaled(out[0], h))
    assert max_err_scaled(xc.grad, xr.grad) < 1e-4, ("dx", max_err_scaled(xc.grad, xr.grad))
    assert max_err_scaled(bn.weight.grad, ref.weight.grad) < 1e-4
    assert max_err_scaled(bn.bias.grad, ref.bias.grad) < 1e-4
    assert max_err_scaled(bn.running_mean, ref.running_mean) < 1e-5
    assert max_err_scaled(bn.running_var, ref.running_var) < 1e-5
    assert int(bn.num_batches_tracked) == int(ref.num_batches_tracked)


def test_bn_act_deterministic_and_rejects_cpu():
    from pcf_b200 import fused_mlp
    bn = nn.BatchNorm1d(64).cuda().train()
    x = torch.randn(1, 30000, 64, device="cuda")
    a = fused_mlp.bn_act(x, bn, 1)
    b = fused_mlp.bn_act(x, bn, 1)
    assert torch.equal(a, b)
    with pytest.raises(RuntimeError):
        fused_mlp.bn_act(x.cpu(), bn, 1)


@pytest.mark.parametrize("small", [True, False])
@pytest.mark.parametrize("after", [False, True])
def test_bn_act_with_fused_residual_matches_float64(after, small, monkeypatch):
    """act(bn(x) + r) (the tail of a PointConvFormer block, /root/reference/layers.py:413-415) and act(bn(x)) + r (the
    decoder's skip connection, layers.py:1096-1097) in the BatchNorm apply pass, forward and backward, against float64."""
    import copy
    from pcf_b200 import fused_mlp
    monkeypatch.setattr(fused_mlp, "SMALL_BN", small)
    g = torch.Generator().manual_seed(9)
    rows, C = (1001 if small else 3001), 96
    x = torch.randn(1, rows, C, generator=g) * 1.5 + 0.2
    r = torch.randn(1, rows, C, generator=g)
    go = torch.randn(1, rows, C, generator=g)
    bn = torch.nn.BatchNorm1d(C)
    torch.nn.init.uniform_(bn.weight, 0.5, 1.5); torch.nn.init.uniform_(bn.bias, -0.5, 0.5)
    ref = copy.deepcopy(bn).double()
    xr, rr = x.double().requires_grad_(True), r.double().requires_grad_(True)
    z = ref(xr.reshape(rows, C)).reshape(1, rows, C)
    yr = torch.nn.functional.leaky_relu(z, 0.1) + rr if after else torch.nn.functional.leaky_relu(z + rr, 0.1)
    (yr * go.double()).sum().backward()
    bn.cuda()
    xc, rc = x.cuda().requires_grad_(True), r.cuda().requires_grad_(True)
    y = fused_mlp.bn_act(xc, bn, fused_mlp.ACT_LEAKY, residual=rc, residual_after_act=after)
    (y * go.cuda()).sum().backward()
    assert max_err_scaled(y, yr) < 1e-5
    assert max_err_scaled(xc.grad, xr.grad) < 1e-4 and max_err_scaled(rc.grad, rr.grad) < 1e-5
    assert max_err_scaled(bn.weight.grad, ref.weight.grad) < 1e-4 and max_err_scaled(bn.bias.grad, ref.bias.grad) < 1e-4


@pytest.mark.parametrize("c0", [12, 3])
@pytest.mark.parametrize("bn", [True, False])
def test_inference_chain_in_one_kernel(c0, bn, monkeypatch):
    """model.eval() + torch.no_grad(): the WeightNet chain c0 -> 8 -> 8 -> 16 runs as ONE kernel (pcfb_mlp_chain_eval): same
    result as the layer-by-layer kernels to rounding, and as Linear -> BatchNorm(running statistics) -> ReLU in float64;
    with BatchNorm modules and with them folded away (replace_batchnorm leaves bare Linears); strided input rows; and the
    one-kernel path must NOT be taken when a gradient is recorded."""
    import copy
    from pcf_b200 import fused_mlp
    dims, acts = [c0, 8, 8, 16], [1, 1, 1]
    assert fused_mlp.lib().pcfb_mlp_chain_eval_supported(*dims)
    mods = build(dims, bn, 11 + c0)
    ref = copy.deepcopy(mods)
    rows = 40003
    g = torch.Generator().manual_seed(3)
    wide = torch.randn(rows, c0 + 5, generator=g) * 0.7 + 0.1
    h = wide[:, :c0].double()
    for (lin, norm), a in zip(ref, acts):
        h = lin.double()(h)
        if norm is not None:
            h = norm.double().eval()(h)
        h = act_ref(h, a)
    spec = []
    for (lin, norm), a in zip(mods, acts):
        lin.cuda()
        if norm is not None:
            norm.cuda().eval()
        spec.append((lin, norm, a))
    xc = wide.cuda()[:, :c0]                                             # row stride c0 + 5: not contiguous
    with torch.no_grad():
        monkeypatch.setattr(fused_mlp, "CHAIN_EVAL", True)
        one = fused_mlp.mlp_chain(xc, spec)
        monkeypatch.setattr(fused_mlp, "CHAIN_EVAL", False)
        many = fused_mlp.mlp_chain(xc, spec)
    assert one.shape == (rows, 16)
    assert max_err_scaled(one, h) < 1e-5, max_err_scaled(one, h)
    assert max_err_scaled(one, many.double()) < 2e-6                      # same maths, accumulation order may differ
    monkeypatch.setattr(fused_mlp, "CHAIN_EVAL", True)
    xg = wide.cuda()[:, :c0].clone().requires_grad_(True)
    out = fused_mlp.mlp_chain(xg, spec)                                   # gradient recorded: the per-layer path with a backward
    out.sum().backward()
    assert xg.grad is not None and torch.isfinite(xg.grad).all()
