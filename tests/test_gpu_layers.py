"""GPU parity of the drop-in layer modules and of the whole segmentation model against golden vectors that
were produced by the UNMODIFIED reference (tests/golden/make_golden.py): same parameters (state-dict keys
are the reference's), same inputs, train-mode forward + backward and eval-mode forward."""
import os

import numpy as np
import pytest
import torch

from gpu_util import cuda, max_err_scaled

pytestmark = pytest.mark.gpu


def _cfg(**kw):
    from pcf_b200.model_architecture import EasyDict
    cfg = EasyDict(USE_VI=True, USE_PE=True, BATCH_NORM=True, USE_CUDA_KERNEL=True, PCONV_OPT=False,
                   drop_path_rate=0., dropout_rate=0., attention_type='subtraction', layer_norm_guidance=False)
    cfg.update(kw)
    return cfg


def _cases():
    from pcf_b200 import layers as L
    c = _cfg()
    cn = _cfg(USE_VI=False, USE_PE=False, BATCH_NORM=False)
    return {
        "pointconv": (lambda: L.PointConv(6, 32, c, [12, 16]), lambda m, a: m(a['xyz'], a['feats'], a['nei'], a['nrm'])),
        "pointconv_single": (lambda: L.PointConv(3, 32, cn, [3, 16]), lambda m, a: m(a['xyz'], a['feats'], a['nei'])),
        "stridepe_self": (lambda: L.PointConvStridePE(32, 32, c, [12, 16]), lambda m, a: m(a['xyz'], a['feats'], a['nei'], a['nrm'])),
        "stridepe_strided": (lambda: L.PointConvStridePE(32, 64, c, [12, 16]),
                             lambda m, a: m(a['xyz'], a['feats'], a['nei'], a['nrm'], a['sxyz'], a['snrm'])),
        "pcf_self": (lambda: L.PCFLayer(64, 64, c, [12, 16], 8), lambda m, a: m(a['xyz'], a['feats'], a['nei'], a['nrm'])),
        "pcf_strided": (lambda: L.PCFLayer(32, 64, c, [12, 16], 8),
                        lambda m, a: m(a['xyz'], a['feats'], a['nei'], a['nrm'], a['sxyz'], a['snrm'])),
        "transpose": (lambda: L.PointConvTransposePE(64, 32, c, [12, 1], [32, 32]),
                      lambda m, a: m(a['sxyz'], a['feats'], a['nei'], a['snrm'], a['xyz'], a['nrm'],
                                     cuda(np.random.default_rng(7).standard_normal((1, 600, 32)).astype(np.float32)))),
        "transpose_mid3": (lambda: L.PointConvTransposePE(64, 32, c, [12, 3], [32, 32]),
                           lambda m, a: m(a['sxyz'], a['feats'], a['nei'], a['snrm'], a['xyz'], a['nrm'])),
    }


# Gradients behind chains of train-mode BatchNorms amplify rounding ~1e4-1e5 x (the reference's own float32 CPU run is off
# by 0.5-1 % of the largest entry on most parameters of the full-width model, tests/golden/model_normal.npz: grad32_err).
# The yardstick is that CPU float32 error on the same draw; the factor is 2 (two correct float32 evaluations differ by
# that much) x 4 (the tensor-core products carry 22 mantissa bits per operand, 3xTF32: unit round-off 2^-22 vs 2^-24).
# scripts/diag_parity.py prints the measured distribution: median 1.9, 90 % below 3.4.
GRAD_FACTOR = 8.0

NAMES = ["pointconv", "pointconv_single", "stridepe_self", "stridepe_strided", "pcf_self", "pcf_strided", "transpose", "transpose_mid3"]


@pytest.mark.parametrize("use_kernel", [True, False])
@pytest.mark.parametrize("name", NAMES)
def test_layer_matches_reference_golden(golden_dir, name, use_kernel):
    g = np.load(os.path.join(golden_dir, "layer_%s.npz" % name))
    ctor, call = _cases()[name]
    layer = ctor().cuda()
    layer.cfg.USE_CUDA_KERNEL = use_kernel
    sd = {k[6:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("param.")}
    layer.load_state_dict(sd, strict=True)                         # reference key names load unchanged
    a = {k: cuda(g[k])[None] for k in ("xyz", "nrm", "sxyz", "snrm", "nei", "feats")}
    a["feats"].requires_grad_(True)
    layer.train()
    y, wni = call(layer, a)
    # targets: the reference module evaluated in float64 (y_train64 / g_feats64 / grad64.*).  The reference's
    # own fp32 evaluation is NOT a usable gradient target: e.g. in layer_pointconv.npz one ReLU pre-activation
    # is ~1e-7, flips in fp32 and moves g_feats by 4% of its max, while fp64 and this GPU path agree.
    torch.testing.assert_close(y.cpu(), torch.from_numpy(g["y_train64"]), rtol=2e-4, atol=2e-4)
    torch.testing.assert_close(y.cpu(), torch.from_numpy(g["y_train"]), rtol=2e-4, atol=2e-4)
    (y * cuda(g["gout"])).sum().backward()
    # gradients: held to the reference's OWN float32 evaluation of the same draw, GRAD_FACTOR * err_cpu_fp32(vs fp64) per
    # tensor, or the reference's own rtol = atol = 1e-4 (test_kernels.py:1756-1764) where the CPU happened to be luckier
    # (this also covers a bias feeding a train-mode BatchNorm, whose true gradient is zero: cancellation noise only).
    def bound(got, ref64, ref32, what):
        e32 = float((ref32.double() - ref64.double()).abs().max())
        mx = float(ref64.abs().max())
        err = float((got.cpu().double() - ref64.double()).abs().max())
        tol = max(GRAD_FACTOR * e32, 1e-4 * (1.0 + mx))
        assert err <= tol, (what, err, e32, mx)
        return err / tol
    worst = bound(a["feats"].grad, torch.from_numpy(g["g_feats64"]), torch.from_numpy(g["g_feats"]), "g_feats")
    for k in g.files:
        if k.startswith("grad64."):
            got = dict(layer.named_parameters())[k[7:]].grad
            worst = max(worst, bound(got, torch.from_numpy(g[k]), torch.from_numpy(g["grad." + k[7:]]), k))
    print("layer_%s use_kernel=%s: worst gradient err / tol = %.2f" % (name, use_kernel, worst))
    layer.load_state_dict(sd, strict=True)
    layer.eval()
    with torch.no_grad():
        ye, _ = call(layer, a)
    torch.testing.assert_close(ye.cpu(), torch.from_numpy(g["y_eval"]), rtol=2e-4, atol=2e-4)


def _build_model(g, variant, pconv_opt, **extra):
    """Our PointConvFormer_Segmentation for a golden variant with the reference's parameters loaded (strict)."""
    import model_variants
    from pcf_b200 import model_architecture as MA
    c = dict(model_variants.cfg_of(variant), USE_CUDA_KERNEL=True, PCONV_OPT=pconv_opt, **extra)
    cfg = MA.get_default_configs(MA.EasyDict(c), c["num_level"], c["base_dim"])
    model = MA.PointConvFormer_Segmentation(cfg).cuda()
    sd = {k[6:]: torch.from_numpy(g[k]) for k in g if k.startswith("param.")}
    if pconv_opt:
        own = set(model.state_dict().keys())
        ren = {}
        for k, v in sd.items():
            k2 = k.replace(".linear.c.", ".pconv_linear_opt.linear.").replace(".linear.bn.", ".bn.")
            ren[k2 if (k2 in own and k not in own) else k] = v
        sd = ren
    model.load_state_dict(sd, strict=True)
    return model, sd


def _model_inputs(g):
    pcs = [cuda(g["pc%d" % l]) for l in range(5)]
    nrm = [cuda(g["nrm%d" % l]) for l in range(5)]
    es = [cuda(g["es%d" % l]) for l in range(5)]
    ef = [cuda(g["ef%d" % l]) for l in range(4)]
    ep = [cuda(g["ep%d" % l]) for l in range(4)]
    return pcs, nrm, es, ef, ep


@pytest.mark.parametrize("pconv_opt", [False, True])
@pytest.mark.parametrize("variant", ["small", "lite", "ptf2", "routing"])
def test_model_matches_reference_golden(golden_dir, variant, pconv_opt):
    """Whole PointConvFormer_Segmentation (small dims) fwd + bwd vs the reference, for the structure of every shipped
    config family (tests/model_variants.py: PCF_Normal 10cm/5cm, 10cm_lite with mid_dim 4, 2cm_PTF2 with
    use_level_1 False + mid_dim_back 3, and the guided_level / resblocks_back branches); with PCONV_OPT the parameters
    are renamed to the reference's other spelling (linear.c -> pconv_linear_opt.linear, linear.bn -> bn)."""
    import model_variants
    g = model_variants.load(golden_dir, variant)
    model, sd = _build_model(g, variant, pconv_opt)
    pcs, nrm, es, ef, ep = _model_inputs(g)
    from pcf_b200 import common_util as CU
    inv = CU.compute_knn_inverse(pcs, es, ef, ep) if pconv_opt else (None, None, None)
    model.train()
    logits = model(cuda(g["feats"]), pcs, es, ef, ep, nrm, *inv)
    torch.testing.assert_close(logits.cpu(), torch.from_numpy(g["logits_train"]), rtol=2e-3, atol=2e-3)
    loss = torch.nn.functional.cross_entropy(logits[0], cuda(g["target"]), label_smoothing=0.2)
    assert abs(loss.item() - float(g["loss"])) < 1e-3
    loss.backward()
    names, norms = g["grad_names"].tolist(), g["grad_norms"].tolist()
    params = dict(model.named_parameters())
    bad = []
    for k, ref in zip(names, norms):
        k2 = k
        if pconv_opt and k not in params:
            k2 = k.replace(".linear.c.", ".pconv_linear_opt.linear.").replace(".linear.bn.", ".bn.")
        got = float(params[k2].grad.norm())
        if abs(got - ref) > 3e-2 * max(ref, 1e-2) + 1e-3:
            bad.append((k, got, ref))
    assert not bad, bad[:5]
    model.load_state_dict(sd, strict=True)
    model.eval()
    with torch.no_grad():
        le = model(cuda(g["feats"]), pcs, es, ef, ep, nrm)
    torch.testing.assert_close(le.cpu(), torch.from_numpy(g["logits_eval"]), rtol=2e-3, atol=2e-3)


@pytest.mark.parametrize("pconv_opt", [False, True])
def test_model_normal_full_width_vs_float64(golden_dir, pconv_opt):
    """BASELINE configs[2]'s own model -- PCF_Normal configPCF_Opt_10cm at its real width (feat_dim 64..384, 8 heads,
    resblocks [0,2,4,6,6], 5.4 M parameters; configs/configPCF_Opt_10cm.yaml:27-43, model_architecture.py:406-502) --
    against the UNMODIFIED reference evaluated in float64 on the same two-scene pyramid (tests/golden/make_golden.py
    --model normal), in both PCONV_OPT spellings.  Yardstick for every gradient: the reference's OWN float32 CPU
    evaluation of the same draw, whose max error against float64 is stored per parameter (grad32_err):
        err_gpu <= GRAD_FACTOR * err_cpu_fp32   (or rtol = atol = 1e-4, the reference's own tolerance, whichever is larger)
    Parameters whose true gradient is zero (a bias in front of a train-mode BatchNorm) carry only rounding noise on
    both sides and fall under the absolute part."""
    import model_variants
    g = model_variants.load(golden_dir, "normal")
    model, sd = _build_model(g, "normal", pconv_opt)
    pcs, nrm, es, ef, ep = _model_inputs(g)
    from pcf_b200 import common_util as CU
    inv = CU.compute_knn_inverse(pcs, es, ef, ep) if pconv_opt else (None, None, None)
    model.train()
    logits = model(cuda(g["feats"]), pcs, es, ef, ep, nrm, *inv)
    ref = torch.from_numpy(g["logits_train64"])
    err_logits = float((logits.cpu() - ref).abs().max())
    assert err_logits <= max(GRAD_FACTOR * float(g["logits32_err"]), 1e-4 * (1.0 + float(ref.abs().max()))), (err_logits, float(g["logits32_err"]))
    loss = torch.nn.functional.cross_entropy(logits[0], cuda(g["target"]), label_smoothing=0.2)
    assert abs(loss.item() - float(g["loss64"])) < 2e-5
    loss.backward()
    params = dict(model.named_parameters())
    bad, over, worst, n_params = [], [], 0.0, 0
    for k, mx, e32 in zip(g["grad_names"].tolist(), g["grad64_max"].tolist(), g["grad32_err"].tolist()):
        k2 = k
        if pconv_opt and k not in params:
            k2 = k.replace(".linear.c.", ".pconv_linear_opt.linear.").replace(".linear.bn.", ".bn.")
        got = model_variants.grad_sample(params[k2].grad.flatten()).cpu().double()
        err = float((got - torch.from_numpy(g["g64." + k]).double()).abs().max())
        tol = max(GRAD_FACTOR * e32, 1e-4 * (1.0 + mx))
        worst = max(worst, err / tol)
        n_params += 1
        if err > 2 * tol:
            bad.append((k, err, e32, mx))
        elif err > tol:
            over.append((k, err, e32, mx))
    print("model_normal pconv_opt=%s: logits err %.2e (cpu fp32 %.2e), worst gradient err / tol %.2f" %
          (pconv_opt, err_logits, float(g["logits32_err"]), worst))
    # every parameter within 2 x the bound, at least 99 % of the ~650 tensors within the bound itself (the amplified rounding
    # noise is heavy-tailed: scripts/diag_parity.py)
    assert not bad, (len(bad), bad[:8])
    assert len(over) <= 0.01 * n_params, (len(over), over[:8])


@pytest.mark.parametrize("variant", ["small", "ptf2"])
def test_bn_folded_inference_matches_reference_golden(golden_dir, variant):
    """replace_batchnorm (util/common_util.py:237-247, the reference's inference path in test_ScanNet_simple.py) folds
    every Linear_BN into a Linear; the folded model must reproduce the reference's eval-mode logits."""
    import model_variants
    from pcf_b200 import common_util as CU
    from pcf_b200.layer_utils import Linear_BN
    g = model_variants.load(golden_dir, variant)
    model, _ = _build_model(g, variant, False)
    pcs, nrm, es, ef, ep = _model_inputs(g)
    model.eval()
    CU.replace_batchnorm(model)
    assert not any(isinstance(m, Linear_BN) for m in model.modules())
    with torch.no_grad():
        le = model(cuda(g["feats"]), pcs, es, ef, ep, nrm)
    torch.testing.assert_close(le.cpu(), torch.from_numpy(g["logits_eval"]), rtol=2e-3, atol=2e-3)


def test_eval_scene_independence(golden_dir):
    """Edges never cross scenes (knn_post_dataloader_utils.py:194-212) and eval-mode BatchNorm is a fixed affine, so the
    logits of a scene must not depend on what it is packed with: run the two packed scenes of the golden pyramid together
    and scene 0 alone."""
    import model_variants
    g = model_variants.load(golden_dir, "small")
    model, _ = _build_model(g, "small", False)
    model.eval()
    pcs, nrm, es, ef, ep = _model_inputs(g)
    with torch.no_grad():
        both = model(cuda(g["feats"]), pcs, es, ef, ep, nrm)
    n = [int(g["stored"][l][0]) for l in range(5)]            # scene 0 comes first at every level: its indices need no shift
    with torch.no_grad():
        alone = model(cuda(g["feats"])[:, :n[0]], [p[:, :n[l]] for l, p in enumerate(pcs)],
                      [e[:, :n[l]].contiguous() for l, e in enumerate(es)],
                      [e[:, :n[l + 1]].contiguous() for l, e in enumerate(ef)],
                      [e[:, :n[l]].contiguous() for l, e in enumerate(ep)], [x[:, :n[l]] for l, x in enumerate(nrm)])
    torch.testing.assert_close(alone, both[:, :n[0]], rtol=1e-5, atol=1e-5)
