"""Where does the GPU path's fp32 error against the float64 reference come from?  Runs the full-width model golden
(tests/golden/model_normal.npz) with parts of the path swapped for torch ops and prints err_gpu / err_cpu_fp32 statistics."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import pcf_b200  # noqa
import model_variants
from pcf_b200 import fused_mlp, layer_utils, layers, pcf_cuda, model_architecture as MA
import torch.nn.functional as F

g = model_variants.load(os.path.join(ROOT, "tests", "golden"), "normal")
cuda = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()

def run(tag):
    c = dict(model_variants.cfg_of("normal"), USE_CUDA_KERNEL=True, PCONV_OPT=False)
    cfg = MA.get_default_configs(MA.EasyDict(c), c["num_level"], c["base_dim"])
    model = MA.PointConvFormer_Segmentation(cfg).cuda()
    model.load_state_dict({k[6:]: torch.from_numpy(g[k]) for k in g if k.startswith("param.")}, strict=True)
    pcs = [cuda(g["pc%d" % l]) for l in range(5)]; nrm = [cuda(g["nrm%d" % l]) for l in range(5)]
    es = [cuda(g["es%d" % l]) for l in range(5)]; ef = [cuda(g["ef%d" % l]) for l in range(4)]; ep = [cuda(g["ep%d" % l]) for l in range(4)]
    model.train()
    logits = model(cuda(g["feats"]), pcs, es, ef, ep, nrm)
    el = float((logits.cpu() - torch.from_numpy(g["logits_train64"])).abs().max())
    loss = F.cross_entropy(logits[0], cuda(g["target"]), label_smoothing=0.2)
    loss.backward()
    params = dict(model.named_parameters())
    ratios = []
    for k, mx, e32 in zip(g["grad_names"].tolist(), g["grad64_max"].tolist(), g["grad32_err"].tolist()):
        if mx < 1e-12:
            continue
        got = model_variants.grad_sample(params[k].grad.flatten()).cpu().double()
        err = float((got - torch.from_numpy(g["g64." + k]).double()).abs().max())
        ratios.append(err / max(e32, 1e-4 * mx))
    r = np.array(ratios)
    print("%-28s logits err %.2e (cpu %.2e)  grad err/cpu: median %.2f  p90 %.2f  max %.2f" %
          (tag, el, float(g["logits32_err"]), np.median(r), np.percentile(r, 90), r.max()), flush=True)

run("base")
pcf_cuda.FORWARD_VARIANT = 1
run("contraction fwd SIMT fp32")
pcf_cuda.FORWARD_VARIANT = 0
orig_sup = fused_mlp.supported
fused_mlp.supported = lambda dims: False
run("no fused chains")
orig_bn = fused_mlp.bn_supported
fused_mlp.bn_supported = lambda C: False
layers.fused_mlp = fused_mlp
run("no chains, torch BatchNorm")
orig_lin = layer_utils.linear
tl = lambda x, w, b=None: F.linear(x, w, b)
layer_utils.linear = tl; layers.linear = tl
run("+ torch F.linear")
fused_mlp.supported = orig_sup
run("chains on, torch BN+linear")
fused_mlp.bn_supported = orig_bn
layer_utils.linear = orig_lin; layers.linear = orig_lin
fused_mlp.supported = orig_sup
