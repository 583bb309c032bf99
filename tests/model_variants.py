"""The model-level golden variants of tests/golden/make_golden.py (MODEL_VARIANTS there): config overrides on top of
SMALL_MODEL_CFG, shared by the CPU oracle tests and the GPU parity tests.  All variants share model_small.npz's pyramid
and edge tables."""
import os

import numpy as np

SMALL_MODEL_CFG = dict(
    USE_VI=True, USE_PE=True, BATCH_NORM=True, USE_XYZ=True, drop_path_rate=0., dropout_rate=0., dropout_fc=0.,
    attention_type='subtraction', layer_norm_guidance=False, transformer_type='PCF', point_dim=3, num_level=5,
    base_dim=16, feat_dim=[16, 32, 48, 64, 96], mid_dim=[16] * 5, mid_dim_back=1, guided_level=0, num_heads=4,
    resblocks=[0, 1, 2, 1, 1], resblocks_back=[0] * 5, use_level_1=True, num_classes=20)

VARIANTS = {
    "small": dict(),                                                                   # PCF_Normal / configPCF_Opt_10cm, configPCF_5cm
    "lite": dict(mid_dim=[4] * 5, num_heads=8, feat_dim=[16, 32, 64, 64, 96], resblocks=[0, 1, 1, 1, 1]),       # configPCF_10cm_lite
    "ptf2": dict(use_level_1=False, mid_dim_back=3, num_heads=8, feat_dim=[16, 32, 64, 64, 96],
                 resblocks=[0, 1, 2, 1, 1, 1]),                                        # configPCF_2cm_PTF2
    "routing": dict(guided_level=1, resblocks_back=[0, 1, 1, 0, 0], resblocks=[0, 1, 1, 1, 1]),
    # configPCF_Opt_10cm.yaml:27-43 at its real width (BASELINE configs[2]); parameters are NOT stored (21.7 MB) but
    # regenerated from their names by synthetic_state_dict() on both sides
    "normal": dict(base_dim=64, feat_dim=[64, 128, 192, 256, 384], num_heads=8, resblocks=[0, 2, 4, 6, 6]),
}
GRAD_SAMPLES = 1024         # model_normal.npz keeps this many strided entries of every float64 gradient


def grad_sample(flat):
    """The strided sample of a flattened gradient that model_normal.npz stores (same rule on both sides)."""
    n = flat.shape[0]
    return flat[::max(1, n // GRAD_SAMPLES)][:GRAD_SAMPLES]


def synthetic_state_dict(shapes):
    """Deterministic parameters / buffers from their state-dict names and shapes (so the full-width golden need not store
    5.4 M floats): Linear weights ~ U(-1,1)/sqrt(fan_in), biases 0.1 U(-1,1), BatchNorm weight 1 + 0.2 N, bias 0.2 N,
    running_mean 0.2 N, running_var 0.5 + U(0,1), num_batches_tracked 0.  Seeded per tensor by crc32(name)."""
    import zlib
    out = {}
    for name, shape in shapes.items():
        rng = np.random.default_rng(zlib.crc32(name.encode()))
        shape = tuple(int(x) for x in shape)
        if name.endswith("num_batches_tracked"):
            out[name] = np.zeros(shape, np.int64)
        elif name.endswith("running_mean"):
            out[name] = (0.2 * rng.standard_normal(shape)).astype(np.float32)
        elif name.endswith("running_var"):
            out[name] = (0.5 + rng.random(shape)).astype(np.float32)
        elif ".bn." in "." + name:
            out[name] = ((1.0 if name.endswith("weight") else 0.0) + 0.2 * rng.standard_normal(shape)).astype(np.float32)
        elif len(shape) == 2:
            out[name] = ((2 * rng.random(shape) - 1) / np.sqrt(shape[1])).astype(np.float32)
        else:
            out[name] = (0.1 * (2 * rng.random(shape) - 1)).astype(np.float32)
    return out


def cfg_of(variant):
    return dict(SMALL_MODEL_CFG, **VARIANTS[variant])


def load(golden_dir, variant):
    """The variant's golden dict with the shared pyramid / edges of model_small.npz merged in."""
    g = dict(np.load(os.path.join(golden_dir, "model_%s.npz" % variant)))
    if "param_names" in g:                                       # parameters regenerated from their names
        shapes = {str(k): tuple(sh[:nd]) for k, sh, nd in zip(g["param_names"], g["param_shapes"], g["param_ndim"])}
        for k, v in synthetic_state_dict(shapes).items():
            g["param." + k] = v
    if variant != "small":
        base = np.load(os.path.join(golden_dir, "model_small.npz"))
        for k in base.files:
            if k[:2] in ("pc", "es", "ef", "ep") or k.startswith("nrm"):
                g[k] = base[k]
    return g
