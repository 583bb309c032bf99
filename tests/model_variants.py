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
}


def cfg_of(variant):
    return dict(SMALL_MODEL_CFG, **VARIANTS[variant])


def load(golden_dir, variant):
    """The variant's golden dict with the shared pyramid / edges of model_small.npz merged in."""
    g = dict(np.load(os.path.join(golden_dir, "model_%s.npz" % variant)))
    if variant != "small":
        base = np.load(os.path.join(golden_dir, "model_small.npz"))
        for k in base.files:
            if k[:2] in ("pc", "es", "ef", "ep") or k.startswith("nrm"):
                g[k] = base[k]
    return g
