"""Model-relevant keys of the reference's shipped YAML configs (/root/reference/configs/*.yaml), restated as
dicts because the reference tree is not available where the benchmarks run.  Only keys that the model /
edge construction read are kept (SURVEY.md section 5 "config / flags")."""

CONFIG_PCF_OPT_10CM = dict(          # configs/configPCF_Opt_10cm.yaml:3-43
    USE_CUDA_KERNEL=True, PCONV_OPT=True, BATCH_NORM=True, USE_XYZ=True, USE_PE=True, sync_bn=True, post_knn=True,
    K_forward=[16] * 5, K_propagate=[16] * 5, K_self=[16] * 5, point_dim=3, num_level=5,
    grid_size=[0.1, 0.2, 0.4, 0.8, 1.6], base_dim=64, feat_dim=[64, 128, 192, 256, 384], mid_dim=[16] * 5,
    mid_dim_back=1, guided_level=0, num_heads=8, resblocks=[0, 2, 4, 6, 6], resblocks_back=[0] * 5,
    label_smoothing=0.2, ignore_label=-100, drop_path_rate=0., dropout_rate=0., dropout_fc=0.,
    layer_norm_guidance=False, num_classes=20, adamw_decay=0.05, learning_rate=0.02, BATCH_SIZE=16)

CONFIG_PCF_10CM_LITE = dict(         # configs/configPCF_10cm_lite.yaml
    CONFIG_PCF_OPT_10CM, PCONV_OPT=False, USE_CUDA_KERNEL=True, mid_dim=[4] * 5, resblocks=[0, 3, 3, 3, 3])

CONFIG_PCF_5CM = dict(               # configs/configPCF_5cm.yaml:23-25 (K = 16 at every level)
    CONFIG_PCF_OPT_10CM, PCONV_OPT=False, grid_size=[0.05, 0.1, 0.2, 0.4, 0.8], learning_rate=0.01, BATCH_SIZE=3)

CONFIG_PCF_2CM_PTF2 = dict(          # configs/configPCF_2cm_PTF2.yaml
    CONFIG_PCF_OPT_10CM, PCONV_OPT=False, use_level_1=False, mid_dim_back=3,
    grid_size=[0.02, 0.06, 0.15, 0.375, 0.9375], drop_path_rate=0., learning_rate=0.01, BATCH_SIZE=2)


def make_cfg(d):
    from .model_architecture import EasyDict, get_default_configs
    cfg = EasyDict(dict(d))
    return get_default_configs(cfg, cfg.num_level, cfg.base_dim)
