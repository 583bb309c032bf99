"""TEST INFRASTRUCTURE ONLY -- imports the *unmodified* reference modules from /root/reference.

Only usable inside the build container (the GPU box has no /root/reference); used by
tests/golden/make_golden.py to generate the committed golden vectors and by the optional
`tests/test_oracle_vs_reference.py` (skipped when /root/reference is absent).

The reference's `layers.py` / `layer_utils.py` / `model_architecture.py` import three modules that
are not installed here (timm, easydict, pcf_cuda); they are stubbed in sys.modules *before* import
(recipe: SURVEY.md Appendix B).  With USE_CUDA_KERNEL=False / PCONV_OPT=False every layer takes
its pure-torch branch (layers.py:364-366,386-390,686-688,713-719,888-898,1044-1047,1086-1092).
"""
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("PCF_REFERENCE_ROOT", "/root/reference")


class EasyDict(dict):
    """Minimal attr-dict (easydict is not installed); missing keys raise AttributeError."""

    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            self[k] = v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    __setattr__ = dict.__setitem__


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "layers.py"))


def load():
    """Returns (layers, layer_utils, model_architecture) modules of the reference."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    if "timm.models.layers" not in sys.modules:
        tl = types.ModuleType("timm.models.layers")
        tl.DropPath = torch.nn.Identity
        sys.modules.update({"timm": types.ModuleType("timm"),
                            "timm.models": types.ModuleType("timm.models"),
                            "timm.models.layers": tl})
    if "pcf_cuda" not in sys.modules:
        sys.modules["pcf_cuda"] = types.ModuleType("pcf_cuda")
    if "easydict" not in sys.modules:
        ed = types.ModuleType("easydict")
        ed.EasyDict = EasyDict
        sys.modules["easydict"] = ed
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import layers
    import layer_utils
    import model_architecture
    return layers, layer_utils, model_architecture


def load_knn_utils():
    """The reference's knn_post_dataloader_utils.py (unmodified) with its absent third-party imports stubbed
    (pykeops, cuvs, cupy) and `knn_keops` -- whose arithmetic lives in pykeops -- routed to the reference's OWN
    alternative, the sklearn KDTree of `compute_knn(method='sklearn')` (lines 68-72).  Everything else
    (compute_knn's n_ref < K branch, compute_knn_packed's scene x level loop, listToBatch's offsets, tensorize)
    is the reference's code.  KDTree works in float64 with unspecified tie order: use on tie-free clouds only."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    for name, attrs in (("pykeops", {}), ("pykeops.torch", {"LazyTensor": None}), ("cuvs", {}),
                        ("cuvs.neighbors", {"brute_force": None}), ("cupy", {})):
        if name not in sys.modules:
            m = types.ModuleType(name)
            for k, v in attrs.items():
                setattr(m, k, v)
            sys.modules[name] = m
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import knn_post_dataloader_utils as KU

    def knn_keops_via_kdtree(ref_points, query_points, K):
        import numpy as np
        ref = ref_points.numpy() if isinstance(ref_points, torch.Tensor) else np.asarray(ref_points)
        qry = query_points.numpy() if isinstance(query_points, torch.Tensor) else np.asarray(query_points)
        idx = KU.KDTree(ref).query(qry, k=K, return_distance=False)       # the body of the method == 'sklearn' branch
        return torch.from_numpy(idx.astype(np.int64))                     # keops returns an int64 torch tensor
    KU.knn_keops = knn_keops_via_kdtree
    return KU
