"""ctypes binding of libpcf_b200.so (the C ABI declared in include/pcf_b200.h).

This is the only place the shared library is touched.  There is NO fallback: if the library is
missing, or a tensor is not a contiguous CUDA tensor, the call raises (the reference raises
RuntimeError through TORCH_CHECK, cpp_wrappers/cpp_pcf_kernel/include/pcf.h:14-24).
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpcf_b200.so")

c_void_p, c_int, c_size_t, c_float, c_int64, c_uint64 = (ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t,
                                                         ctypes.c_float, ctypes.c_int64, ctypes.c_uint64)


class PconvShape(ctypes.Structure):
    _fields_ = [(n, c_int) for n in ("n_in", "n_out", "K", "C_in", "C_add", "C_mid", "C_out", "H")]


_P = c_void_p
# name -> (restype, argtypes); mirrors include/pcf_b200.h one to one
SIGNATURES = {
    "pcfb_last_error": (ctypes.c_char_p, []),
    "pcfb_version": (ctypes.c_char_p, []),
    "pcfb_launch_count": (c_uint64, []),
    "pcfb_set_pdl": (c_int, [c_int]),
    "pcfb_knn_packed": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, _P, _P]),
    "pcfb_knn_grid_workspace": (c_size_t, [c_int, c_int]),
    "pcfb_knn_grid_build": (c_int, [_P, _P, c_int, c_int, c_float, _P, c_size_t, _P]),
    "pcfb_knn_grid_order": (_P, [c_int, c_int, _P]),
    "pcfb_knn_grid_query": (c_int, [_P, c_int, c_int, _P, _P, c_int, c_int, _P, _P, _P, c_size_t, _P]),
    "pcfb_knn_inverse_workspace": (c_size_t, [c_int, c_int, c_int]),
    "pcfb_knn_inverse": (c_int, [_P, c_int, c_int, c_int, _P, _P, _P, _P, c_size_t, _P]),
    "pcfb_gather": (c_int, [_P, _P, c_int, c_int, c_int, c_int, _P, _P]),
    "pcfb_gather_backward": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, _P, _P]),
    "pcfb_gather_max": (c_int, [_P, _P, c_int, c_int, c_int, c_int, _P, _P, _P]),
    "pcfb_gather_max_backward": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, _P, _P]),
    "pcfb_edge_geometry": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, _P, _P, _P]),
    "pcfb_pconv_forward_supported": (c_int, [ctypes.POINTER(PconvShape), c_int]),
    "pcfb_pconv_forward_workspace": (c_size_t, [ctypes.POINTER(PconvShape), c_int]),
    "pcfb_pconv_forward": (c_int, [ctypes.POINTER(PconvShape), _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_size_t, c_int, _P]),
    "pcfb_pconv_backward_workspace": (c_size_t, [ctypes.POINTER(PconvShape), c_int]),
    "pcfb_pconv_backward": (c_int, [ctypes.POINTER(PconvShape)] + [_P] * 19 + [c_size_t, c_int, _P]),
    "pcfb_gemm_nt_workspace": (c_size_t, [c_int, c_int]),
    "pcfb_gemm_nt": (c_int, [_P, c_int, _P, c_int, c_int, _P, _P, c_int, c_int, c_int, c_int, c_int, _P, c_size_t, _P]),
    "pcfb_gemm_nt_prepare": (c_int, [_P, c_int, c_int, c_int, c_int, c_int, _P, c_size_t, _P]),
    "pcfb_gemm_nt_prepared": (c_int, [_P, c_int, _P, _P, c_int, c_int, c_int, c_int, c_int, _P, c_size_t, _P]),
    "pcfb_gemm_tn_workspace": (c_size_t, [c_int, c_int, c_int, c_int]),
    "pcfb_gemm_tn": (c_int, [_P, c_int, _P, c_int, _P, c_int, _P, c_int, c_int, c_int, _P, c_size_t, _P]),
    "pcfb_mlp_supported": (c_int, [c_int, c_int]),
    "pcfb_mlp_workspace": (c_size_t, [c_int64, c_int, c_int]),
    "pcfb_mlp_forward": (c_int, [_P, c_int, c_int64, c_int, c_int, _P, _P, _P, _P, c_int, _P, c_int, _P, _P, _P]),
    "pcfb_bn_finalize": (c_int, [_P, c_int, c_int, c_int64, _P, _P, _P, _P, c_float, c_float, _P, _P, _P, _P, _P, _P, _P, _P,
                                 _P, c_int, c_int, c_int, ctypes.c_double, _P]),
    "pcfb_bn_reduce_sums": (c_int, [_P, c_int, c_int, _P, _P, _P, c_int, c_int, c_int, ctypes.c_double, _P]),
    "pcfb_bn_eval_affine": (c_int, [_P, _P, _P, _P, c_float, c_int, _P, _P, _P, _P]),
    "pcfb_mlp_chain_eval_supported": (c_int, [c_int, c_int, c_int, c_int]),
    "pcfb_mlp_chain_eval": (c_int, [_P, c_int, c_int64, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P, _P, c_int, _P]),
    "pcfb_bn_small_max_rows": (c_int, []),
    "pcfb_bn_small_forward": (c_int, [_P, c_int64, c_int, _P, _P, _P, c_float, c_float, _P, _P, _P, c_int, _P, c_int, _P, _P, _P, _P, _P,
                                      _P, _P, c_int, c_int, c_int, ctypes.c_double, _P]),
    "pcfb_bn_small_backward": (c_int, [_P, _P, c_int64, c_int, _P, _P, _P, _P, c_int, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int,
                                       ctypes.c_double, _P]),
    "pcfb_syncbn_buffer_bytes": (c_size_t, [c_int]),
    "pcfb_syncbn_channels": (c_int, []),
    "pcfb_bn_act": (c_int, [_P, c_int64, c_int, _P, _P, c_int, _P, _P, c_int, _P]),
    "pcfb_mlp_backward_stats": (c_int, [_P, c_int, _P, c_int, c_int64, c_int, _P, _P, _P, _P, c_int, _P, _P, _P, c_size_t, _P]),
    "pcfb_mlp_backward": (c_int, [_P, c_int, _P, c_int, c_int64, c_int, c_int, _P, _P, _P, _P, _P, _P, c_int,
                                  _P, c_int, _P, _P, c_int, _P, _P, _P, c_int, _P, _P, _P, _P, _P, c_size_t, _P]),
    "pcfb_sum_partials": (c_int, [_P, c_int, c_int, _P, _P]),
    "pcfb_bn_supported": (c_int, [c_int]),
    "pcfb_bn_workspace": (c_size_t, [c_int64, c_int]),
    "pcfb_bn_stats": (c_int, [_P, c_int64, c_int, _P, _P, c_size_t, _P, _P]),
    "pcfb_bn_backward_stats": (c_int, [_P, _P, c_int64, c_int, _P, _P, _P, _P, c_int, _P, _P, _P, _P, c_size_t, _P]),
    "pcfb_bn_backward": (c_int, [_P, _P, c_int64, c_int, _P, _P, _P, _P, _P, c_int, _P, _P, _P, _P, _P]),
    "pcfb_gridsub_workspace": (c_size_t, [c_int, c_int, c_int64]),
    "pcfb_gridsub_bounds": (c_int, [_P, _P, c_int, c_int, c_float, _P, _P, _P, c_size_t, _P]),
    "pcfb_gridsub_count": (c_int, [_P, _P, c_int, c_int, c_float, _P, _P, _P, c_int64, _P, _P, c_size_t, _P]),
    "pcfb_gridsub_emit": (c_int, [_P, _P, c_int, c_int, c_int, c_int64, _P, _P, _P, c_size_t, _P]),
    "pcfb_set_point_kernel_max": (c_int, [c_int]),
    "pcfb_voxelize_workspace": (c_size_t, [c_int, c_int, c_int64]),
    "pcfb_voxelize": (c_int, [_P, _P, c_int, c_int, ctypes.c_double, c_int64, _P, _P, _P, _P, c_size_t, _P]),
    "pcfb_pyramid_level_workspace": (c_size_t, [c_int, c_int, c_int64]),
    "pcfb_pyramid_level": (c_int, [_P, _P, c_int, _P, c_int, c_int, c_float, c_int64, _P, _P, _P, _P, _P, c_size_t, _P]),
    "pcfb_guidance_input": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P, _P]),
    "pcfb_guidance_input_backward": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_int, _P, _P, _P]),
    "pcfb_adamw_workspace": (c_size_t, []),
    "pcfb_adamw_clip_step": (c_int, [_P, _P, _P, _P, c_int64, _P, _P, c_float, c_float, c_float, c_float, c_float, _P, _P, c_size_t, _P]),
    "pcfb_ce_workspace": (c_size_t, [c_int64]),
    "pcfb_ce_forward": (c_int, [_P, _P, _P, c_int64, c_int, c_int64, c_float, _P, _P, _P, c_size_t, _P]),
    "pcfb_ce_backward": (c_int, [_P, _P, _P, c_int64, c_int, c_int64, c_float, _P, _P, _P, _P]),
    "pcfb_selftest_umma": (c_int, [_P, _P, _P, c_int, c_int, c_int, _P, c_uint64, c_int, _P, _P]),
}

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(or make -C ml-pointconvformer_b200/csrc).  There is no CPU fallback." % LIB_PATH)
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().pcfb_last_error().decode("utf-8", "replace")
        raise RuntimeError("pcf_b200 %s failed (code %d): %s" % (what, rc, msg))


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    return 0 if t is None else t.data_ptr()


def require(t, dtype, name):
    """The reference's CHECK_INPUT (pcf.h:22-24): CUDA + contiguous, plus the dtype."""
    if not isinstance(t, torch.Tensor):
        raise RuntimeError("%s must be a tensor" % name)
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor (pcf_b200 has no CPU path)" % name)
    if not t.is_contiguous():
        raise RuntimeError("%s must be contiguous" % name)
    if t.dtype != dtype:
        raise RuntimeError("%s must be %s, got %s" % (name, dtype, t.dtype))
    return t


def workspace(nbytes, device):
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


# Algorithmic work of the step at the op boundary (compulsory reads + writes, useful FLOPs), accumulated by the Python wrappers
# while bench.py runs one eager step with ACCOUNT = {"bytes": 0.0, "flops": 0.0} (None = off, the default).
ACCOUNT = None


def account(nbytes, flops=0.0):
    if ACCOUNT is not None:
        ACCOUNT["bytes"] += float(nbytes)
        ACCOUNT["flops"] += float(flops)


def launch_count():
    return int(lib().pcfb_launch_count())
