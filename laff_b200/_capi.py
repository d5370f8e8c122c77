"""ctypes binding of ``include/laff_b200.h`` (the C ABI of the sm_100a kernels).

There is deliberately no fallback: if the shared library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_lib", "liblaff_b200.so")

F16, BF16, F32 = 0, 1, 2
MAX_FEATURES = 8
MAX_TOPK = 16
MAX_TOPK_DENSE = 2048
ACT = {None: 0, False: 0, "none": 0, "tanh": 1, "relu": 2, "sigmoid": 3}
DIRECTION = {"t2i": 0, "i2t": 1, "bidir": 2}


class LaffError(RuntimeError):
    pass


class PoolSource(C.Structure):
    _fields_ = [
        ("kind", C.c_int),
        ("in_dim", C.c_int),
        ("src", C.c_void_p),
        ("ld", C.c_longlong),
        ("bn_scale", C.c_void_p),
        ("bn_shift", C.c_void_p),
    ]


class PoolDesc(C.Structure):
    _fields_ = [
        ("n_features", C.c_int),
        ("heads", C.c_int),
        ("head_dim", C.c_int),
        ("with_ave", C.c_int),
        ("mul", C.c_int),
        ("omega", C.c_float),
        ("norm_eps", C.c_double),
        ("att_weight", C.c_void_p),
        ("att_bias", C.c_void_p),
        ("src", PoolSource * MAX_FEATURES),
    ]


FUSE_MAX_FC, FUSE_MAX_TILED = 4, 2


class FuseFc(C.Structure):
    _fields_ = [("x16", C.c_void_p), ("ldx", C.c_longlong), ("w16", C.c_void_p), ("ldw", C.c_longlong), ("K", C.c_int),
                ("activation", C.c_int), ("bias", C.c_void_p), ("bn_scale", C.c_void_p), ("bn_shift", C.c_void_p)]


class FuseTiled(C.Structure):
    _fields_ = [("x", C.c_void_p), ("ld", C.c_longlong), ("in_dim", C.c_int), ("bn_scale", C.c_void_p),
                ("bn_shift", C.c_void_p)]


class FuseDesc(C.Structure):
    _fields_ = [("n_fc", C.c_int), ("n_tiled", C.c_int), ("heads", C.c_int), ("head_dim", C.c_int), ("dtype", C.c_int),
                ("norm_eps", C.c_double), ("att_weight", C.c_void_p), ("att_bias", C.c_void_p),
                ("fc", FuseFc * FUSE_MAX_FC), ("tiled", FuseTiled * FUSE_MAX_TILED)]


class OptTensor(C.Structure):
    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("grad_out", C.c_void_p), ("state1", C.c_void_p),
                ("state2", C.c_void_p), ("n", C.c_longlong)]


_vp, _i, _ll, _f, _d, _sz = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_double, C.c_size_t
_ull = C.c_ulonglong

# name -> (restype, argtypes); every symbol declared in include/laff_b200.h
SIGNATURES = {
    "laff_last_error": (C.c_char_p, []),
    "laff_abi_version": (_i, []),
    "laff_launch_count": (C.c_longlong, [_i]),
    "laff_set_tuning": (_i, [_i, _i, _i]),
    "laff_set_sm_limit": (_i, [_i]),
    "laff_get_tuning": (_i, [C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    "laff_debug_fuse_profile": (_i, [_vp, _i]),
    "laff_set_fuse_variant": (_i, [_i]),
    "laff_get_fuse_variant": (_i, []),
    "laff_l2norm_quantize": (_i, [_vp, _ll, _i, _i, _ll, _d, _i, _vp, _ll, _vp]),
    "laff_split3_16": (_i, [_vp, _ll, _i, _ll, _i, _i, _vp, _i, _ll, _vp]),
    "laff_cast_pad_16": (_i, [_vp, _ll, _i, _ll, _i, _vp, _i, _ll, _vp]),
    "laff_sim_dense": (_i, [_vp, _vp, _i, _i, _i, _ll, _ll, _i, _f, _vp, _ll, _vp]),
    "laff_sim_collect": (_i, [_vp, _vp, _i, _i, _i, _ll, _ll, _i, _f, _vp, _i, _i, _vp, _vp, _vp, _vp]),
    "laff_sim_collect_rank": (_i, [_vp, _vp, _i, _i, _i, _ll, _ll, _i, _f, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "laff_debug_gemm": (_i, [_vp, _vp, _i, _i, _i, _ll, _ll, _i, _i, _i, _i, _vp, _vp]),
    "laff_sim_gt_workspace_bytes": (_sz, [_i, _i]),
    "laff_sim_gt_scores": (_i, [_vp, _vp, _i, _i, _i, _ll, _ll, _i, _vp, _vp, _vp, _sz, _vp]),
    "laff_sim_rank_workspace_bytes": (_sz, [_i, _i, _i]),
    "laff_sim_rank_topk": (_i, [_vp, _vp, _i, _i, _i, _ll, _ll, _i, _f, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "laff_topk_merge": (_i, [_vp, _vp, _i, _i, _i, _ll, _i, _f, _vp, _vp, _vp]),
    "laff_rank_from_scores": (_i, [_vp, _i, _i, _ll, _vp, _i, _vp, _vp, _vp, _vp]),
    "laff_topk_dense": (_i, [_vp, _ll, _vp, _ll, _i, _ll, _i, _f, _vp, _vp, _vp]),
    "laff_rank_multi_gt": (_i, [_vp, _ll, _i, _ll, _vp, _vp, _vp, _vp]),
    "laff_multi_gt_metrics": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp]),
    "laff_rank_metrics": (_i, [_vp, _i, _vp, _vp]),
    "laff_transform_train_forward": (_i, [_vp, _ll, _i, _i, _i, _f, _ull, _vp, _vp, _vp, _vp, _vp, _f, _f, _i, _vp, _ll, _vp,
                                          _vp, _vp, _vp]),
    "laff_transform_train_backward": (_i, [_vp, _ll, _vp, _ll, _vp, _ll, _i, _vp, _f, _i, _i, _vp, _vp, _vp, _i, _i, _vp, _ll,
                                           _vp, _vp, _vp, _vp]),
    "laff_attention_pool_backward": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _ll, _ll, _f, _i, _i, _f, _vp, _vp, _vp, _vp, _vp,
                                          _vp]),
    "laff_transpose_16": (_i, [_vp, _ll, _i, _i, _i, _i, _i, _vp, _ll, _vp]),
    "laff_fold_tiles": (_i, [_vp, _ll, _i, _i, _i, _vp, _ll, _vp]),
    "laff_frame_pool_backward": (_i, [_vp, _ll, _i, _i, _vp, _vp, _ll, _d, _vp, _vp, _vp, _vp, _vp]),
    "laff_optimizer_blocks": (_i, [_vp, _i, _vp, _vp, _i]),
    "laff_optimizer_step": (_i, [_vp, _vp, _vp, _i, _i, _f, _f, _f, _f, _ll, _f, _vp, _vp, _vp, _vp, _vp]),
    "laff_optimizer_step_scaled": (_i, [_vp, _vp, _vp, _i, _i, _f, _f, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp, _f, _f, _i, _f, _vp,
                                        _vp]),
    "laff_vocab_create": (_vp, [_vp, _vp, _vp, _i]),
    "laff_vocab_destroy": (None, [_vp]),
    "laff_tokenize_lookup": (_ll, [_vp, _vp, _i, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _ll]),
    "laff_bow_project": (_i, [_vp, _vp, _ll, _i, _i, _vp, _ll, _i, _vp, _i, _vp, _vp, _vp, _vp, _ll, _vp]),
    "laff_bow_counts": (_i, [_vp, _vp, _i, _i, _vp, _ll, _vp]),
    "laff_gather_mean": (_i, [_vp, _ll, _ll, _vp, _vp, _i, _i, _vp, _ll, _vp]),
    "laff_gather_rows": (_i, [_vp, _ll, _ll, _vp, _ll, _i, _vp, _ll, _vp]),
    "laff_gru_cell": (_i, [_vp, _ll, _vp, _ll, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "laff_mean_over_length": (_i, [_vp, _vp, _i, _i, _vp]),
    "laff_gru_cell_backward": (_i, [_vp, _ll, _vp, _ll, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _ll, _vp, _ll, _vp]),
    "laff_scatter_add_rows": (_i, [_vp, _ll, _vp, _ll, _i, _ll, _vp, _ll, _vp]),
    "laff_column_sum": (_i, [_vp, _ll, _ll, _i, _vp, _vp]),
    "laff_label_metrics": (_i, [_vp, _i, _i, _ll, _vp, _vp, _vp, _vp]),
    "laff_project": (_i, [_vp, _vp, _ll, _i, _i, _ll, _ll, _i, _vp, _i, _vp, _vp, _vp, _ll, _vp]),
    "laff_bn_fold": (_i, [_vp, _vp, _vp, _vp, _d, _i, _vp, _vp, _vp]),
    "laff_attention_pool": (_i, [C.POINTER(PoolDesc), _ll, _vp, _ll, _vp, _i, _ll, _vp, _vp]),
    "laff_fuse_forward": (_i, [C.POINTER(FuseDesc), _ll, _vp, _ll, _vp, _i, _ll, _vp]),
    "laff_frame_pool": (_i, [_vp, _ll, _i, _i, _vp, _f, _i, _i, _f, _d, _vp, _ll, _vp]),
    "laff_mrl_workspace_bytes": (_sz, [_i, _i, _i]),
    "laff_mrl_forward_backward": (_i, [_vp, _vp, _i, _i, _i, _f, _i, _i, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "laff_mrl_score_forward_backward": (_i, [_vp, _i, _ll, _f, _i, _i, _i, _vp, _vp, _vp]),
    "laff_dsl_forward_backward": (_i, [_vp, _vp, _i, _i, _i, _f, _vp, _vp, _vp, _vp, _sz, _vp]),
}

_lib = None


def lib() -> C.CDLL:
    """Load liblaff_b200.so (built by ``laff_b200.build.build()``); raises if it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LaffError(
                "laff_b200: %s is missing. Build it with `python -m laff_b200.build` (nvcc, sm_100a). "
                "There is no CPU fallback." % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)  # AttributeError if the header and the library disagree
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().laff_last_error()
        raise LaffError("%s failed (rc=%d): %s" % (what or "laff_b200 call", rc, msg.decode() if msg else "?"))


def call(name: str, *args) -> None:
    check(getattr(lib(), name)(*args), name)
