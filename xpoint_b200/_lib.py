"""ctypes binding of libxpoint_b200.so (the C ABI declared in include/xpoint_b200.h).

There is no CPU fallback anywhere in this package: if the library is missing, or a tensor is not
on a CUDA device, the call raises.  Build the library with ``python __graft_entry__.py`` or
``make -C xpoint_b200/csrc``.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int32, c_int64, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libxpoint_b200.so")

XP_F32, XP_F16, XP_BF16 = 0, 1, 2
XP_OK, XP_ERR_INVALID_ARG, XP_ERR_UNSUPPORTED, XP_ERR_CUDA, XP_ERR_WORKSPACE = 0, -1, -2, -3, -4

_DTYPES = {torch.float32: XP_F32, torch.float16: XP_F16, torch.bfloat16: XP_BF16}


class ScanArgs(Structure):
    """Mirror of xp_scan_args (include/xpoint_b200.h)."""
    _fields_ = (
        [(n, c_void_p) for n in ("u", "delta", "A", "B", "C", "D", "z", "delta_bias", "out", "last_state")]
        + [(n, c_int64) for n in ("batch", "dim", "delta_dim", "groups", "dstate", "seqlen",
                                  "u_batch_stride", "u_dim_stride", "delta_batch_stride", "delta_dim_stride",
                                  "B_batch_stride", "B_group_stride", "B_state_stride",
                                  "C_batch_stride", "C_group_stride", "C_state_stride",
                                  "z_batch_stride", "z_dim_stride", "out_batch_stride", "out_dim_stride")]
        + [(n, c_int32) for n in ("in_dtype", "out_dtype", "delta_softplus", "force_generic")]
        + [("u_group_stride", c_int64), ("u_group_div", c_int64), ("reverse_group_mask", ctypes.c_uint64)]
        + [("dt_weight", c_void_p), ("dt_rank", c_int64), ("dt_group_stride", c_int64)]
    )


class CoreArgs(Structure):
    """Mirror of xp_ss2d_core_args (include/xpoint_b200.h)."""
    _fields_ = (
        [(n, c_void_p) for n in ("xx", "delta", "B", "C", "A", "D", "delta_bias", "out")]
        + [(n, c_int64) for n in ("batch", "d_inner", "dstate", "H", "W", "B_batch_stride", "B_group_stride", "B_state_stride",
                                  "C_batch_stride", "C_group_stride", "C_state_stride")]
        + [(n, c_int32) for n in ("in_dtype", "delta_softplus")]
    )


_SIGNATURES = {
    "xp_abi_version": (ctypes.c_int, []),
    "xp_last_error": (c_char_p, []),
    "xp_check_device": (ctypes.c_int, []),
    "xp_selective_scan_fwd": (ctypes.c_int, [POINTER(ScanArgs), c_void_p]),
    "xp_selective_scan_bwd": (ctypes.c_int, []),
    "xp_cross_scan": (ctypes.c_int, [c_void_p, c_void_p] + [c_int64] * 4 + [c_int32] * 5 + [c_void_p]),
    "xp_cross_merge": (ctypes.c_int, [c_void_p, c_void_p] + [c_int64] * 4 + [c_int32] * 5 + [c_void_p]),
    "xp_merge_norm_gate_workspace_bytes": (c_int64, [c_int64] * 4),
    "xp_merge_norm_gate": (ctypes.c_int, [c_void_p] * 5 + [c_int64] * 4 + [c_int32, c_int32, c_float, c_void_p, c_int64,
                                                                         c_void_p]),
    "xp_ss2d_pack": (ctypes.c_int, [c_void_p, c_void_p] + [c_int64] * 4 + [c_int32, c_void_p]),
    "xp_ss2d_dwconv_pack": (ctypes.c_int, [c_void_p] * 4 + [c_int64] * 5 + [c_int32, c_int32, c_void_p]),
    "xp_ss2d_dt_proj": (ctypes.c_int, [c_void_p] * 3 + [c_int64] * 8 + [c_int32, c_void_p]),
    "xp_ss2d_merge_norm": (ctypes.c_int, [c_void_p] * 5 + [c_int64] * 4 + [c_int32, c_float, c_void_p]),
    "xp_ss2d_core_channels": (c_int32, [c_int64] * 4 + [c_int32]),
    "xp_ss2d_core": (ctypes.c_int, [POINTER(CoreArgs), c_void_p]),
    "xp_ss2d_plane_norm": (ctypes.c_int, [c_void_p] * 5 + [c_int64] * 4 + [c_int32, c_float, c_void_p]),
    "xp_layer_norm": (ctypes.c_int, [c_void_p] * 4 + [c_int64, c_int64, c_int32, c_int32, c_float, c_void_p]),
    "xp_add_layer_norm": (ctypes.c_int, [c_void_p] * 7 + [c_int64, c_int64] + [c_int32] * 4 + [c_float, c_void_p]),
    "xp_patch_embed_stem": (ctypes.c_int, [c_void_p] * 6 + [c_int64] * 5 + [c_float, c_int32, c_int32, c_void_p]),
    "xp_linear_act": (ctypes.c_int, [c_void_p] * 4 + [c_int64] * 3 + [c_int32, c_int32, c_void_p]),
    "xp_linear_res_ln": (ctypes.c_int, [c_void_p] * 8 + [c_int64] * 3 + [c_int32, c_float, c_void_p]),
    "xp_mlp_res_ln": (ctypes.c_int, [c_void_p] * 10 + [c_int64] * 2 + [c_int32, c_float, c_void_p]),
    "xp_detector_post": (ctypes.c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int32, c_int32, c_void_p]),
    "xp_l2_normalize": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int32, c_void_p]),
    "xp_detector_post_cl": (ctypes.c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int32, c_int64, c_int32, c_void_p]),
    "xp_l2_normalize_cl": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int32, c_void_p]),
    "xp_encoder_tail": (ctypes.c_int, [c_void_p] * 4 + [c_int64] * 5 + [c_int32] * 3 + [c_void_p]),
    "xp_nms_workspace_bytes": (c_int64, [c_int64] * 3),
    "xp_box_nms": (ctypes.c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_float, c_float, c_float, c_int64,
                                  c_float, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p]),
    "xp_sample_descriptors": (ctypes.c_int, [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_int32] + [c_int64] * 5
                              + [c_void_p, c_void_p]),
    "xp_match_workspace_bytes": (c_int64, [c_int64] * 4),
    "xp_estimate_homography": (ctypes.c_int, [c_void_p] * 4 + [c_int64] * 4 + [c_int32, c_float, c_int32, ctypes.c_uint32]
                               + [c_void_p] * 4),
    "xp_warp_keypoints": (ctypes.c_int, [c_void_p] * 3 + [c_int64] * 4 + [c_void_p] * 3 + [c_int32, c_void_p]),
    "xp_repeatability_counts": (ctypes.c_int, [c_void_p] * 5 + [c_int64] * 2 + [c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "xp_match_score_counts": (ctypes.c_int, [c_void_p] * 6 + [c_int64] * 2 + [c_void_p, c_int64] + [c_void_p] * 5),
    "xp_mnn_match": (ctypes.c_int, [c_void_p] * 4 + [c_int64] * 4 + [c_void_p] * 5 + [c_int32, c_void_p, c_int64, c_void_p]),
}

_lib = None

# Bookkeeping for bench.py: number of kernels of THIS library launched (each wrapper adds the launches its C-ABI
# call performs) and, when `scan_profile` is a list, (start_event, end_event, algorithmic_bytes, shape) per scan.
launch_count = 0
scan_profile = None


def count_launches(n: int) -> None:
    global launch_count
    launch_count += n


def exported_symbols():
    """Names declared in include/xpoint_b200.h that the library must export."""
    return sorted(_SIGNATURES)


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: the CUDA extension has not been built "
                "(run `python -c 'import __graft_entry__ as g; g.build()'` or `make -C xpoint_b200/csrc`). "
                "xpoint_b200 has no CPU fallback.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)  # raises AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        if handle.xp_abi_version() != 5:
            raise RuntimeError("libxpoint_b200.so ABI version mismatch")
        _lib = handle
    return _lib


def check(rc: int) -> None:
    """Map xp_status to the exceptions the reference raises (TORCH_CHECK -> RuntimeError)."""
    if rc == XP_OK:
        return
    msg = lib().xp_last_error().decode("utf-8", "replace")
    if rc == XP_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise RuntimeError(msg)


def dtype_code(t: torch.Tensor) -> int:
    try:
        return _DTYPES[t.dtype]
    except KeyError:
        raise RuntimeError(f"xpoint_b200: unsupported dtype {t.dtype} (fp32, fp16, bf16 only)") from None


def require_cuda(*tensors) -> torch.device:
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("xpoint_b200: expected a CUDA tensor (this package has no CPU path)")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError("xpoint_b200: all tensors must be on the same device")
    return dev


def stream_ptr(dev: torch.device) -> c_void_p:
    return c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def ptr(t) -> c_void_p:
    return c_void_p(0 if t is None else t.data_ptr())
