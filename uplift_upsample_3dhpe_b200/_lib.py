"""Builds (nvcc, sm_100a) and loads ``libuu3d.so`` through ctypes.

The shared library is the product; there is no Python/CPU fallback.  ``load()``
raises if the library is missing and cannot be built, and every compute entry
point of the library itself fails without a CUDA device.
"""
from __future__ import annotations

import ctypes
import os
import shutil
import subprocess
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_int64, c_uint8, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "libuu3d.so")
INCLUDE = os.path.join(os.path.dirname(_HERE), "include")
SOURCES = ["kernels_f32.cu", "spatial_tc.cu", "attention_tc.cu", "gemm_tc.cu", "train_kernels.cu", "uu_train.cu",
           "uu_api.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-diag-suppress", "550,177"]

UU_MAX_STRIDED = 8
PRECISION = {"fp32": 0, "bf16": 1}
KINDS = ["gather", "spatial", "token_fill", "layernorm", "attention", "gemm_tc", "gemm_f32", "cast"]


class UUSpec(ctypes.Structure):
    _fields_ = [
        ("n_tok", c_int32), ("n_joints", c_int32), ("d_spatial", c_int32), ("d_temporal", c_int32),
        ("spatial_depth", c_int32), ("temporal_depth", c_int32), ("num_heads", c_int32),
        ("h_spatial", c_int32), ("h_temporal", c_int32), ("n_strided", c_int32),
        ("strides", c_int32 * UU_MAX_STRIDED), ("pad_left", c_int32 * UU_MAX_STRIDED),
        ("pad_right", c_int32 * UU_MAX_STRIDED),
        ("has_strided_input", c_int32), ("first_strided_token_attention_layer", c_int32),
        ("full_output", c_int32),
    ]


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(INCLUDE, "uu3d.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a into libuu3d.so (in-tree, so it travels with the repo)."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libuu3d.so")
    cmd = [nvcc] + NVCC_FLAGS + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", LIB_PATH]
    if verbose:
        print(" ".join(cmd))
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    return LIB_PATH


_lib = None


def load() -> ctypes.CDLL:
    """Load the library (building it first when it is missing and nvcc is present)."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("UU3D_LIB") or LIB_PATH      # UU3D_LIB: development override (A/B builds of the library)
    if path == LIB_PATH and not os.path.exists(LIB_PATH):
        build()
    lib = ctypes.CDLL(path)
    _declare(lib)
    _lib = lib
    return lib


def _declare(lib) -> None:
    P = POINTER
    lib.uu_last_error.restype = c_char_p
    lib.uu_last_error.argtypes = []
    lib.uu_version.restype = c_int
    lib.uu_create.argtypes = [P(UUSpec), c_int, P(c_void_p)]
    lib.uu_destroy.argtypes = [c_void_p]
    lib.uu_set_precision.argtypes = [c_void_p, c_int]
    lib.uu_get_precision.argtypes = [c_void_p]
    lib.uu_weight_count.argtypes = [c_void_p]
    lib.uu_param_count.argtypes = [c_void_p]
    lib.uu_param_count.restype = c_int64
    lib.uu_weight_info.argtypes = [c_void_p, c_int, c_char_p, c_int, P(c_int), P(c_int64), P(c_int)]
    lib.uu_set_weight.argtypes = [c_void_p, c_char_p, c_int, c_void_p, P(c_int64), c_int]
    lib.uu_get_weight.argtypes = [c_void_p, c_char_p, c_int, c_void_p, c_int64]
    lib.uu_forward.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]
    lib.uu_forward_host.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p]
    lib.uu_last_launch_count.argtypes = [c_void_p]
    lib.uu_set_profiling.argtypes = [c_void_p, c_int]
    lib.uu_get_profile.argtypes = [c_void_p, c_void_p, c_void_p, c_int]
    lib.uu_train_config.argtypes = [c_void_p, c_int, c_int, c_float, c_float, P(c_float), c_int, ctypes.c_uint64]
    lib.uu_train_forward_backward.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int64, c_void_p, c_void_p]
    lib.uu_grad_buffer.argtypes = [c_void_p, P(c_void_p), P(c_int64)]
    lib.uu_get_grad.argtypes = [c_void_p, c_char_p, c_int, c_void_p, c_int64]
    lib.uu_get_droppath_scale.argtypes = [c_void_p, c_int, c_int, c_void_p, c_int64, P(c_float)]
    lib.uu_adamw_step.argtypes = [c_void_p, c_float, c_float, c_float, c_float, c_float, c_int64, c_float, c_void_p]
    lib.uu_get_ema_weight.argtypes = [c_void_p, c_char_p, c_int, c_void_p, c_int64]
    lib.uu_stride_mask.argtypes = [c_int, c_int, c_int, c_int64, c_void_p]
    lib.uu_op_build_gather.argtypes = [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.uu_op_token_fill.argtypes = [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.uu_op_layernorm.argtypes = [c_void_p, c_int, c_int, c_void_p, c_void_p, c_float, c_void_p, c_int, c_void_p,
                                    c_int, c_void_p]
    lib.uu_op_attention.argtypes = [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p]
    lib.uu_forward_video.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                     c_void_p]
    lib.uu_forward_video_host.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                                          c_void_p]
    lib.uu_set_flip_order.argtypes = [c_void_p, c_void_p, c_int]
    lib.uu_forward_tta.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]
    lib.uu_forward_video_tta.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                         c_void_p]
    lib.uu_op_keyframe_interp.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]
    lib.uu_op_window_gather.argtypes = [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                        c_void_p, c_void_p, c_void_p]
    lib.uu_op_spatial.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_void_p, P(c_int32), c_void_p]
    lib.uu_op_gemm_f32.argtypes = [c_void_p, c_int64, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p,
                                   c_int64, c_void_p, c_int64, c_void_p]
    lib.uu_op_gemm_bf16.argtypes = [c_void_p, c_int64, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_int,
                                    c_void_p, c_int64, c_void_p, c_int, c_int64, c_void_p]
    lib.uu_op_world_to_cam_and_2d.argtypes = [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]
    lib.uu_op_pose_metrics.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.uu_train_set_math.argtypes = [c_void_p, c_int]
    lib.uu_train_set_token_masking.argtypes = [c_void_p, c_float]
    lib.uu_get_token_mask.argtypes = [c_void_p, c_void_p, c_int64]
    lib.uu_op_gemm_tf32.argtypes = [c_void_p, c_int64, c_int, c_int, c_void_p, c_int64, c_int, c_void_p, c_int, c_void_p,
                                    c_int64, c_void_p, c_int64, c_void_p]
    lib.uu_op_ln_gemm_bf16.argtypes = [c_void_p, c_int, c_int, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_int,
                                       c_int, c_void_p, c_void_p]
    lib.uu_op_resid_gemm_bf16.argtypes = [c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p,
                                          c_int, c_void_p, c_void_p]
    for name in EXPORTS:          # fail at load time, not at first use, if a symbol is missing
        getattr(lib, name)


# every symbol include/uu3d.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "uu_last_error", "uu_version", "uu_create", "uu_destroy", "uu_set_precision", "uu_get_precision",
    "uu_weight_count", "uu_param_count", "uu_weight_info", "uu_set_weight", "uu_get_weight",
    "uu_forward", "uu_forward_host", "uu_forward_video", "uu_forward_video_host", "uu_op_window_gather",
    "uu_set_flip_order", "uu_forward_tta", "uu_forward_video_tta", "uu_op_keyframe_interp", "uu_op_pose_metrics", "uu_op_world_to_cam_and_2d", "uu_last_launch_count", "uu_set_profiling", "uu_get_profile", "uu_stride_mask",
    "uu_train_config", "uu_train_forward_backward", "uu_grad_buffer", "uu_get_grad", "uu_get_droppath_scale",
    "uu_adamw_step", "uu_get_ema_weight", "uu_train_set_math", "uu_train_set_token_masking", "uu_get_token_mask",
    "uu_op_build_gather", "uu_op_token_fill", "uu_op_layernorm", "uu_op_attention", "uu_op_spatial", "uu_op_gemm_f32",
    "uu_op_gemm_bf16", "uu_op_gemm_tf32", "uu_op_ln_gemm_bf16", "uu_op_resid_gemm_bf16",
]


class UUError(RuntimeError):
    pass


def check(rc: int) -> None:
    if rc != 0:
        raise UUError(load().uu_last_error().decode("utf8", "replace"))


def make_spec(spec) -> UUSpec:
    s = UUSpec()
    s.n_tok, s.n_joints = spec.n_tok, spec.n_joints
    s.d_spatial, s.d_temporal = spec.d_spatial, spec.d_temporal
    s.spatial_depth, s.temporal_depth = spec.spatial_depth, spec.temporal_depth
    s.num_heads = spec.num_heads
    s.h_spatial, s.h_temporal = spec.h_spatial, spec.h_temporal
    if len(spec.strides) > UU_MAX_STRIDED:
        raise ValueError("too many strided blocks")
    s.n_strided = len(spec.strides)
    for i, (st, p) in enumerate(zip(spec.strides, spec.paddings)):
        s.strides[i], s.pad_left[i], s.pad_right[i] = st, p[0], p[1]
    s.has_strided_input = int(spec.has_strided_input)
    s.first_strided_token_attention_layer = spec.first_strided_token_attention_layer
    s.full_output = int(spec.full_output)
    return s
